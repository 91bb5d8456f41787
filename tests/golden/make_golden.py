#!/usr/bin/env python
"""Mint golden vectors by running the UNMODIFIED reference loss modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ``/root/reference/src/open_clip/loss.py`` and
``/root/reference/src/models/components/losses.py`` (plus the legacy twin
``src/open_clip_train/spatial_loss.py``) by file path behind a stub ``open_clip``
package (the real ``open_clip/__init__`` needs ftfy/omegaconf/timm, absent here;
recipe from SURVEY.md §8c), feeds them the seeded inputs of
``spatial_clip_b200.synth`` and stores what they return:

    loss per rank, d loss/d image_features, d loss/d text_features,
    d loss_r/d logit_scale, and (single-rank cases) the dense soft-label matrices
    captured from the ``F.normalize(p=1)`` call.

Multi-rank cases use a single-process emulation (monkey-patched
``gather_features`` / ``dist.all_gather``); one case is additionally run for
real over gloo with 2 processes and asserted equal to the emulation.

The fixtures hold only outputs and the generator arguments; inputs are
regenerated from the seed at test time.  Nothing here is imported by the product.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types
from pathlib import Path
from unittest import mock

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from spatial_clip_b200.synth import make_spot_batch, shuffled_text_ids  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    pkg = types.ModuleType("open_clip")
    pkg.__path__ = [str(REF / "src/open_clip")]
    sys.modules["open_clip"] = pkg
    oc_loss = _load("open_clip.loss", REF / "src/open_clip/loss.py")
    pkg.loss = oc_loss
    pkg.ClipLoss = oc_loss.ClipLoss
    ref = _load("ref_losses", REF / "src/models/components/losses.py")
    legacy = _load("ref_legacy_spatial", REF / "src/open_clip_train/spatial_loss.py")
    return oc_loss, ref, legacy


def fake_gather(all_img, all_txt, b):
    """Single-process stand-in for gather_features (loss.py:21-65 semantics)."""

    def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False,
                        rank=0, world_size=1, use_horovod=False):
        if gather_with_grad:
            return all_img, all_txt
        gi = list(all_img.detach().chunk(world_size, dim=0))
        gt = list(all_txt.detach().chunk(world_size, dim=0))
        if not local_loss:
            gi[rank] = image_features
            gt[rank] = text_features
        return torch.cat(gi, dim=0), torch.cat(gt, dim=0)

    return gather_features


def run_spatial(ref_mod, cls_name, batch, scale, world, ctor, logit_bias=None, capture_labels=False, legacy=False,
                text_ids=None):
    n = batch.image_features.shape[0]
    b = n // world
    img = batch.image_features.clone().requires_grad_(True)
    txt = batch.text_features.clone().requires_grad_(True)
    s = torch.tensor(float(scale), dtype=torch.float32, requires_grad=True)
    ids = batch.tile_ids
    losses, ds = [], []
    captured = []
    real_normalize = torch.nn.functional.normalize

    def spy_normalize(x, p=2.0, dim=1, **kw):
        if p == 1:
            captured.append(x.detach().clone().numpy())
        return real_normalize(x, p=p, dim=dim, **kw)

    gather_calls = [0]

    def fake_all_gather(lst, t):
        # the reference gathers the image-side ids first, then the text-side ids (losses.py:63-68)
        src = ids if (text_ids is None or gather_calls[0] % 2 == 0) else text_ids
        gather_calls[0] += 1
        for q in range(world):
            lst[q] = src[q * b:(q + 1) * b].clone()

    total = 0.0
    with mock.patch.object(ref_mod, "gather_features", fake_gather(img, txt, b)), \
            mock.patch.object(ref_mod.dist, "all_gather", fake_all_gather), \
            mock.patch.object(ref_mod.F, "normalize", spy_normalize):
        for r in range(world):
            sl = slice(r * b, (r + 1) * b)
            m = getattr(ref_mod, cls_name)(rank=r, world_size=world, **ctor)
            m.rank, m.world_size = r, world
            if legacy:
                out = m(img[sl], txt[sl], ids[sl], ids[sl], batch.neighbor_tile_ids[sl],
                        batch.neighbor_alphas[sl], s, logit_bias, output_dict=True)
            else:
                out = m(image_features=img[sl], text_features=txt[sl], logit_scale=s,
                        image_tile_ids=ids[sl], text_tile_ids=(ids if text_ids is None else text_ids)[sl],
                        neighbor_tile_ids=batch.neighbor_tile_ids[sl],
                        neighbor_alphas=batch.neighbor_alphas[sl], logit_bias=logit_bias)
            loss_r = out["contrastive_loss"]
            ds.append(float(torch.autograd.grad(loss_r, s, retain_graph=True)[0]))
            losses.append(float(loss_r.detach()))
            total = total + loss_r
    total.backward()
    res = dict(loss=np.array(losses, np.float64), d_scale=np.array(ds, np.float64),
               d_image=img.grad.numpy().copy(), d_text=txt.grad.numpy().copy())
    if capture_labels and world == 1:
        res["labels_i_t"] = captured[0]
        res["labels_t_i"] = captured[1]
    return res


def run_clip(oc_loss, ref_mod, batch, scale, world, ctor, logit_bias=None):
    n = batch.image_features.shape[0]
    b = n // world
    img = batch.image_features.clone().requires_grad_(True)
    txt = batch.text_features.clone().requires_grad_(True)
    s = torch.tensor(float(scale), dtype=torch.float32, requires_grad=True)
    losses, ds = [], []
    total = 0.0
    with mock.patch.object(oc_loss, "gather_features", fake_gather(img, txt, b)):
        for r in range(world):
            sl = slice(r * b, (r + 1) * b)
            m = ref_mod.ClipLoss(rank=r, world_size=world, **ctor)
            out = m(image_features=img[sl], text_features=txt[sl], logit_scale=s, logit_bias=logit_bias)
            loss_r = out["contrastive_loss"]
            ds.append(float(torch.autograd.grad(loss_r, s, retain_graph=True)[0]))
            losses.append(float(loss_r.detach()))
            total = total + loss_r
    total.backward()
    return dict(loss=np.array(losses, np.float64), d_scale=np.array(ds, np.float64),
                d_image=img.grad.numpy().copy(), d_text=txt.grad.numpy().copy())


def _gloo_worker(rank, world, gen, scale, ctor, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, ref, _ = load_reference()
    batch = make_spot_batch(**gen)
    loc = batch.rank_slice(rank, world)
    img = loc.image_features.clone().requires_grad_(True)
    txt = loc.text_features.clone().requires_grad_(True)
    s = torch.tensor(float(scale), requires_grad=True)
    m = ref.SpatialLoss(**ctor)  # picks rank/world from the initialised process group
    loss = m(img, txt, s, loc.tile_ids, loc.tile_ids, loc.neighbor_tile_ids, loc.neighbor_alphas)["contrastive_loss"]
    loss.backward()
    q.put((rank, float(loss.detach()), img.grad.numpy(), txt.grad.numpy(), float(s.grad)))
    dist.barrier()
    dist.destroy_process_group()


def run_spatial_gloo(gen, scale, world, ctor, port=29731):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, gen, scale, ctor, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join()
    return dict(loss=np.array([g[1] for g in got]), d_image=np.concatenate([g[2] for g in got]),
                d_text=np.concatenate([g[3] for g in got]), d_scale=np.array([g[4] for g in got]))


SPATIAL_DEFAULT = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
                       neighbor_alpha_scale=0.5, float32_logits=True)  # configs/loss/spatial.yaml:6-11


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    oc_loss, ref, legacy = load_reference()
    cases = {}

    def add(name, kind, gen, scale, world, ctor, res, **extra):
        meta = dict(kind=kind, gen=gen, scale=scale, world=world, ctor=ctor, **extra)
        np.savez_compressed(OUT / f"{name}.npz", meta=json.dumps(meta), **res)
        cases[name] = meta
        print(f"{name:38s} loss={res['loss']}")

    if len(sys.argv) > 1 and sys.argv[1] == "--asym-only":
        # later addition, minted on its own so that the earlier fixtures (zip timestamps) stay byte-identical:
        # image-side and text-side tile ids differ (world 1: the id all-gather is not exercised)
        gasym = dict(n=96, d=64, k=8, seed=1008, dup_frac=0.05, self_loops=True)
        ba = make_spot_batch(**gasym)
        add("spatial_n96_asym_text_ids", "spatial", gasym, 30.0, 1, SPATIAL_DEFAULT,
            run_spatial(ref, "SpatialLoss", ba, 30.0, 1, SPATIAL_DEFAULT, capture_labels=True,
                        text_ids=shuffled_text_ids(ba.tile_ids, 77)), text_ids_seed=77)
        idx = json.loads((OUT / "index.json").read_text())
        idx.update(cases)
        (OUT / "index.json").write_text(json.dumps(idx, indent=1, sort_keys=True))
        return

    if len(sys.argv) > 1 and sys.argv[1] == "--asym-w2-only":
        # as above, over two ranks: the TEXT-side ids travel through the second id all-gather (losses.py:63-68)
        gasym = dict(n=192, d=64, k=8, seed=1009, dup_frac=0.05, self_loops=True)
        ba = make_spot_batch(**gasym)
        add("spatial_n192_w2_asym_text_ids", "spatial", gasym, 30.0, 2, SPATIAL_DEFAULT,
            run_spatial(ref, "SpatialLoss", ba, 30.0, 2, SPATIAL_DEFAULT,
                        text_ids=shuffled_text_ids(ba.tile_ids, 78)), text_ids_seed=78)
        idx = json.loads((OUT / "index.json").read_text())
        idx.update(cases)
        (OUT / "index.json").write_text(json.dumps(idx, indent=1, sort_keys=True))
        return

    g64 = dict(n=64, d=64, k=8, seed=1001)
    # cfg1 smoke shape: N=64, K=8 and K=6, hydra defaults
    add("spatial_n64_k8_default", "spatial", g64, 1 / 0.07, 1, SPATIAL_DEFAULT,
        run_spatial(ref, "SpatialLoss", make_spot_batch(**g64), 1 / 0.07, 1, SPATIAL_DEFAULT, capture_labels=True))
    g64k6 = dict(n=64, d=128, k=6, seed=1002)
    add("spatial_n64_k6_capactive", "spatial", g64k6, 55.0, 1, SPATIAL_DEFAULT,
        run_spatial(ref, "SpatialLoss", make_spot_batch(**g64k6), 55.0, 1, SPATIAL_DEFAULT, capture_labels=True))
    nocap = dict(SPATIAL_DEFAULT, cap_logit_scale=None, temp_reg_weight=0.0, neighbor_alpha_scale=1.0)
    add("spatial_n64_nocap_noreg_s100", "spatial", g64, 100.0, 1, nocap,
        run_spatial(ref, "SpatialLoss", make_spot_batch(**g64), 100.0, 1, nocap, capture_labels=True))
    # edge cases of the mask: duplicate ids, self loops, negative alphas
    gedge = dict(n=96, d=64, k=8, seed=1003, dup_frac=0.08, self_loops=True, negative_alphas=True)
    add("spatial_n96_edges", "spatial", gedge, 30.0, 1, SPATIAL_DEFAULT,
        run_spatial(ref, "SpatialLoss", make_spot_batch(**gedge), 30.0, 1, SPATIAL_DEFAULT, capture_labels=True))
    # ragged tiny batch (validation tail), K=0-like all padding handled by valid count 0 rows
    gtiny = dict(n=5, d=64, k=8, seed=1004)
    add("spatial_n5_tiny", "spatial", gtiny, 1 / 0.07, 1, SPATIAL_DEFAULT,
        run_spatial(ref, "SpatialLoss", make_spot_batch(**gtiny), 1 / 0.07, 1, SPATIAL_DEFAULT, capture_labels=True))
    # scalar logit_bias (cancels in both softmaxes)
    add("spatial_n64_bias", "spatial", g64, 20.0, 1, SPATIAL_DEFAULT,
        run_spatial(ref, "SpatialLoss", make_spot_batch(**g64), 20.0, 1, SPATIAL_DEFAULT,
                    logit_bias=torch.tensor(-3.5)), logit_bias=-3.5)
    # legacy twin, positional order (spatial_loss.py:37-48)
    leg_ctor = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
                    float32_logits=True, neighbor_alpha_scale=0.5)
    add("legacy_n64_k8_default", "spatial", g64, 1 / 0.07, 1, leg_ctor,
        run_spatial(legacy, "GlobalMappingMultiPositiveClipLoss", make_spot_batch(**g64), 1 / 0.07, 1, leg_ctor,
                    legacy=True))
    # multi-rank emulation
    g256 = dict(n=256, d=128, k=8, seed=1005, dup_frac=0.01)
    for world in (2, 4, 8):
        add(f"spatial_n256_w{world}", "spatial", g256, 55.0, world, SPATIAL_DEFAULT,
            run_spatial(ref, "SpatialLoss", make_spot_batch(**g256), 55.0, world, SPATIAL_DEFAULT))
    for ll, gwg in ((True, False), (False, False), (False, True)):
        ctor = dict(SPATIAL_DEFAULT, local_loss=ll, gather_with_grad=gwg)
        add(f"spatial_n256_w4_ll{int(ll)}_gwg{int(gwg)}", "spatial", g256, 25.0, 4, ctor,
            run_spatial(ref, "SpatialLoss", make_spot_batch(**g256), 25.0, 4, ctor))
    # real 2-process gloo run of the reference == emulation
    emu = run_spatial(ref, "SpatialLoss", make_spot_batch(**g256), 55.0, 2, SPATIAL_DEFAULT)
    real = run_spatial_gloo(g256, 55.0, 2, SPATIAL_DEFAULT)
    for key in ("loss", "d_image", "d_text", "d_scale"):
        err = np.abs(emu[key] - real[key]).max()
        assert err < 5e-6, (key, err)
        print(f"gloo-vs-emulation {key}: max abs diff {err:.2e}")
    add("spatial_n256_w2_gloo", "spatial", g256, 55.0, 2, SPATIAL_DEFAULT, real, source="gloo 2 procs")

    # ClipLoss (configs/loss/clip.yaml:6-8) and the other flag combinations
    gclip = dict(n=128, d=128, k=0, seed=1006)
    cl = dict(local_loss=True, gather_with_grad=True, cache_labels=True)
    for scale in (1 / 0.07, 100.0):
        add(f"clip_n128_w1_s{int(scale)}", "clip", gclip, scale, 1, cl,
            run_clip(oc_loss, ref, make_spot_batch(**gclip), scale, 1, cl))
    for world in (2, 4):
        for ll in (True, False):
            for gwg in (True, False):
                ctor = dict(local_loss=ll, gather_with_grad=gwg, cache_labels=False)
                add(f"clip_n128_w{world}_ll{int(ll)}_gwg{int(gwg)}", "clip", gclip, 10.0, world, ctor,
                    run_clip(oc_loss, ref, make_spot_batch(**gclip), 10.0, world, ctor))
    add("clip_n128_bias", "clip", gclip, 10.0, 1, cl,
        run_clip(oc_loss, ref, make_spot_batch(**gclip), 10.0, 1, cl, logit_bias=torch.tensor(2.0)), logit_bias=2.0)

    # a tile-sized case for the GPU kernels (full 128-row MMA tiles + ragged tail)
    g300 = dict(n=300, d=256, k=8, seed=1007, dup_frac=0.01, self_loops=True)
    add("spatial_n300_d256", "spatial", g300, 55.0, 1, SPATIAL_DEFAULT,
        run_spatial(ref, "SpatialLoss", make_spot_batch(**g300), 55.0, 1, SPATIAL_DEFAULT))
    add("clip_n300_d256", "clip", dict(g300, k=0), 1 / 0.07, 1, cl,
        run_clip(oc_loss, ref, make_spot_batch(**dict(g300, k=0)), 1 / 0.07, 1, cl))

    (OUT / "index.json").write_text(json.dumps(cases, indent=1, sort_keys=True))
    print(f"wrote {len(cases)} fixtures to {OUT}")


if __name__ == "__main__":
    main()
