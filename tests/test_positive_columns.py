"""Data-side soft-target resolution: spatial_clip_b200.positives vs the oracle's restatement of the
reference's label loop (bit-exact), and SpatialLossFromColumns vs the reference goldens (host logic, CPU)."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden_names, load_golden, text_ids_for
from emulated_ops import EmulatedOps
from oracle.contrastive_oracle import dense_labels, soft_label_triples
from spatial_clip_b200 import SpatialLossFromColumns, losses
from spatial_clip_b200.positives import collate_positive_columns, resolve_positive_columns
from spatial_clip_b200.synth import make_spot_batch


@pytest.mark.parametrize("seed", range(40))
def test_columns_weights_probs_bit_exact_vs_oracle(seed):
    r = random.Random(seed)
    world, b, k = r.choice([1, 2, 4]), r.choice([1, 3, 16, 50]), r.choice([0, 1, 6, 8])
    bt = make_spot_batch(n=world * b, d=64, k=k, seed=seed, dup_frac=r.choice([0, 0.1, 0.5]),
                         self_loops=r.random() < 0.5, negative_alphas=r.random() < 0.5)
    scale = r.choice([1.0, 0.5, 2.0])
    for rank in range(world):
        sl = slice(rank * b, (rank + 1) * b)
        rows, sums = soft_label_triples(bt.tile_ids.numpy(), bt.neighbor_tile_ids[sl].numpy(),
                                        bt.neighbor_alphas[sl].numpy(), scale, rank)
        col, w, q = resolve_positive_columns(bt.tile_ids, bt.neighbor_tile_ids[sl], bt.neighbor_alphas[sl], scale, rank)
        assert col.dtype == torch.int32 and w.dtype == torch.float32 and tuple(col.shape) == (b, k + 1)
        for i, lst in enumerate(rows):
            want_c = np.array([c for c, _ in lst], dtype=np.int32)
            want_w = np.array([x for _, x in lst], dtype=np.float32)
            want_q = want_w * (np.float32(1.0) / max(sums[i], np.float32(1e-12)))
            n = len(lst)
            assert np.array_equal(col[i, :n].numpy(), want_c)
            assert np.array_equal(w[i, :n].numpy().view(np.uint32), want_w.view(np.uint32))
            assert np.array_equal(q[i, :n].numpy().view(np.uint32), want_q.astype(np.float32).view(np.uint32))
            assert (col[i, n:] == -1).all() and (w[i, n:] == 0).all() and (q[i, n:] == 0).all()


@pytest.mark.parametrize("name", [n for n in golden_names("spatial", world=1) if "bias" not in n and "legacy" not in n])
def test_weights_reproduce_reference_dense_labels(name):
    """The un-normalised weights ARE the reference's dense label rows (captured before F.normalize), bit for bit."""
    meta, gold = load_golden(name)
    if "labels_i_t" not in gold:
        pytest.skip("no dense labels stored")
    b = make_spot_batch(**meta["gen"])
    scale = meta["ctor"].get("neighbor_alpha_scale", 1.0)
    for id_map, key in ((text_ids_for(meta, b), "labels_i_t"), (b.tile_ids, "labels_t_i")):
        col, w, _ = resolve_positive_columns(id_map, b.neighbor_tile_ids, b.neighbor_alphas, scale, 0)
        rows = [[(int(c), np.float32(x)) for c, x in zip(cr.tolist(), wr.tolist()) if c >= 0] for cr, wr in zip(col, w)]
        assert np.array_equal(dense_labels(rows, len(b.tile_ids)).view(np.uint32), gold[key].view(np.uint32)), key


def _run(meta, rank, world, ops_calls=None):
    full = make_spot_batch(**meta["gen"])
    b = full.rank_slice(rank, world)
    bl = b.tile_ids.shape[0]
    c = dict(meta["ctor"])
    c.pop("cache_labels", None)
    scale = c.get("neighbor_alpha_scale", 1.0)
    txt_ids = text_ids_for(meta, full)
    it = resolve_positive_columns(txt_ids, b.neighbor_tile_ids, b.neighbor_alphas, scale, rank)
    kw = {}
    if "text_ids_seed" in meta:  # image / text id vectors differ: the text rows resolve in the IMAGE id map
        ti = resolve_positive_columns(full.tile_ids, b.neighbor_tile_ids, b.neighbor_alphas, scale, rank)
        kw = dict(positive_columns_text=ti[0], positive_probs_text=ti[2])
    img = b.image_features.clone().requires_grad_(True)
    txt = b.text_features.clone().requires_grad_(True)
    s = torch.tensor(float(meta["scale"]), requires_grad=True)
    mod = SpatialLossFromColumns(rank=rank if world > 1 else None, world_size=world if world > 1 else None, **c)
    out = mod(image_features=img, text_features=txt, logit_scale=s, positive_columns=it[0], positive_probs=it[2],
              positive_weights=it[1], **kw)
    assert set(out) == {"contrastive_loss"} and bl == img.shape[0]
    out["contrastive_loss"].backward()
    return float(out["contrastive_loss"].detach()), img.grad.numpy(), txt.grad.numpy(), float(s.grad)


def _assert_close(gold, rank, b, loss, gi, gt, ds, scale):
    np.testing.assert_allclose(loss, gold["loss"][rank], rtol=3e-6, atol=2e-6 + 2e-7 * scale)
    np.testing.assert_allclose(ds, gold["d_scale"][rank], rtol=3e-4, atol=2e-6)
    sl = slice(rank * b, (rank + 1) * b)
    floor = 3e-6 * scale * 0.5 / b
    for got, ref in ((gi, gold["d_image"][sl]), (gt, gold["d_text"][sl])):
        assert np.abs(got - ref).max() <= 3e-5 * np.abs(ref).max() + floor


@pytest.mark.parametrize("name", [n for n in golden_names("spatial", world=1) if "bias" not in n and "legacy" not in n])
def test_module_from_columns_matches_reference_single_rank(name):
    ops = EmulatedOps(round_bf16=False)
    prev = losses._set_ops_for_testing(ops)
    try:
        meta, gold = load_golden(name)
        loss, gi, gt, ds = _run(meta, 0, 1)
        _assert_close(gold, 0, meta["gen"]["n"], loss, gi, gt, ds, meta["scale"])
        assert "forward_all_precomputed" in ops.calls and "build_positives" not in ops.calls
    finally:
        losses._set_ops_for_testing(prev)


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops = EmulatedOps(round_bf16=False)
        losses._set_ops_for_testing(ops)
        meta, _ = load_golden(name)
        res = _run(meta, rank, world)
        assert "build_positives" not in ops.calls
        q.put((rank,) + res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["spatial_n256_w2", "spatial_n256_w4_ll0_gwg1", "spatial_n192_w2_asym_text_ids"])
def test_module_from_columns_matches_reference_over_gloo(name):
    meta, gold = load_golden(name)
    world = meta["world"]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = meta["gen"]["n"] // world
    for rank, loss, gi, gt, ds in got:
        _assert_close(gold, rank, b, loss, gi, gt, ds, meta["scale"])


def test_collate_helper_and_argument_checks():
    b = make_spot_batch(n=12, d=64, k=4, seed=9, dup_frac=0.2)
    batch = dict(image_tile_ids=b.tile_ids, text_tile_ids=b.tile_ids, neighbor_tile_ids=b.neighbor_tile_ids,
                 neighbor_alphas=b.neighbor_alphas)
    out = collate_positive_columns(batch, 0.5)
    assert {"positive_columns", "positive_weights", "positive_probs"} <= set(out) and set(batch) <= set(out)
    assert torch.equal(out["positive_columns"][:, 0], torch.arange(12, dtype=torch.int32))
    np.testing.assert_allclose(out["positive_probs"].sum(1).numpy(), 1.0, rtol=1e-6)
    with pytest.raises(ValueError):
        resolve_positive_columns(b.tile_ids[:5], b.neighbor_tile_ids, b.neighbor_alphas)
    with pytest.raises(ValueError):
        SpatialLossFromColumns()(b.image_features, b.text_features, torch.tensor(10.0),
                                 positive_columns=out["positive_columns"][:5], positive_probs=out["positive_probs"])


def test_caller_supplied_lists_are_validated():
    """Lists from outside the library: out-of-range columns, a slot 0 that is not the row's own column
    (e.g. rank 0's columns used on another rank) and weight on unused slots raise instead of reaching the kernels."""
    b = make_spot_batch(n=16, d=64, k=4, seed=13)
    col, w, q = resolve_positive_columns(b.tile_ids, b.neighbor_tile_ids, b.neighbor_alphas, 0.5)
    prev = losses._set_ops_for_testing(EmulatedOps(round_bf16=False))
    try:
        def run(c, p, **kw):
            return SpatialLossFromColumns(**kw)(b.image_features, b.text_features, torch.tensor(10.0),
                                                positive_columns=c, positive_probs=p)
        run(col, q)  # the producer's own output is accepted
        bad = col.clone()
        bad[3, 1] = 16  # == N: one past the last column
        with pytest.raises(ValueError, match="outside"):
            run(bad, q)
        with pytest.raises(ValueError, match="own column"):
            run(col.roll(1, 0), q)  # lists of other rows (what a wrong rank offset produces)
        bad_q = q.clone()
        bad_q[col < 0] = 0.25
        with pytest.raises(ValueError, match="unused slot"):
            run(col, bad_q)
    finally:
        losses._set_ops_for_testing(prev)


def test_collate_requires_global_ids_when_distributed(monkeypatch):
    import spatial_clip_b200.positives as P

    b = make_spot_batch(n=8, d=64, k=2, seed=3)
    batch = dict(text_tile_ids=b.tile_ids, neighbor_tile_ids=b.neighbor_tile_ids, neighbor_alphas=b.neighbor_alphas)
    monkeypatch.setattr(P.dist, "is_initialized", lambda: True)
    monkeypatch.setattr(P.dist, "get_world_size", lambda: 2)
    with pytest.raises(ValueError, match="all_tile_ids"):
        collate_positive_columns(batch, 0.5)
    out = collate_positive_columns(batch, 0.5, all_tile_ids=torch.cat([b.tile_ids, b.tile_ids + 1000]), rank=1)
    assert torch.equal(out["positive_columns"][:, 0], torch.arange(8, 16, dtype=torch.int32))
