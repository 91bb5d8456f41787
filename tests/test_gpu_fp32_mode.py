"""GPU parity of the fp32-accurate mode (``precision="fp32"``: bf16 hi/lo operand pairs, three tensor-core products
per GEMM, dL/dz as two bf16 tiles) against the REFERENCE's fp32 goldens on the same (unrounded) fp32 inputs.

Stated tolerance of the mode (measured on a B200, round 2): gradients 1e-4 of ||grad||_inf, d logit_scale rel 3e-4,
loss |err| <= 1e-5 |loss| + 1e-6 s_eff.  The second term is the tensor cores' fp32 accumulation: it truncates instead
of rounding, which biases every similarity by ~ -5e-7 |z| (measured: similarities within 3e-6 of fp64, loss low by
4.6e-7 s on the saturated fixtures) -- two orders below the bf16 mode's 1e-3, one above an IEEE fp32 GEMM."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, text_ids_for
from spatial_clip_b200.synth import make_spot_batch

pytestmark = pytest.mark.gpu

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


@pytest.fixture(scope="module")
def ops():
    from spatial_clip_b200 import losses
    from spatial_clip_b200._cuda import CudaOps

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    o = CudaOps()
    prev = losses._set_ops_for_testing(o)
    yield o
    losses._set_ops_for_testing(prev)


def _split(x):
    h = x.float().bfloat16()
    l = (x.float() - h.float()).bfloat16()
    return h, l


def _fp32_pair(m, n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.nn.functional.normalize(torch.randn(m, d, generator=g), dim=-1)
    y = torch.nn.functional.normalize(0.5 * torch.randn(n, d, generator=g) + x[:1], dim=-1)
    return x.cuda(), y.cuda()


@pytest.mark.parametrize("rows,d,dtype", [(300, 256, torch.float32), (5, 64, torch.float32), (1000, 512, torch.float16),
                                          (257, 128, torch.bfloat16)])
def test_split_operands_are_exact(ops, rows, d, dtype):
    x = torch.randn(rows, d, generator=torch.Generator().manual_seed(rows + d)).to(dtype).cuda()
    r, c = ops.split_cast(x)
    torch.cuda.synchronize()
    h, l = _split(x)
    assert torch.equal(r, torch.cat([h, h, l], 1)) and torch.equal(c, torch.cat([h, l, h], 1))
    assert (x.float() - h.float() - l.float()).abs().max().item() <= 2.0 ** -17 * x.float().abs().max().item()


@pytest.mark.parametrize("m,n,d", [(128, 256, 64), (300, 300, 256), (1000, 2500, 512), (5, 5, 64), (257, 1025, 1024)])
def test_split_similarity_and_row_statistics(ops, m, n, d):
    from dense_checker import row_stats

    x, y = _fp32_pair(m, n, d, seed=m + n + d)
    xr, _ = ops.split_cast(x, want_cols=False)
    _, yc = ops.split_cast(y, want_rows=False)
    s = 40.0
    scal = ops.prep_scalars(torch.tensor([s], device="cuda"), None)
    part, plan, z = ops.fwd_rowstats(xr, yc, scal, debug_z=True)
    torch.cuda.synchronize()
    ref = x.double() @ y.double().t()
    assert (z.double() - ref).abs().max().item() < 4e-6, "three-product similarity must be fp32-grade"
    col = torch.full((m, 1), -1, dtype=torch.int32, device="cuda")
    q = torch.zeros((m, 1), device="cuda")
    stats = ops.row_finalize(part, plan, xr, yc, col, q).double()
    lse = torch.logsumexp(s * ref, 1)
    assert (stats[:, 0] * LN2 - lse).abs().max().item() < 1e-4
    p = torch.softmax(s * ref, 1)
    assert (stats[:, 1] - (p * ref).sum(1)).abs().max().item() < 5e-6


@pytest.mark.parametrize("m,n,d", [(128, 128, 64), (300, 1000, 512), (5, 5, 64), (640, 2000, 384), (300, 700, 768)])
def test_split_bwd_rows_matches_dense_formula(ops, m, n, d):
    from emulated_ops import EmulatedOps

    x, y = _fp32_pair(m, n, d, seed=3 * m + d)
    xr, _ = ops.split_cast(x, want_cols=False)
    _, yc = ops.split_cast(y, want_rows=False)
    s = 30.0
    scal = ops.prep_scalars(torch.tensor([s], device="cuda"), None)
    g = torch.Generator().manual_seed(5)
    z = x.double() @ y.double().t()
    rs = torch.stack([(torch.logsumexp(s * z, 1) * LOG2E).float(), 0.1 * torch.rand(m, generator=g).cuda(),
                      torch.zeros(m).cuda(), torch.zeros(m).cuda()], 1).contiguous()
    cs = torch.stack([(torch.logsumexp(s * z, 0) * LOG2E).float(), 0.1 * torch.rand(n, generator=g).cuda(),
                      torch.zeros(n).cuda(), torch.zeros(n).cuda()], 1).contiguous()
    col = torch.full((m, 1), -1, dtype=torch.int32, device="cuda")
    q = torch.zeros((m, 1), device="cuda")
    ocol = torch.full((n, 1), -1, dtype=torch.int32, device="cuda")
    oq = torch.zeros((n, 1), device="cuda")
    gaps = torch.tensor([0.3], device="cuda")
    go = torch.tensor([1.7], device="cuda")
    args = (rs, cs, col, q, ocol, oq, max(m, n), 0, gaps, scal, go, 0.5 / m, 0.05, 1.0, 2, torch.float32)
    got = ops.bwd_rows(xr, yc, *args, opp_q_local=torch.zeros((m, 1), device="cuda"), split=True)
    torch.cuda.synchronize()
    cpu = [a.cpu() if torch.is_tensor(a) else a for a in args]
    want = EmulatedOps(round_bf16=False).bwd_rows(x.cpu(), y.cpu(), *cpu)
    err = (got.cpu() - want).abs().max().item()
    ref = want.abs().max().item()
    assert err <= 1e-4 * ref, f"split bwd_rows err {err} vs max {ref}"


def _run(meta, **extra):
    from spatial_clip_b200 import ClipLoss, SpatialLoss

    b = make_spot_batch(**meta["gen"])
    img = b.image_features.cuda().requires_grad_(True)
    txt = b.text_features.cuda().requires_grad_(True)
    s = torch.tensor(float(meta["scale"]), device="cuda", requires_grad=True)
    c = dict(meta["ctor"], precision="fp32", **extra)
    if meta["kind"] == "spatial":
        c.pop("cache_labels", None)
        out = SpatialLoss(**c)(img, txt, s, b.tile_ids.cuda(), text_ids_for(meta, b).cuda(), b.neighbor_tile_ids.cuda(),
                               b.neighbor_alphas.cuda())
    else:
        out = ClipLoss(**c)(img, txt, s)
    loss = out["contrastive_loss"]
    loss.backward()
    torch.cuda.synchronize()
    return float(loss.detach()), img.grad.cpu().numpy(), txt.grad.cpu().numpy(), float(s.grad)


@pytest.mark.parametrize("name", [n for n in golden_names(world=1) if "legacy" not in n])
def test_fp32_mode_modules_match_reference(ops, name):
    """vs the reference's own fp32 outputs on the same fp32 inputs: loss 1e-5 |loss| + 1e-6 s_eff (module docstring),
    d_scale rel 3e-4, grads 1e-4 of ||grad||_inf."""
    meta, gold = load_golden(name)
    loss, gi, gt, ds = _run(meta)
    scale, n = meta["scale"], meta["gen"]["n"]
    rep = (loss, gold["loss"][0], ds, gold["d_scale"][0], np.abs(gi - gold["d_image"]).max() / np.abs(gold["d_image"]).max(),
           np.abs(gt - gold["d_text"]).max() / np.abs(gold["d_text"]).max())
    print(name, rep)
    s_eff = min(scale, meta["ctor"].get("cap_logit_scale") or scale)
    assert abs(loss - gold["loss"][0]) <= 1e-5 * abs(gold["loss"][0]) + 2e-6 + 1e-6 * s_eff, rep
    assert abs(ds - gold["d_scale"][0]) <= 3e-4 * abs(gold["d_scale"][0]) + 2e-6, rep
    floor = 3e-6 * scale * 0.5 / n
    for got, ref in ((gi, gold["d_image"]), (gt, gold["d_text"])):
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max() + floor, rep


def test_fp32_mode_mid_size_vs_dense_checker(ops):
    from dense_checker import clip_loss_and_grads
    from spatial_clip_b200 import ClipLoss

    n, d, s = 4096, 512, 14.2857
    b = make_spot_batch(n=n, d=d, k=0, seed=2000 + n)
    img = b.image_features.cuda().requires_grad_(True)
    txt = b.text_features.cuda().requires_grad_(True)
    sc = torch.tensor(s, device="cuda", requires_grad=True)
    loss = ClipLoss(precision="fp32")(img, txt, sc)["contrastive_loss"]
    loss.backward()
    want_loss, wi, wt, wds = clip_loss_and_grads(img.detach(), txt.detach(), s)
    assert abs(float(loss) - float(want_loss)) <= 1e-5 * float(want_loss) + 2e-6 + 1e-6 * s
    for got, ref in ((img.grad, wi), (txt.grad, wt)):
        assert (got.double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


def _rank_worker(rank, world, port, name, q):
    import traceback

    import torch.distributed as dist

    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from spatial_clip_b200 import SpatialLoss

        torch.cuda.set_device(0)
        meta, _ = load_golden(name)
        b = make_spot_batch(**meta["gen"]).rank_slice(rank, world)
        img = b.image_features.cuda().requires_grad_(True)
        txt = b.text_features.cuda().requires_grad_(True)
        s = torch.tensor(float(meta["scale"]), device="cuda", requires_grad=True)
        out = SpatialLoss(**meta["ctor"], precision="fp32")(img, txt, s, b.tile_ids.cuda(), b.tile_ids.cuda(),
                                                             b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda())
        out["contrastive_loss"].backward()
        torch.cuda.synchronize()
        q.put((rank, float(out["contrastive_loss"].detach()), img.grad.cpu().numpy(), txt.grad.cpu().numpy(),
               float(s.grad)))
    except Exception:
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("name", ["spatial_n256_w2", "spatial_n256_w4_ll0_gwg0"])
def test_fp32_mode_multi_rank_on_one_gpu(ops, name):
    """The (h|l|h) column operands travel through the all-gather; statistics exchange as in the bf16 mode."""
    import socket

    import torch.multiprocessing as mp

    meta, gold = load_golden(name)
    world = meta["world"]
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, world, port, name, q), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    got = []
    try:
        for _ in range(world):
            item = q.get(timeout=150)
            assert len(item) == 5, f"rank {item[0]} raised:\n{item[1]}"
            got.append(item)
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    bl = meta["gen"]["n"] // world
    for rank, loss, gi, gt, ds in got:
        sl = slice(rank * bl, (rank + 1) * bl)
        s_eff = min(meta["scale"], meta["ctor"].get("cap_logit_scale") or meta["scale"])
        assert abs(loss - gold["loss"][rank]) <= 1e-5 * abs(gold["loss"][rank]) + 2e-6 + 1e-6 * s_eff
        assert abs(ds - gold["d_scale"][rank]) <= 3e-4 * abs(gold["d_scale"][rank]) + 2e-6
        floor = 3e-6 * meta["scale"] * 0.5 / bl
        for got_g, ref in ((gi, gold["d_image"][sl]), (gt, gold["d_text"][sl])):
            assert np.abs(got_g - ref).max() <= 1e-4 * np.abs(ref).max() + floor
