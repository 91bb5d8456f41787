"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libscl_b200.so via ctypes); the checker is the CPU oracle / goldens minted from the reference."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, text_ids_for
from oracle.contrastive_oracle import (clip_loss_oracle, dense_labels, soft_label_triples,
                                       spatial_loss_oracle)
from spatial_clip_b200.synth import make_spot_batch

pytestmark = pytest.mark.gpu

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


@pytest.fixture(scope="module")
def ops():
    """The CudaOps object the modules under test use too."""
    from spatial_clip_b200 import losses
    from spatial_clip_b200._cuda import CudaOps

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    o = CudaOps()
    prev = losses._set_ops_for_testing(o)
    yield o
    losses._set_ops_for_testing(prev)


def _bf16_pair(m, n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.nn.functional.normalize(torch.randn(m, d, generator=g), dim=-1)
    y = torch.nn.functional.normalize(0.5 * torch.randn(n, d, generator=g) + (x[:1] if m else 0), dim=-1)
    return x.cuda().to(torch.bfloat16), y.cuda().to(torch.bfloat16)


# ---------------------------------------------------------------- layer 1: TMA + UMMA + TMEM plumbing
@pytest.mark.parametrize("m,n,d", [(128, 256, 64), (128, 256, 512), (300, 300, 256), (1000, 2500, 512),
                                   (5, 5, 64), (257, 1025, 128)])
def test_similarity_tiles_match_matmul(ops, m, n, d):
    x, y = _bf16_pair(m, n, d, seed=m + n + d)
    scal = ops.prep_scalars(torch.tensor([20.0], device="cuda"), None)
    part, plan, z = ops.fwd_rowstats(x, y, scal, debug_z=True)
    torch.cuda.synchronize()
    ref = x.float() @ y.float().t()
    err = (z - ref).abs().max().item()
    assert err < 2e-5, f"max |z - x y^T| = {err}"


# ---------------------------------------------------------------- layer 2: online softmax statistics
@pytest.mark.parametrize("m,n,d,s", [(300, 300, 256, 14.2857), (1000, 2500, 512, 55.0), (64, 5000, 128, 100.0),
                                     (5, 5, 64, 14.2857)])
def test_row_statistics(ops, m, n, d, s):
    from dense_checker import row_stats

    x, y = _bf16_pair(m, n, d, seed=7 * m + n)
    scal = ops.prep_scalars(torch.tensor([s], device="cuda"), None)
    part, plan = ops.fwd_rowstats(x, y, scal)
    col = torch.full((m, 1), -1, dtype=torch.int32, device="cuda")
    q = torch.zeros((m, 1), device="cuda")
    stats = ops.row_finalize(part, plan, x, y, col, q).double()
    torch.cuda.synchronize()
    _, lse, mu, var = row_stats(x, y, s)
    assert (stats[:, 0] * LN2 - lse).abs().max().item() < 2e-5 * max(1.0, s)
    assert (stats[:, 1] - mu).abs().max().item() < 2e-5
    assert (stats[:, 2] - var).abs().max().item() < 2e-5
    assert stats[:, 3].abs().max().item() == 0.0


# ---------------------------------------------------------------- layer 3: integer path, bit exact
@pytest.mark.parametrize("name", [n for n in golden_names("spatial", world=1) if "bias" not in n and "legacy" not in n])
def test_positive_lists_bit_exact_vs_reference_labels(ops, name):
    meta, gold = load_golden(name)
    if "labels_i_t" not in gold:
        pytest.skip("no dense labels stored")
    b = make_spot_batch(**meta["gen"])
    n, k = b.neighbor_tile_ids.shape
    # image rows resolve their neighbours in the TEXT id map (losses.py:92,102-108)
    col, w, q = ops.build_positives(text_ids_for(meta, b).cuda(), b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda(),
                                    n, k, meta["ctor"].get("neighbor_alpha_scale", 1.0), 0, b.image_features.cuda())
    torch.cuda.synchronize()
    col, w, q = col.cpu().numpy(), w.cpu().numpy(), q.cpu().numpy()
    dense = np.zeros((n, n), dtype=np.float32)
    for i in range(n):
        seen = set()
        for t in range(k + 1):
            if col[i, t] >= 0:
                assert col[i, t] not in seen, "duplicate column not merged"
                seen.add(col[i, t])
                dense[i, col[i, t]] = w[i, t]
    assert np.array_equal(dense.view(np.uint32), gold["labels_i_t"].view(np.uint32))
    rs = dense.sum(1)
    np.testing.assert_allclose(q.sum(1), np.ones(n), rtol=1e-6)
    np.testing.assert_allclose((q * (col >= 0)).sum(1) * rs, rs, rtol=1e-6)


def test_positive_lists_multi_rank_offsets_and_sentinel_id(ops):
    b = make_spot_batch(n=96, d=64, k=8, seed=11, dup_frac=0.05, self_loops=True)
    ids = b.tile_ids.clone()
    ids[5] = torch.iinfo(torch.int64).min  # collides with the hash sentinel
    nbr = b.neighbor_tile_ids.clone()
    nbr[7, 0] = ids[5]
    alpha = b.neighbor_alphas.clone()
    alpha[7, 0] = 0.25
    world, bl = 3, 32
    for r in range(world):
        sl = slice(r * bl, (r + 1) * bl)
        col, w, _ = ops.build_positives(ids.cuda(), nbr[sl].cuda().contiguous(), alpha[sl].cuda().contiguous(), bl, 8,
                                        0.5, r, b.image_features.cuda())
        rows, _ = soft_label_triples(ids.numpy(), nbr[sl].numpy(), alpha[sl].numpy(), 0.5, r)
        want = dense_labels(rows, 96)
        got = np.zeros_like(want)
        col, w = col.cpu().numpy(), w.cpu().numpy()
        for i in range(bl):
            for t in range(9):
                if col[i, t] >= 0:
                    got[i, col[i, t]] = w[i, t]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


# ---------------------------------------------------------------- layer 4: fused backward pass alone
@pytest.mark.parametrize("m,n,d", [(128, 128, 64), (300, 300, 256), (300, 1000, 512), (5, 5, 64), (640, 2000, 384)])
def test_bwd_rows_matches_dense_formula(ops, m, n, d):
    from emulated_ops import EmulatedOps

    x, y = _bf16_pair(m, n, d, seed=3 * m + d)
    s = 30.0
    scal = ops.prep_scalars(torch.tensor([s], device="cuda"), None)
    g = torch.Generator().manual_seed(5)
    # arbitrary but plausible statistics: LSE-like offsets, small mu
    z = x.float() @ y.float().t()
    rs = torch.stack([torch.logsumexp(s * z, 1) * LOG2E, 0.1 * torch.rand(m, generator=g).cuda(),
                      torch.zeros(m).cuda(), torch.zeros(m).cuda()], 1).contiguous()
    cs = torch.stack([torch.logsumexp(s * z, 0) * LOG2E, 0.1 * torch.rand(n, generator=g).cuda(),
                      torch.zeros(n).cuda(), torch.zeros(n).cuda()], 1).contiguous()
    col = torch.full((m, 1), -1, dtype=torch.int32, device="cuda")
    q = torch.zeros((m, 1), device="cuda")
    ocol = torch.full((n, 1), -1, dtype=torch.int32, device="cuda")
    oq = torch.zeros((n, 1), device="cuda")
    gaps = torch.tensor([0.3], device="cuda")
    go = torch.tensor([1.7], device="cuda")
    args = (rs, cs, col, q, ocol, oq, max(m, n), 0, gaps, scal, go, 0.5 / m, 0.05, 1.0, 2, torch.float32)
    got = ops.bwd_rows(x, y, *args, opp_q_local=torch.zeros((m, 1), device="cuda"))
    torch.cuda.synchronize()
    cpu = [a.cpu() if torch.is_tensor(a) else a for a in args]
    want = EmulatedOps().bwd_rows(x.cpu(), y.cpu(), *cpu)
    err = (got.cpu() - want).abs().max().item()
    ref = want.abs().max().item()
    assert err <= 5e-3 * ref, f"bwd_rows err {err} vs max {ref}"  # G is rounded to bf16 for the 2nd GEMM


# ---------------------------------------------------------------- layer 5: the modules, vs the reference
def _module_run(meta, dtype=torch.float32):
    from spatial_clip_b200 import ClipLoss, SpatialLoss

    b = make_spot_batch(**meta["gen"])
    img = b.image_features.cuda().to(dtype).requires_grad_(True)
    txt = b.text_features.cuda().to(dtype).requires_grad_(True)
    s = torch.tensor(float(meta["scale"]), device="cuda", requires_grad=True)
    c = dict(meta["ctor"])
    if meta["kind"] == "legacy":  # positional order of open_clip_train, bare-tensor return
        from spatial_clip_b200 import GlobalMappingMultiPositiveClipLoss

        mod = GlobalMappingMultiPositiveClipLoss(**c)
        out = {"contrastive_loss": mod(img, txt, b.tile_ids.cuda(), text_ids_for(meta, b).cuda(),
                                       b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda(), s)}
    elif meta["kind"] == "spatial":
        mod = SpatialLoss(**c)
        out = mod(img, txt, s, b.tile_ids.cuda(), text_ids_for(meta, b).cuda(), b.neighbor_tile_ids.cuda(),
                  b.neighbor_alphas.cuda())
    else:
        mod = ClipLoss(**c)
        out = mod(img, txt, s)
    loss = out["contrastive_loss"]
    loss.backward()
    torch.cuda.synchronize()
    return b, float(loss.detach()), img.grad.float().cpu().numpy(), txt.grad.float().cpu().numpy(), float(s.grad)


def _oracle_on_bf16_inputs(meta, b):
    img = b.image_features.to(torch.bfloat16).float().numpy()
    txt = b.text_features.to(torch.bfloat16).float().numpy()
    c = meta["ctor"]
    if meta["kind"] in ("spatial", "legacy"):
        return spatial_loss_oracle(img, txt, meta["scale"], b.tile_ids.numpy(), text_ids_for(meta, b).numpy(),
                                   b.neighbor_tile_ids.numpy(), b.neighbor_alphas.numpy(), 1,
                                   c.get("cap_logit_scale"), c.get("temp_reg_weight", 0.0),
                                   c.get("neighbor_alpha_scale", 1.0))
    return clip_loss_oracle(img, txt, meta["scale"])


@pytest.mark.parametrize("name", golden_names(world=1))
def test_modules_match_reference(ops, name):
    """Gate (SURVEY 8d "parity gates", bf16 mode): against the oracle evaluated in fp64 on the SAME bf16-rounded
    inputs -- loss rel 2e-5 (+ the fp32-LSE floor), d_scale rel 1e-3, grads 5e-3 of ||grad||_inf (dL/dz is
    rounded to bf16 for the second GEMM; measured 1.4e-3 .. 2.5e-3).  Second check against the reference's fp32
    goldens (un-rounded fp32 inputs): loss rel 1e-3, grads 3e-2 of ||grad||_inf.  A saturated fixture (loss ~ 0,
    gradients at rounding level) cannot fail a relative gate meaningfully: it is checked in absolute terms instead
    (|grad| itself below the floor that the bf16 rounding of dL/dz allows)."""
    meta, gold = load_golden(name)
    if name.startswith("legacy"):  # same arithmetic through the open_clip_train positional signature
        meta = dict(meta, kind="legacy")
    b, loss, gi, gt, ds = _module_run(meta)
    scale = meta["scale"]
    orc = _oracle_on_bf16_inputs(meta, b)
    r0 = orc.ranks[0]
    report = {"loss": (loss, r0.loss, gold["loss"][0]), "ds": (ds, r0.d_scale, gold["d_scale"][0])}
    for nm, got, ref, gref in (("d_image", gi, orc.d_image, gold["d_image"]), ("d_text", gt, orc.d_text, gold["d_text"])):
        report[nm] = (np.abs(got - ref).max() / np.abs(ref).max(), np.abs(got - gref).max() / np.abs(gref).max())
    print(name, report)
    assert abs(loss - r0.loss) <= 2e-5 * abs(r0.loss) + 2e-6 * scale, report
    assert abs(ds - r0.d_scale) <= 1e-3 * abs(r0.d_scale) + 2e-6, report
    floor = 3e-6 * scale * 0.5 / len(orc.d_image)  # rounding level of dL/dz (bf16) x |Y| for one row
    for got, ref in ((gi, orc.d_image), (gt, orc.d_text)):
        if np.abs(ref).max() <= 4 * floor:  # saturated fixture (clip_n128_w1_s100): nothing relative to compare
            assert np.abs(got).max() <= 8 * floor, report
            continue
        assert np.abs(got - ref).max() <= 5e-3 * np.abs(ref).max() + floor, report
    assert abs(loss - gold["loss"][0]) <= 1e-3 * abs(gold["loss"][0]) + 2e-6 * scale, report
    for got, ref in ((gi, gold["d_image"]), (gt, gold["d_text"])):
        assert np.abs(got - ref).max() <= 3e-2 * np.abs(ref).max() + 3e-6 * scale * 0.5 / len(ref), report


def test_bf16_inputs_give_bf16_grads(ops):
    meta, gold = load_golden("spatial_n300_d256")
    from spatial_clip_b200 import SpatialLoss

    b = make_spot_batch(**meta["gen"])
    img = b.image_features.cuda().bfloat16().requires_grad_(True)
    txt = b.text_features.cuda().bfloat16().requires_grad_(True)
    s = torch.tensor(float(meta["scale"]), device="cuda", requires_grad=True)
    out = SpatialLoss(**meta["ctor"])(img, txt, s, b.tile_ids.cuda(), b.tile_ids.cuda(), b.neighbor_tile_ids.cuda(),
                                      b.neighbor_alphas.cuda())
    out["contrastive_loss"].backward()
    assert img.grad.dtype == torch.bfloat16 and txt.grad.dtype == torch.bfloat16
    ref = gold["d_image"]
    assert np.abs(img.grad.float().cpu().numpy() - ref).max() <= 2e-2 * np.abs(ref).max()


def test_fp16_inputs_give_fp16_grads(ops):
    """Half-precision features take the same path (cast to bf16 operands, fp32 statistics); the gradient comes back
    in the input dtype through bwd_gather_kernel<__half>."""
    meta, gold = load_golden("spatial_n300_d256")
    from spatial_clip_b200 import SpatialLoss

    b = make_spot_batch(**meta["gen"])
    img = b.image_features.cuda().half().requires_grad_(True)
    txt = b.text_features.cuda().half().requires_grad_(True)
    s = torch.tensor(float(meta["scale"]), device="cuda", requires_grad=True)
    out = SpatialLoss(**meta["ctor"])(img, txt, s, b.tile_ids.cuda(), b.tile_ids.cuda(), b.neighbor_tile_ids.cuda(),
                                      b.neighbor_alphas.cuda())
    out["contrastive_loss"].backward()
    assert img.grad.dtype == torch.float16 and txt.grad.dtype == torch.float16
    assert abs(float(out["contrastive_loss"].detach()) - gold["loss"][0]) <= 2e-3 * abs(gold["loss"][0])
    for got, ref in ((img.grad, gold["d_image"]), (txt.grad, gold["d_text"])):
        assert np.abs(got.float().cpu().numpy() - ref).max() <= 2e-2 * np.abs(ref).max()


def test_no_grad_forward_only(ops):
    meta, gold = load_golden("clip_n300_d256")
    from spatial_clip_b200 import ClipLoss

    b = make_spot_batch(**meta["gen"])
    with torch.no_grad():
        out = ClipLoss(**meta["ctor"])(b.image_features.cuda(), b.text_features.cuda(),
                                       torch.tensor(meta["scale"], device="cuda"))
    assert abs(float(out["contrastive_loss"]) - gold["loss"][0]) <= 1e-3 * gold["loss"][0]


# ---------------------------------------------------------------- larger sizes: dense torch checker + identities
@pytest.mark.parametrize("n,d,s", [(4096, 512, 14.2857), (8192, 512, 25.0)])
def test_clip_mid_size_vs_dense_checker(ops, n, d, s):
    from dense_checker import clip_loss_and_grads
    from spatial_clip_b200 import ClipLoss

    b = make_spot_batch(n=n, d=d, k=0, seed=2000 + n)
    img = b.image_features.cuda().bfloat16().float().requires_grad_(True)
    txt = b.text_features.cuda().bfloat16().float().requires_grad_(True)
    sc = torch.tensor(s, device="cuda", requires_grad=True)
    loss = ClipLoss()(img, txt, sc)["contrastive_loss"]
    loss.backward()
    want_loss, wi, wt, wds = clip_loss_and_grads(img.detach(), txt.detach(), s)
    assert abs(float(loss) - float(want_loss)) <= 2e-5 * float(want_loss) + 2e-6 * s
    assert abs(float(sc.grad) - float(wds)) <= 2e-3 * abs(float(wds)) + 1e-6
    for got, ref in ((img.grad, wi), (txt.grad, wt)):
        assert (got.double() - ref).abs().max().item() <= 5e-3 * ref.abs().max().item()
    # Euler identity: sum_i <x_i, dL/dx_i> = s * dL/ds for both modalities (size independent)
    e_i = (img.grad.double() * img.detach().double()).sum().item()
    e_t = (txt.grad.double() * txt.detach().double()).sum().item()
    assert abs(e_i - s * float(wds)) <= 5e-3 * abs(s * float(wds)) + 1e-5
    assert abs(e_t - s * float(wds)) <= 5e-3 * abs(s * float(wds)) + 1e-5


def test_full_size_identities_n32768(ops):
    """BASELINE size (N=32768, D=512, K=8): size-independent properties instead of a dense oracle."""
    from spatial_clip_b200 import SpatialLoss

    n, d, k = 32768, 512, 8
    b = make_spot_batch(n=n, d=d, k=k, seed=1004)
    img = b.image_features.cuda().requires_grad_(True)
    txt = b.text_features.cuda().requires_grad_(True)
    sc = torch.tensor(55.0, device="cuda", requires_grad=True)
    mod = SpatialLoss(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.0,
                      neighbor_alpha_scale=0.5, float32_logits=True)
    ids = b.tile_ids.cuda()
    loss = mod(img, txt, sc, ids, ids.clone(), b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda())["contrastive_loss"]
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss) and torch.isfinite(img.grad).all() and torch.isfinite(txt.grad).all()
    # sampled rows: exact LSE of 64 rows in fp64 -> per-row loss bound check through the Euler identity
    s_eff = 40.0
    e_i = (img.grad.double() * img.detach().bfloat16().double()).sum().item()
    e_t = (txt.grad.double() * txt.detach().bfloat16().double()).sum().item()
    want = s_eff * float(sc.grad)  # with temp_reg_weight == 0: sum <x, dx> = s_eff * dL/ds
    assert abs(e_i - want) <= 1e-2 * abs(want) + 1e-5
    assert abs(e_t - want) <= 1e-2 * abs(want) + 1e-5
    # positives: row weights normalised, own column first
    col, w, q = mod.last_positives
    assert (col[:, 0].cpu() == torch.arange(n, dtype=torch.int32)).all()
    assert (q.sum(1) - 1).abs().max().item() < 1e-5
    # a sampled row block against the dense checker
    from dense_checker import row_stats

    rows = torch.arange(0, n, n // 64, device="cuda")[:64]
    _, lse, _, _ = row_stats(img.detach()[rows].bfloat16(), txt.detach().bfloat16(), s_eff)
    zq = torch.zeros(64, dtype=torch.float64, device="cuda")
    xi = img.detach()[rows].bfloat16().double()
    tb = txt.detach().bfloat16().double()
    for t in range(k + 1):
        c = col[rows, t].long()
        zz = (xi * tb[c.clamp(min=0)]).sum(1)
        zq += torch.where(c >= 0, q[rows, t].double() * zz, torch.zeros_like(zz))
    per_row = lse - s_eff * zq  # image-direction loss terms of the sampled rows
    assert torch.isfinite(per_row).all() and (per_row > -1e-6).all()


# ---------------------------------------------------------------- BASELINE sizes: blockwise fp64 oracle
@pytest.mark.parametrize("n,same_ids", [(16384, True), (32768, True), (16384, False)])
def test_baseline_size_loss_and_sampled_gradients(ops, n, same_ids):
    """BASELINE configs[2] (N=16384, K=8) and the metric's batch (N=32768) with the shipped hyper-parameters
    (cap 40, temperature regulariser 0.05, alpha scale 0.5): loss, d logit_scale and 64 sampled rows of dImage / dGene
    against the blockwise fp64 oracle (oracle/blockwise_oracle.py) on the same bf16-rounded inputs.
    Gates: loss rel 2e-5, d_scale rel 1e-3, sampled gradient rows 5e-3 of the largest sampled gradient entry."""
    from oracle.blockwise_oracle import blockwise_oracle, sample_rows_for
    from spatial_clip_b200 import SpatialLoss
    from spatial_clip_b200.synth import shuffled_text_ids

    d, k = 512, 8
    b = make_spot_batch(n=n, d=d, k=k, seed=1004)
    tids = b.tile_ids if same_ids else shuffled_text_ids(b.tile_ids, 77)
    img = b.image_features.cuda().requires_grad_(True)
    txt = b.text_features.cuda().requires_grad_(True)
    sc = torch.tensor(55.0, device="cuda", requires_grad=True)
    cfg = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
               neighbor_alpha_scale=0.5, float32_logits=True)
    loss = SpatialLoss(**cfg)(img, txt, sc, b.tile_ids.cuda(), tids.cuda(), b.neighbor_tile_ids.cuda(),
                              b.neighbor_alphas.cuda())["contrastive_loss"]
    loss.backward()
    torch.cuda.synchronize()
    rows = sample_rows_for(n, 8, 8, seed=n)
    ref = blockwise_oracle(b.image_features.bfloat16().float().numpy(), b.text_features.bfloat16().float().numpy(),
                           55.0, b.tile_ids.numpy(), tids.numpy(), b.neighbor_tile_ids.numpy(),
                           b.neighbor_alphas.numpy(), 1, 40.0, 0.05, 0.5, True, True, rows)
    gi = img.grad[rows].double().cpu().numpy()
    gt = txt.grad[rows].double().cpu().numpy()
    rep = dict(loss=(float(loss), ref.loss[0]), ds=(float(sc.grad), ref.d_scale[0]),
               gi=np.abs(gi - ref.d_image_rows).max() / np.abs(ref.d_image_rows).max(),
               gt=np.abs(gt - ref.d_text_rows).max() / np.abs(ref.d_text_rows).max())
    print("baseline-size parity", n, same_ids, rep)
    assert abs(float(loss) - ref.loss[0]) <= 2e-5 * abs(ref.loss[0]), rep
    assert abs(float(sc.grad) - ref.d_scale[0]) <= 1e-3 * abs(ref.d_scale[0]), rep
    assert rep["gi"] <= 5e-3 and rep["gt"] <= 5e-3, rep


# ---------------------------------------------------------------- ranks sharing this GPU (gloo transport)
def _rank_worker(rank, world, port, name, q):
    import os
    import traceback

    import torch.distributed as dist

    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from spatial_clip_b200 import ClipLoss, SpatialLoss

        torch.cuda.set_device(0)
        meta, _ = load_golden(name)
        b = make_spot_batch(**meta["gen"]).rank_slice(rank, world)
        img = b.image_features.cuda().requires_grad_(True)
        txt = b.text_features.cuda().requires_grad_(True)
        s = torch.tensor(float(meta["scale"]), device="cuda", requires_grad=True)
        if meta["kind"] == "spatial":
            out = SpatialLoss(**meta["ctor"])(img, txt, s, b.tile_ids.cuda(), b.tile_ids.cuda(),
                                              b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda())
        else:
            out = ClipLoss(**dict(meta["ctor"]))(img, txt, s)
        out["contrastive_loss"].backward()
        torch.cuda.synchronize()
        q.put((rank, float(out["contrastive_loss"].detach()), img.grad.cpu().numpy(), txt.grad.cpu().numpy(),
               float(s.grad)))
    except Exception:  # surface the traceback in the parent instead of hanging the other ranks
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("name", ["spatial_n256_w2", "spatial_n256_w4_ll0_gwg0", "clip_n128_w2_ll0_gwg1",
                                  "clip_n128_w4_ll1_gwg0"])
def test_multi_rank_on_one_gpu(ops, name):
    """world_size 2/4, one process per rank, all on this GPU; collectives over gloo (NCCL needs one GPU per
    rank).  No kernel waits on another rank's kernel: the exchange is host-driven."""
    import socket

    import torch.multiprocessing as mp

    meta, gold = load_golden(name)
    world = meta["world"]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
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
        rep = (name, rank, loss, gold["loss"][rank], ds, gold["d_scale"][rank],
               np.abs(gi - gold["d_image"][sl]).max() / np.abs(gold["d_image"][sl]).max(),
               np.abs(gt - gold["d_text"][sl]).max() / np.abs(gold["d_text"][sl]).max())
        print(rep)
        # vs the reference's fp32 goldens (inputs NOT pre-rounded to bf16): bf16-mode tolerances
        assert abs(loss - gold["loss"][rank]) <= 1e-3 * abs(gold["loss"][rank]) + 2e-6 * meta["scale"], rep
        assert abs(ds - gold["d_scale"][rank]) <= 3e-2 * abs(gold["d_scale"][rank]) + 1e-5, rep
        assert rep[6] <= 3e-2 and rep[7] <= 3e-2, rep


# ---------------------------------------------------------------- wide embeddings (ViT-L / ViT-H: D = 768, 1024)
@pytest.mark.parametrize("m,n,d", [(300, 700, 768), (257, 1025, 1024)])
def test_wide_embeddings_kernels(ops, m, n, d):
    """D > 512: streamed-X forward, two-slice backward."""
    from dense_checker import row_stats
    from emulated_ops import EmulatedOps

    x, y = _bf16_pair(m, n, d, seed=m + n + d)
    s = 30.0
    scal = ops.prep_scalars(torch.tensor([s], device="cuda"), None)
    part, plan, z = ops.fwd_rowstats(x, y, scal, debug_z=True)
    assert (z - x.float() @ y.float().t()).abs().max().item() < 3e-5
    col = torch.full((m, 1), -1, dtype=torch.int32, device="cuda")
    q = torch.zeros((m, 1), device="cuda")
    stats = ops.row_finalize(part, plan, x, y, col, q).double()
    _, lse, mu, var = row_stats(x, y, s)
    assert (stats[:, 0] * LN2 - lse).abs().max().item() < 2e-5 * s
    assert (stats[:, 1] - mu).abs().max().item() < 2e-5
    # backward pass alone
    zz = x.float() @ y.float().t()
    rs = torch.stack([torch.logsumexp(s * zz, 1) * LOG2E, 0.1 * torch.rand(m).cuda(), torch.zeros(m).cuda(),
                      torch.zeros(m).cuda()], 1).contiguous()
    cs = torch.stack([torch.logsumexp(s * zz, 0) * LOG2E, 0.1 * torch.rand(n).cuda(), torch.zeros(n).cuda(),
                      torch.zeros(n).cuda()], 1).contiguous()
    ocol = torch.full((n, 1), -1, dtype=torch.int32, device="cuda")
    oq = torch.zeros((n, 1), device="cuda")
    gaps = torch.tensor([0.2], device="cuda")
    go = torch.tensor([1.3], device="cuda")
    args = (rs, cs, col, q, ocol, oq, max(m, n), 0, gaps, scal, go, 0.5 / m, 0.05, 1.0, 2, torch.float32)
    got = ops.bwd_rows(x, y, *args, opp_q_local=torch.zeros((m, 1), device="cuda"))
    torch.cuda.synchronize()
    want = EmulatedOps().bwd_rows(x.cpu(), y.cpu(), *[a.cpu() if torch.is_tensor(a) else a for a in args])
    assert (got.cpu() - want).abs().max().item() <= 5e-3 * want.abs().max().item()


@pytest.mark.parametrize("d", [640, 768, 1024, 1280])
def test_wide_embeddings_module_vs_oracle(ops, d):
    from spatial_clip_b200 import SpatialLoss

    gen = dict(n=300, d=d, k=8, seed=4000 + d, dup_frac=0.01, self_loops=True)
    b = make_spot_batch(**gen)
    cfg = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
               neighbor_alpha_scale=0.5, float32_logits=True)
    img = b.image_features.cuda().requires_grad_(True)
    txt = b.text_features.cuda().requires_grad_(True)
    s = torch.tensor(55.0, device="cuda", requires_grad=True)
    ids = b.tile_ids.cuda()
    loss = SpatialLoss(**cfg)(img, txt, s, ids, ids, b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda())[
        "contrastive_loss"]
    loss.backward()
    torch.cuda.synchronize()
    orc = spatial_loss_oracle(b.image_features.bfloat16().float().numpy(), b.text_features.bfloat16().float().numpy(),
                              55.0, b.tile_ids.numpy(), b.tile_ids.numpy(), b.neighbor_tile_ids.numpy(),
                              b.neighbor_alphas.numpy(), 1, 40.0, 0.05, 0.5)
    r0 = orc.ranks[0]
    assert abs(float(loss.detach()) - r0.loss) <= 2e-5 * abs(r0.loss) + 1e-4
    assert abs(float(s.grad) - r0.d_scale) <= 1e-3 * abs(r0.d_scale) + 2e-6
    for got, ref in ((img.grad.cpu().numpy(), orc.d_image), (txt.grad.cpu().numpy(), orc.d_text)):
        assert np.abs(got - ref).max() <= 5e-3 * np.abs(ref).max()


# ---------------------------------------------------------------- run-to-run determinism
def test_backward_is_bitwise_reproducible(ops):
    """No atomics anywhere on the gradient path: chunk partials are summed in chunk order and the opposite-direction
    soft-target terms in ascending entry order (bwd_gather), so two runs give bit-identical results -- also with
    duplicate tile ids and hub rows that many rows name as a neighbour."""
    from spatial_clip_b200 import SpatialLoss

    b = make_spot_batch(n=4096, d=256, k=8, seed=99, dup_frac=0.05, self_loops=True)
    nbr = b.neighbor_tile_ids.clone()
    alpha = b.neighbor_alphas.clone()
    nbr[::3, 0] = b.tile_ids[7]  # a hub: ~1365 rows list row 7 (bucket > 32 entries: the selection path)
    alpha[::3, 0] = 0.2
    cfg = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
               neighbor_alpha_scale=0.5, float32_logits=True)

    def run():
        img = b.image_features.cuda().requires_grad_(True)
        txt = b.text_features.cuda().requires_grad_(True)
        s = torch.tensor(55.0, device="cuda", requires_grad=True)
        ids = b.tile_ids.cuda()
        loss = SpatialLoss(**cfg)(img, txt, s, ids, ids.clone(), nbr.cuda(), alpha.cuda())["contrastive_loss"]
        loss.backward()
        torch.cuda.synchronize()
        return loss.detach().clone(), img.grad.clone(), txt.grad.clone(), s.grad.clone()

    first = run()
    for _ in range(2):
        again = run()
        for a, c in zip(first, again):
            assert torch.equal(a, c)
    # and the hub row's gradient is right (dense fp64 oracle on the bf16-rounded inputs)
    orc = spatial_loss_oracle(b.image_features.bfloat16().float().numpy(), b.text_features.bfloat16().float().numpy(),
                              55.0, b.tile_ids.numpy(), b.tile_ids.numpy(), nbr.numpy(), alpha.numpy(), 1, 40.0, 0.05, 0.5)
    gi = first[1].cpu().numpy()
    assert np.abs(gi - orc.d_image).max() <= 5e-3 * np.abs(orc.d_image).max()
    assert np.abs(gi[7] - orc.d_image[7]).max() <= 5e-3 * np.abs(orc.d_image[7]).max()
    gt = first[2].cpu().numpy()
    assert np.abs(gt - orc.d_text).max() <= 5e-3 * np.abs(orc.d_text).max()


# ---------------------------------------------------------------- CUDA-graph replay of the step
@pytest.mark.parametrize("kind", ["spatial", "spatial_same_ids", "clip"])
def test_graph_replay_is_bitwise_identical_to_eager(ops, kind):
    """After two eager calls per configuration the modules replay forward and backward as CUDA graphs: same kernels
    on the same values, so every step must equal the eager module bit for bit -- with inputs that change from step to
    step and live at different addresses."""
    from spatial_clip_b200 import ClipLoss, SpatialLoss, release_cuda_graphs

    release_cuda_graphs()  # independent of what earlier tests of this process captured (the cache is bounded)
    n, d, k = 1500, 256, 8
    cfg = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
               neighbor_alpha_scale=0.5, float32_logits=True)
    if kind == "clip":
        graphed, eager = ClipLoss(), ClipLoss(cuda_graphs=False)
    else:
        graphed, eager = SpatialLoss(**cfg), SpatialLoss(**cfg, cuda_graphs=False)
    keep = []
    for step in range(6):
        b = make_spot_batch(n=n, d=d, k=k, seed=300 + step, dup_frac=0.02, self_loops=True)
        res = []
        for mod in (graphed, eager):
            img = b.image_features.cuda().requires_grad_(True)
            txt = b.text_features.cuda().requires_grad_(True)
            s = torch.tensor(30.0 + step, device="cuda", requires_grad=True)
            keep.append((img, txt))  # keep earlier inputs alive so that later ones get fresh addresses
            if kind == "clip":
                out = mod(img, txt, s)
            else:
                ids = b.tile_ids.cuda()
                tids = ids if kind == "spatial_same_ids" else ids.clone()
                out = mod(img, txt, s, ids, tids, b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda())
            loss = out["contrastive_loss"]
            (loss * (1.0 + 0.5 * step)).backward()  # a different upstream gradient every step
            torch.cuda.synchronize()
            res.append((loss.detach().clone(), img.grad.clone(), txt.grad.clone(), s.grad.clone()))
        for a, c in zip(*res):
            assert torch.equal(a, c), (kind, step)
    from spatial_clip_b200 import losses

    assert any(st.fwd is not None and st.bwd is not None for st in losses._GRAPHS.values()), "no graph was captured"


def test_graph_buffers_guard_against_out_of_order_backward(ops):
    from spatial_clip_b200 import ClipLoss, release_cuda_graphs

    release_cuda_graphs()
    mod = ClipLoss()
    b = make_spot_batch(n=512, d=128, k=0, seed=5)

    def fwd():
        img = b.image_features.cuda().requires_grad_(True)
        txt = b.text_features.cuda().requires_grad_(True)
        return mod(img, txt, torch.tensor(20.0, device="cuda", requires_grad=True))["contrastive_loss"]

    for _ in range(3):
        fwd().backward()
    first = fwd()
    second = fwd()  # overwrites the graph's saved activations
    with pytest.raises(RuntimeError, match="overwritten"):
        first.backward()
    second.backward()


def test_no_grad_forward_replays_too(ops):
    from spatial_clip_b200 import ClipLoss

    mod, ref = ClipLoss(), ClipLoss(cuda_graphs=False)
    for step in range(5):
        b = make_spot_batch(n=700, d=128, k=0, seed=40 + step)
        with torch.no_grad():
            s = torch.tensor(25.0, device="cuda")
            a = mod(b.image_features.cuda(), b.text_features.cuda(), s)["contrastive_loss"]
            c = ref(b.image_features.cuda(), b.text_features.cuda(), s)["contrastive_loss"]
        assert torch.equal(a, c)


# ---------------------------------------------------------------- more neighbour slots than the default 8
@pytest.mark.parametrize("k", [12, 20, 31])
def test_positive_lists_with_many_neighbour_slots(ops, k):
    """K > 8 runs the 16- / 32-lane builder groups: the lists must equal the oracle's dense labels and the host
    producer's lists bit for bit, and the module must still match the oracle."""
    from spatial_clip_b200 import SpatialLoss
    from spatial_clip_b200.positives import resolve_positive_columns

    n, d = 384, 128
    b = make_spot_batch(n=n, d=d, k=k, seed=700 + k, dup_frac=0.05, self_loops=True, negative_alphas=True)
    col, w, q = ops.build_positives(b.tile_ids.cuda(), b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda(), n, k, 0.5, 0,
                                    b.image_features.cuda())
    torch.cuda.synchronize()
    rows, _ = soft_label_triples(b.tile_ids.numpy(), b.neighbor_tile_ids.numpy(), b.neighbor_alphas.numpy(), 0.5, 0)
    want = dense_labels(rows, n)
    got = np.zeros_like(want)
    c_np, w_np = col.cpu().numpy(), w.cpu().numpy()
    for i in range(n):
        for t in range(k + 1):
            if c_np[i, t] >= 0:
                got[i, c_np[i, t]] = w_np[i, t]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    hc, hw, hq = resolve_positive_columns(b.tile_ids, b.neighbor_tile_ids, b.neighbor_alphas, 0.5)
    assert torch.equal(col.cpu(), hc) and torch.equal(w.cpu().view(torch.int32), hw.view(torch.int32))
    assert torch.equal(q.cpu().view(torch.int32), hq.view(torch.int32))
    img = b.image_features.cuda().requires_grad_(True)
    txt = b.text_features.cuda().requires_grad_(True)
    s = torch.tensor(30.0, device="cuda", requires_grad=True)
    ids = b.tile_ids.cuda()
    loss = SpatialLoss(cap_logit_scale=40.0, temp_reg_weight=0.05, neighbor_alpha_scale=0.5)(
        img, txt, s, ids, ids.clone(), b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda())["contrastive_loss"]
    loss.backward()
    orc = spatial_loss_oracle(b.image_features.bfloat16().float().numpy(), b.text_features.bfloat16().float().numpy(), 30.0,
                              b.tile_ids.numpy(), b.tile_ids.numpy(), b.neighbor_tile_ids.numpy(),
                              b.neighbor_alphas.numpy(), 1, 40.0, 0.05, 0.5)
    assert abs(float(loss.detach()) - orc.ranks[0].loss) <= 2e-5 * abs(orc.ranks[0].loss) + 6e-5
    for got_g, ref in ((img.grad.cpu().numpy(), orc.d_image), (txt.grad.cpu().numpy(), orc.d_text)):
        assert np.abs(got_g - ref).max() <= 5e-3 * np.abs(ref).max()
