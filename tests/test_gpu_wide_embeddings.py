"""Embedding widths 640 / 1152 / 1280 / 1536 (the remaining open_clip model_configs widths): same kernels as
768 / 1024 with more D slices in the backward (640 -> 2 x 320 needs a 64-column accumulator group)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


@pytest.fixture(scope="module")
def ops():
    from spatial_clip_b200._cuda import CudaOps

    return CudaOps()


@pytest.mark.parametrize("m,n,d", [(300, 700, 640), (257, 1025, 1152), (130, 520, 1280), (300, 300, 1536)])
def test_wider_embeddings_kernels(ops, m, n, d):
    from dense_checker import row_stats
    from emulated_ops import EmulatedOps

    g = torch.Generator().manual_seed(m + n + d)
    x = torch.nn.functional.normalize(torch.randn(m, d, generator=g), dim=-1).cuda().bfloat16()
    y = torch.nn.functional.normalize(0.5 * torch.randn(n, d, generator=g) + x[:1].float().cpu(), dim=-1).cuda().bfloat16()
    s = 30.0
    scal = ops.prep_scalars(torch.tensor([s], device="cuda"), None)
    part, plan, z = ops.fwd_rowstats(x, y, scal, debug_z=True)
    assert (z - x.float() @ y.float().t()).abs().max().item() < 3e-5
    col = torch.full((m, 1), -1, dtype=torch.int32, device="cuda")
    q = torch.zeros((m, 1), device="cuda")
    stats = ops.row_finalize(part, plan, x, y, col, q).double()
    _, lse, mu, _ = row_stats(x, y, s)
    assert (stats[:, 0] * LN2 - lse).abs().max().item() < 2e-5 * s
    assert (stats[:, 1] - mu).abs().max().item() < 2e-5
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
