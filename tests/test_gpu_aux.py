"""GPU tests of the HBM-bound side passes through the C ABI (cast / normalise / transposed copy)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from spatial_clip_b200._cuda import CudaOps

    return CudaOps()


@pytest.mark.parametrize("rows,d", [(300, 256), (5, 64), (4096, 512), (1000, 128)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_cast_and_transposed_copy_are_exact(ops, rows, d, dtype):
    g = torch.Generator().manual_seed(rows + d)
    x = torch.randn(rows, d, generator=g).cuda().to(dtype)
    ld_t = (rows + 7) // 8 * 8
    y, y_t = ops.cast_bf16(x, want_rows=True, want_t=True, ld_t=ld_t)
    torch.cuda.synchronize()
    want = x.float().to(torch.bfloat16)
    assert torch.equal(y, want)  # bit exact: same round-to-nearest-even as torch
    assert torch.equal(y_t[:, :rows], want.t())
    assert (y_t[:, rows:] == 0).all()


def test_normalize_matches_f_normalize(ops):
    x = torch.randn(777, 512, generator=torch.Generator().manual_seed(1)).cuda()
    y, _ = ops.cast_bf16(x, normalize=True)
    want = torch.nn.functional.normalize(x, dim=-1)
    # reference: open_clip model.py:326-345 normalises in fp32; we round the result to bf16
    assert (y.float() - want).abs().max().item() <= 2 ** -8 * want.abs().max().item() + 1e-6  # bf16 half-ulp
    norms = y.float().norm(dim=-1)
    assert (norms - 1).abs().max().item() < 5e-3
