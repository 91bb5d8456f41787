"""GPU tests of the HBM-bound side passes through the C ABI (cast / normalise / caller-list validation)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from spatial_clip_b200._cuda import CudaOps

    return CudaOps()


@pytest.mark.parametrize("rows,d", [(300, 256), (5, 64), (4096, 512), (1000, 128), (32768, 512)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_cast_is_exact(ops, rows, d, dtype):
    g = torch.Generator().manual_seed(rows + d)
    x = torch.randn(rows, d, generator=g).cuda().to(dtype)
    y = ops.cast_bf16(x)
    torch.cuda.synchronize()
    assert torch.equal(y, x.float().to(torch.bfloat16))  # bit exact: same round-to-nearest-even as torch


def test_prepare_casts_both_modalities_and_caps_the_scale(ops):
    g = torch.Generator().manual_seed(3)
    a = torch.randn(777, 512, generator=g).cuda()
    b = torch.randn(777, 512, generator=g).cuda()
    img, txt, img_c, txt_c, scal = ops.prepare(a, b, torch.tensor([55.0], device="cuda"), 40.0)
    torch.cuda.synchronize()
    assert torch.equal(img, a.to(torch.bfloat16)) and torch.equal(txt, b.to(torch.bfloat16))
    assert img_c.data_ptr() == img.data_ptr() and txt_c.data_ptr() == txt.data_ptr()
    assert scal.tolist() == pytest.approx([40.0, 40.0 * 1.4426950408889634, 55.0], rel=1e-6)


def test_normalize_matches_f_normalize(ops):
    x = torch.randn(777, 512, generator=torch.Generator().manual_seed(1)).cuda()
    y = ops.cast_bf16(x, normalize=True)
    want = torch.nn.functional.normalize(x, dim=-1)
    # reference: open_clip model.py:326-345 normalises in fp32; we round the result to bf16
    assert (y.float() - want).abs().max().item() <= 2 ** -8 * want.abs().max().item() + 1e-6  # bf16 half-ulp
    norms = y.float().norm(dim=-1)
    assert (norms - 1).abs().max().item() < 5e-3


def test_check_positives_sanitises_and_flags(ops):
    col = torch.tensor([[0, 3, -1], [1, 99, 2], [2, -1, -7], [0, 1, -1]], dtype=torch.int32, device="cuda")
    q = torch.tensor([[0.5, 0.5, 0.25], [0.5, 0.25, 0.25], [1.0, 0.0, 0.3], [1.0, 0.0, 0.0]], device="cuda")
    c2, q2, flag = ops.check_positives(col, q, 4, 0)
    torch.cuda.synchronize()
    assert int(flag) == 1 | 2 | 4  # 99 / -7 out of range, row 3's slot 0 is not column 3, weight on an unused slot
    assert c2.tolist() == [[0, 3, -1], [1, -1, 2], [2, -1, -1], [0, 1, -1]]
    assert q2.tolist() == [[0.5, 0.5, 0.0], [0.5, 0.0, 0.25], [1.0, 0.0, 0.0], [1.0, 0.0, 0.0]]
    good = torch.tensor([[4, 1], [5, -1]], dtype=torch.int32, device="cuda")
    _, _, flag = ops.check_positives(good, torch.tensor([[0.5, 0.5], [1.0, 0.0]], device="cuda"), 8, 2)
    assert int(flag) == 0
