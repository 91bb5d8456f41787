"""GPU parity of the in-pass retrieval ranks: the counts of fwd_rowstats_pair_kernel<1> against the
reference metric's logits-matmul + topk on the same bf16-rounded features.
"""
import pytest
import torch

from spatial_clip_b200.synth import make_spot_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from spatial_clip_b200 import losses
    from spatial_clip_b200._cuda import CudaOps

    o = CudaOps()
    prev = losses._set_ops_for_testing(o)
    yield o
    losses._set_ops_for_testing(prev)


@pytest.mark.parametrize("n,d,k_nbr", [(300, 256, 8), (5, 64, 8), (1000, 512, 0), (4096, 512, 8)])
def test_ranks_match_dense_count(ops, n, d, k_nbr):
    from spatial_clip_b200 import ClipLoss, SpatialLoss
    from spatial_clip_b200.metrics import recall_at_k

    b = make_spot_batch(n=n, d=d, k=k_nbr, seed=50 + n)
    noise = torch.nn.functional.normalize(torch.randn(n, d, generator=torch.Generator().manual_seed(n)), dim=-1)
    img = b.image_features.cuda()
    txt = torch.nn.functional.normalize(0.3 * b.image_features + noise, dim=-1).cuda()
    s = torch.tensor(20.0, device="cuda")
    with torch.no_grad():
        if k_nbr:
            mod = SpatialLoss(track_retrieval_ranks=True, temp_reg_weight=0.05)
            ids = b.tile_ids.cuda()
            mod(img, txt, s, ids, ids, b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda())
        else:
            mod = ClipLoss(track_retrieval_ranks=True)
            mod(img, txt, s)
    ranks = mod.last_retrieval_ranks
    torch.cuda.synchronize()
    z = img.bfloat16().double() @ txt.bfloat16().double().t()
    diag = z.diagonal()[:, None]
    want = ((z > diag).sum(1) - 0).to(torch.int32)  # the diagonal itself is never > itself
    # fp32 accumulation order differs between the tensor-core pass and the own-pair dot product: allow near-ties
    near = ((z - diag).abs() < 2e-6).sum(1) - 1
    assert ((ranks - want).abs() <= near.to(torch.int32)).all(), (ranks - want).abs().max().item()
    for k in (1, 5, 10):
        got = recall_at_k(ranks, k).item()
        ref = (want < min(k, n)).float().mean().item()
        assert abs(got - ref) <= 2.0 / n
