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


def test_ranks_under_graph_replay(ops):
    """From the third call on the module replays a captured graph; the ranks buffer is then the graph's static output
    ("of the last call") and must equal the eager module's on every step, with gradients flowing as usual."""
    from spatial_clip_b200 import SpatialLoss

    n, d, k = 1200, 256, 8
    graphed = SpatialLoss(track_retrieval_ranks=True, temp_reg_weight=0.05)
    eager = SpatialLoss(track_retrieval_ranks=True, temp_reg_weight=0.05, cuda_graphs=False)
    for step in range(5):
        b = make_spot_batch(n=n, d=d, k=k, seed=900 + step)
        noise = torch.nn.functional.normalize(torch.randn(n, d, generator=torch.Generator().manual_seed(step)), dim=-1)
        txt0 = torch.nn.functional.normalize(0.3 * b.image_features + noise, dim=-1)
        got = []
        for mod in (graphed, eager):
            img = b.image_features.cuda().requires_grad_(True)
            txt = txt0.cuda().requires_grad_(True)
            s = torch.tensor(20.0, device="cuda", requires_grad=True)
            ids = b.tile_ids.cuda()
            out = mod(img, txt, s, ids, ids, b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda())
            out["contrastive_loss"].backward()
            torch.cuda.synchronize()
            got.append((mod.last_retrieval_ranks.clone(), img.grad.clone()))
        assert torch.equal(got[0][0], got[1][0]), step
        assert torch.equal(got[0][1], got[1][1]), step
        assert got[0][0].shape == (n,) and int(got[0][0].max()) > 0
