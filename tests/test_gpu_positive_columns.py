"""GPU parity of the data-side soft-target route: SpatialLossFromColumns (phases = 6 of scl_fwd_all,
lists resolved by spatial_clip_b200.positives on the host) must give what SpatialLoss gives from the ids -- the same
kernels run in both and the sparse finish adds in a fixed order, so loss and gradients must be bit-identical -- and
the device builder's lists must equal the host producer's bit for bit.
"""
import pytest
import torch

from spatial_clip_b200.synth import make_spot_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,k", [(300, 256, 8), (5, 64, 8), (1024, 512, 6), (4096, 512, 8)])
def test_from_columns_equals_from_ids(n, d, k):
    from spatial_clip_b200 import SpatialLoss, SpatialLossFromColumns
    from spatial_clip_b200.positives import resolve_positive_columns

    b = make_spot_batch(n=n, d=d, k=k, seed=70 + n, dup_frac=0.02, self_loops=True)
    cfg = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
               neighbor_alpha_scale=0.5, float32_logits=True)
    ids = b.tile_ids.cuda()

    def run(from_columns):
        img = b.image_features.cuda().requires_grad_(True)
        txt = b.text_features.cuda().requires_grad_(True)
        s = torch.tensor(55.0, device="cuda", requires_grad=True)
        if from_columns:
            col, w, q = resolve_positive_columns(b.tile_ids, b.neighbor_tile_ids, b.neighbor_alphas, 0.5)
            mod = SpatialLossFromColumns(**cfg)
            out = mod(img, txt, s, positive_columns=col.cuda(), positive_probs=q.cuda(), positive_weights=w.cuda())
        else:
            mod = SpatialLoss(**cfg)
            out = mod(img, txt, s, ids, ids.clone(), b.neighbor_tile_ids.cuda(), b.neighbor_alphas.cuda())
        out["contrastive_loss"].backward()
        torch.cuda.synchronize()
        return out["contrastive_loss"].detach(), img.grad, txt.grad, s.grad, mod.last_positives

    l0, gi0, gt0, ds0, pos0 = run(False)
    l1, gi1, gt1, ds1, pos1 = run(True)
    # integer path: the device builder and the host producer agree bit for bit
    assert torch.equal(pos0[0].cpu(), pos1[0].cpu())
    assert torch.equal(pos0[1].cpu().view(torch.int32), pos1[1].cpu().view(torch.int32))
    assert torch.equal(pos0[2].cpu().view(torch.int32), pos1[2].cpu().view(torch.int32))
    assert torch.equal(l0, l1) and torch.equal(ds0, ds1)
    # gradients: identical kernels and a deterministic sparse finish (no atomics) -> bitwise equal
    assert torch.equal(gi0, gi1) and torch.equal(gt0, gt1)
