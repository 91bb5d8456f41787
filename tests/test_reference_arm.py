"""The reference arm of bench.py: oracle/_ref (bytecode compiled from the reference's own loss modules by
oracle/make_ref.py) loads, is the reference's code, and one emulated rank-step through it agrees with the oracle's
torch port and with the goldens minted from the reference.  CPU only; skips when oracle/_ref has not been built."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ref_loader
from oracle.torch_port import spatial_rank_step
from spatial_clip_b200.synth import make_spot_batch

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not built (python oracle/make_ref.py)")


def test_loaded_modules_are_the_reference_classes():
    oc_loss, comp, legacy = ref_loader.load_reference()
    assert comp.SpatialLoss.__module__ == "scl_reference_components_losses"
    assert comp.SpatialLoss.forward.__code__.co_filename.endswith("src/models/components/losses.py")
    assert hasattr(oc_loss, "gather_features") and hasattr(legacy, "GlobalMappingMultiPositiveClipLoss")
    import inspect

    names = list(inspect.signature(comp.SpatialLoss.forward).parameters)
    assert names[1:8] == ["image_features", "text_features", "logit_scale", "image_tile_ids", "text_tile_ids",
                          "neighbor_tile_ids", "neighbor_alphas"]


@pytest.mark.parametrize("world,rank", [(1, 0), (4, 2)])
def test_reference_rank_step_matches_port_and_goldens(world, rank):
    name = "spatial_n256_w4" if world == 4 else "spatial_n64_k8_default"
    meta, gold = load_golden(name)
    full = make_spot_batch(**meta["gen"])
    n = full.image_features.shape[0]
    b = n // world
    sl = slice(rank * b, (rank + 1) * b)
    c = meta["ctor"]
    ctor = dict(local_loss=c["local_loss"], gather_with_grad=c["gather_with_grad"],
                cap_logit_scale=c.get("cap_logit_scale"), temp_reg_weight=c.get("temp_reg_weight", 0.0),
                float32_logits=c.get("float32_logits", False), neighbor_alpha_scale=c.get("neighbor_alpha_scale", 1.0))

    def leaves():
        return (full.image_features[sl].clone().requires_grad_(True), full.text_features[sl].clone().requires_grad_(True),
                torch.tensor(float(meta["scale"]), requires_grad=True))

    img, txt, s = leaves()
    loss = ref_loader.reference_rank_step(img, txt, full.image_features, full.text_features, s, full.tile_ids,
                                          full.tile_ids[sl], full.neighbor_tile_ids[sl], full.neighbor_alphas[sl], rank,
                                          world, ctor)
    assert abs(float(loss) - gold["loss"][rank]) <= 1e-6 * abs(gold["loss"][rank]) + 1e-7
    assert abs(float(s.grad) - gold["d_scale"][rank]) <= 1e-5 * abs(gold["d_scale"][rank]) + 1e-7
    img2, txt2, s2 = leaves()
    loss2 = spatial_rank_step(img2, txt2, full.image_features, full.text_features, s2, full.tile_ids,
                              full.neighbor_tile_ids[sl], full.neighbor_alphas[sl], rank, cap=ctor["cap_logit_scale"],
                              temp_reg_weight=ctor["temp_reg_weight"], alpha_scale=ctor["neighbor_alpha_scale"])
    assert abs(float(loss) - float(loss2)) <= 2e-6 * abs(float(loss))
    if world == 1:  # the goldens hold d sum_r loss_r: at one rank that is this step's gradient
        np.testing.assert_allclose(img.grad.numpy(), gold["d_image"], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(txt.grad.numpy(), gold["d_text"], rtol=2e-4, atol=1e-7)
    else:
        assert img.grad.shape == (b, full.image_features.shape[1]) and torch.isfinite(img.grad).all()
