"""Pins the error budget of the fp32-accurate device mode (precision="fp32") on the CPU: the exact arithmetic the
kernels perform (bf16 hi/lo operand pairs, three products, fp32 accumulation, two-tile dL/dz) is restated in
``emulated_ops.SplitArithmeticOps`` and run through the real modules against the goldens minted from the
reference (fp32 inputs).  Gate = the one the GPU tests apply: loss rel 1e-5, grads 1e-4 of ||grad||_inf."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, text_ids_for
from emulated_ops import SplitArithmeticOps
from spatial_clip_b200 import ClipLoss, SpatialLoss, losses
from spatial_clip_b200.synth import make_spot_batch


@pytest.mark.parametrize("name", [n for n in golden_names(world=1) if "legacy" not in n])
def test_split_arithmetic_meets_the_fp32_gate(name):
    prev = losses._set_ops_for_testing(SplitArithmeticOps(round_bf16=False))
    try:
        meta, gold = load_golden(name)
        b = make_spot_batch(**meta["gen"])
        img = b.image_features.clone().requires_grad_(True)
        txt = b.text_features.clone().requires_grad_(True)
        s = torch.tensor(float(meta["scale"]), requires_grad=True)
        c = dict(meta["ctor"], precision="fp32")
        if meta["kind"] == "spatial":
            c.pop("cache_labels", None)
            out = SpatialLoss(**c)(img, txt, s, b.tile_ids, text_ids_for(meta, b), b.neighbor_tile_ids, b.neighbor_alphas)
        else:
            out = ClipLoss(**c)(img, txt, s)
        loss = out["contrastive_loss"]
        loss.backward()
    finally:
        losses._set_ops_for_testing(prev)
    scale, n = meta["scale"], meta["gen"]["n"]
    assert abs(float(loss.detach()) - gold["loss"][0]) <= 1e-5 * abs(gold["loss"][0]) + 2e-6 + 2e-7 * scale
    assert abs(float(s.grad) - gold["d_scale"][0]) <= 3e-4 * abs(gold["d_scale"][0]) + 2e-6
    floor = 3e-6 * scale * 0.5 / n
    for got, ref in ((img.grad.numpy(), gold["d_image"]), (txt.grad.numpy(), gold["d_text"])):
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max() + floor
