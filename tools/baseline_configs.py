#!/usr/bin/env python
"""The two single-GPU BASELINE.json configs that bench.py (the metric's configuration) does not time:
configs[1] symmetric InfoNCE (ClipLoss) at B = 4096, D = 512, bf16, and configs[2] the multi-positive SpatialLoss at
B = 16384, D = 512, K = 8 -- ms per fwd+bwd step (CUDA events, median) and parity against the CPU oracle.
Developer tool (needs a B200):  python tools/baseline_configs.py > profiles/r2_configs.jsonl
"""
import json
import statistics
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.blockwise_oracle import blockwise_oracle, sample_rows_for  # noqa: E402  (checker only)
from spatial_clip_b200 import ClipLoss, SpatialLoss  # noqa: E402
from spatial_clip_b200.synth import make_spot_batch  # noqa: E402

PEAK = 1618.8
p = ROOT / "MEASURED_PEAKS.json"
if p.exists():
    PEAK = float(json.loads(p.read_text())["bf16_tflops"])


def run(kind, n, d, k, scale, iters=30, warmup=8):
    dev = torch.device("cuda", 0)
    b = make_spot_batch(n=n, d=d, k=k, seed=1004)
    img0, txt0 = b.image_features.to(dev), b.text_features.to(dev)
    s = torch.tensor(scale, device=dev, requires_grad=True)
    if kind == "clip":
        mod = ClipLoss()
        call = lambda i, t: mod(i, t, s)  # noqa: E731
    else:
        mod = SpatialLoss(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
                          neighbor_alpha_scale=0.5, float32_logits=True)
        ids, tids = b.tile_ids.to(dev), b.tile_ids.clone().to(dev)
        nbr, alpha = b.neighbor_tile_ids.to(dev), b.neighbor_alphas.to(dev)
        call = lambda i, t: mod(i, t, s, ids, tids, nbr, alpha)  # noqa: E731
    times = []
    keep = None
    for it in range(warmup + iters):
        img = img0.detach().requires_grad_(True)
        txt = txt0.detach().requires_grad_(True)
        s.grad = None
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = call(img, txt)["contrastive_loss"]
        loss.backward()
        e.record()
        torch.cuda.synchronize()
        keep = (loss, img, txt)
        if it >= warmup:
            times.append(a.elapsed_time(e))
    loss, img, txt = keep
    ms = statistics.median(times)
    rows = sample_rows_for(n, 8, 4, seed=n)
    ref = blockwise_oracle(b.image_features.bfloat16().float().numpy(), b.text_features.bfloat16().float().numpy(), scale,
                           b.tile_ids.numpy(), b.tile_ids.numpy(), b.neighbor_tile_ids.numpy() if k else None,
                           b.neighbor_alphas.numpy() if k else None, 1, 40.0 if kind == "spatial" else None,
                           0.05 if kind == "spatial" else 0.0, 0.5, True, True, rows, kind=kind)
    gi = img.grad[rows].double().cpu().numpy()
    alg = 6.0 * n * n * d / (ms * 1e-3) / 1e12
    print(json.dumps({"config": f"{kind} N={n} D={d} K={k} logit_scale={scale}", "ms_per_step": ms,
                      "pairs_per_s": n / (ms * 1e-3), "algorithmic_tflops": alg, "algorithmic_frac": alg / PEAK,
                      "executed_frac": 2 * alg / PEAK, "loss": float(loss.detach()),
                      "loss_rel_vs_oracle": abs(float(loss.detach()) - ref.loss[0]) / abs(ref.loss[0]),
                      "d_scale_rel_vs_oracle": abs(float(s.grad) - ref.d_scale[0]) / abs(ref.d_scale[0]),
                      "d_image_rows_err_of_max": float(np.abs(gi - ref.d_image_rows).max() / np.abs(ref.d_image_rows).max())}),
          flush=True)


if __name__ == "__main__":
    run("clip", 4096, 512, 0, 14.2857)      # BASELINE configs[1]
    run("spatial", 16384, 512, 8, 55.0)     # BASELINE configs[2]
