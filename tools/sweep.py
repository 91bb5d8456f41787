#!/usr/bin/env python
"""BASELINE configs[4] sweep on one GPU: N in {8192 .. 131072}, D in {512, 768, 1024}, precision bf16 / fp32.
Prints one JSON line per point:
ms per fwd+bwd step (CUDA events, median of --iters after --warmup), pairs/s/GPU, algorithmic and executed
tensor fractions of MEASURED_PEAKS.json.  Needs a B200; developer tool, not part of the bench contract.

    python tools/sweep.py --n 8192 32768 131072 --d 512 1024 --precision bf16 fp32
"""
import argparse
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from spatial_clip_b200 import SpatialLoss  # noqa: E402
from spatial_clip_b200.synth import make_spot_batch  # noqa: E402

CFG = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
           neighbor_alpha_scale=0.5, float32_logits=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[8192, 16384, 32768, 65536, 131072])
    ap.add_argument("--d", type=int, nargs="+", default=[512, 768, 1024])
    ap.add_argument("--precision", nargs="+", default=["bf16", "fp32"])
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("bf16_tflops", 1590.0))
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for n in args.n:
        for d in args.d:
            b = make_spot_batch(n=n, d=d, k=args.k, seed=1004)
            img0, txt0 = b.image_features.to(dev), b.text_features.to(dev)
            ids, nbr, alpha = b.tile_ids.to(dev), b.neighbor_tile_ids.to(dev), b.neighbor_alphas.to(dev)
            for prec in args.precision:
                mod = SpatialLoss(**CFG, precision=prec)
                scale = torch.tensor(55.0, device=dev, requires_grad=True)
                times = []
                for it in range(args.warmup + args.iters):
                    img = img0.detach().requires_grad_(True)
                    txt = txt0.detach().requires_grad_(True)
                    scale.grad = None
                    flush.fill_(1)
                    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    loss = mod(img, txt, scale, ids, ids, nbr, alpha)["contrastive_loss"]
                    loss.backward()
                    e.record()
                    torch.cuda.synchronize()
                    if it >= args.warmup:
                        times.append(a.elapsed_time(e))
                ms = statistics.median(times)
                alg = 6.0 * n * n * d / (ms * 1e-3) / 1e12
                mma = 3.0 if prec == "fp32" else 1.0
                print(json.dumps({"n": n, "d": d, "k": args.k, "precision": prec, "ms_per_step": ms,
                                  "pairs_per_s": n / (ms * 1e-3), "loss": float(loss.detach()),
                                  "algorithmic_tflops": alg, "algorithmic_frac": alg / peak,
                                  "executed_tflops": 2.0 * mma * alg, "tensor_pipe_util": 2.0 * mma * alg / peak}),
                      flush=True)
            del img0, txt0
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
