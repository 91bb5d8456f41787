#!/usr/bin/env python
"""BASELINE configs[4] sweep on one GPU: N in {8192 .. 131072}, D in {512, 768, 1024}, precision bf16 / fp32.
Prints one JSON line per point: ms per fwd+bwd step (CUDA events, median of --iters after --warmup), pairs/s/GPU,
algorithmic and executed tensor fractions of MEASURED_PEAKS.json and -- for N <= --parity-max-n -- loss / d_scale /
sampled gradient rows against the blockwise fp64 CPU oracle (oracle/blockwise_oracle.py) on the operand values the
mode sees (bf16-rounded inputs for bf16, the fp32 inputs for the fp32-accurate mode).  Beyond that size the Euler
identity  sum_i <x_i, dL/dx_i> = s_eff dL/ds  (exact when the temperature regulariser is off; reported as a relative
residual otherwise) stands in.  Needs a B200; developer tool, not part of the bench contract.

    python tools/sweep.py --n 8192 32768 131072 --d 768 1024 --precision bf16 fp32 > profiles/r2_sweep.jsonl
"""
import argparse
import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.blockwise_oracle import blockwise_oracle, sample_rows_for  # noqa: E402  (checker only)
from spatial_clip_b200 import SpatialLoss  # noqa: E402
from spatial_clip_b200.synth import make_spot_batch  # noqa: E402

CFG = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
           neighbor_alpha_scale=0.5, float32_logits=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[8192, 16384, 32768, 65536, 131072])
    ap.add_argument("--d", type=int, nargs="+", default=[768, 1024])
    ap.add_argument("--precision", nargs="+", default=["bf16", "fp32"])
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--parity-max-n", type=int, default=16384)
    args = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("bf16_tflops", 1590.0))
    dev = torch.device("cuda", 0)
    for n in args.n:
        for d in args.d:
            b = make_spot_batch(n=n, d=d, k=args.k, seed=1004)
            img0, txt0 = b.image_features.to(dev), b.text_features.to(dev)
            ids, nbr, alpha = b.tile_ids.to(dev), b.neighbor_tile_ids.to(dev), b.neighbor_alphas.to(dev)
            tids = ids.clone()
            for prec in args.precision:
                mod = SpatialLoss(**CFG, precision=prec)
                scale = torch.tensor(55.0, device=dev, requires_grad=True)
                times = []
                for it in range(args.warmup + args.iters):
                    img = img0.detach().requires_grad_(True)
                    txt = txt0.detach().requires_grad_(True)
                    scale.grad = None
                    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    loss = mod(img, txt, scale, ids, tids, nbr, alpha)["contrastive_loss"]
                    loss.backward()
                    e.record()
                    torch.cuda.synchronize()
                    if it >= args.warmup:
                        times.append(a.elapsed_time(e))
                ms = statistics.median(times)
                alg = 6.0 * n * n * d / (ms * 1e-3) / 1e12
                mma = 3.0 if prec == "fp32" else 1.0
                slices = 1 if d <= 512 else 2  # the backward recomputes the similarity tile once per D slice
                executed = (2.0 + 2.0 * (1 + slices)) * mma * n * n * d * 2 / (ms * 1e-3) / 1e12
                out = {"n": n, "d": d, "k": args.k, "precision": prec, "ms_per_step": ms,
                       "pairs_per_s": n / (ms * 1e-3), "loss": float(loss.detach()), "algorithmic_tflops": alg,
                       "algorithmic_frac": alg / peak, "executed_tflops": executed, "executed_frac": executed / peak}
                s_eff = 40.0
                xr = img0 if prec == "fp32" else img0.bfloat16().float()
                e_i = float((img.grad.double() * xr.double()).sum())
                out["euler_residual"] = abs(e_i - s_eff * float(scale.grad)) / abs(s_eff * float(scale.grad))
                if n <= args.parity_max_n:
                    t0 = time.perf_counter()
                    rows = sample_rows_for(n, 8, 4, seed=n + d)
                    conv = (lambda t: t) if prec == "fp32" else (lambda t: t.bfloat16().float())
                    ref = blockwise_oracle(conv(b.image_features).numpy(), conv(b.text_features).numpy(), 55.0,
                                           b.tile_ids.numpy(), b.tile_ids.numpy(), b.neighbor_tile_ids.numpy(),
                                           b.neighbor_alphas.numpy(), 1, 40.0, 0.05, 0.5, True, True, rows)
                    gi = img.grad[rows].double().cpu().numpy()
                    gt = txt.grad[rows].double().cpu().numpy()
                    out["parity"] = {
                        "loss_rel": abs(float(loss) - ref.loss[0]) / abs(ref.loss[0]),
                        "d_scale_rel": abs(float(scale.grad) - ref.d_scale[0]) / abs(ref.d_scale[0]),
                        "d_image_rows_err_of_max": float(np.abs(gi - ref.d_image_rows).max() / np.abs(ref.d_image_rows).max()),
                        "d_text_rows_err_of_max": float(np.abs(gt - ref.d_text_rows).max() / np.abs(ref.d_text_rows).max()),
                        "sampled_rows": len(rows), "oracle_seconds": time.perf_counter() - t0}
                print(json.dumps(out), flush=True)
                del mod
                import spatial_clip_b200

                spatial_clip_b200.release_cuda_graphs()
            del img0, txt0
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
