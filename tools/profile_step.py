#!/usr/bin/env python
"""Developer tool: Kineto timeline summary of one fwd+bwd step (works under torchrun for W > 1).

    python tools/profile_step.py            # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/profile_step.py
"""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from spatial_clip_b200 import SpatialLoss  # noqa: E402
from spatial_clip_b200.synth import make_spot_batch  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
full = make_spot_batch(n=bench.N_GLOBAL, d=bench.D, k=bench.K, seed=bench.SEED)
loc = full.rank_slice(rank, world)
inp = {k: v.to(dev) for k, v in dict(img=loc.image_features, txt=loc.text_features, ids=loc.tile_ids,
                                     nbr=loc.neighbor_tile_ids, alpha=loc.neighbor_alphas).items()}
scale = torch.tensor(bench.SCALE, device=dev, requires_grad=True)
mod = SpatialLoss(**bench.SPATIAL_CFG)


def step():
    img = inp["img"].detach().requires_grad_(True)
    txt = inp["txt"].detach().requires_grad_(True)
    scale.grad = None
    out = mod(img, txt, scale, inp["ids"], inp["ids"], inp["nbr"], inp["alpha"])["contrastive_loss"]
    out.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
n_steps = 5
t0 = time.perf_counter()
for _ in range(n_steps):
    step()
cpu_ms = (time.perf_counter() - t0) * 1e3 / n_steps
torch.cuda.synchronize()
if os.environ.get("SCL_CPROFILE") and rank == 0:
    import cProfile
    import pstats

    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(40)
elif os.environ.get("SCL_CPROFILE"):
    for _ in range(20):
        step()
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(n_steps):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = {}
    for e in evs:
        tot.setdefault(e.name[:70], [0, 0.0])
        tot[e.name[:70]][0] += 1
        tot[e.name[:70]][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    busy = sum(v[1] for v in tot.values())
    span = (max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs))
    print(f"world {world}: CPU launch time per step {cpu_ms:.3f} ms; GPU busy {busy / n_steps / 1e3:.3f} ms/step; "
          f"GPU span {span / n_steps / 1e3:.3f} ms/step")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {k:70s} n/step={v[0] / n_steps:5.1f} us/step={v[1] / n_steps:9.1f}")
if world > 1:
    dist.destroy_process_group()
