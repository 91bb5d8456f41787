#!/usr/bin/env python
"""Headline benchmark: fused multi-positive CLIP loss fwd+bwd, global batch 32768, D=512, K=8.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = SpatialLoss forward + backward (grads to image features, gene features, logit_scale) over
one synthetic global batch (SURVEY.md §8d generator), through the reference-facing module API.
Prints ONE JSON line on rank 0.  Metric/unit/config follow BASELINE.json; ``value`` is whole-job
pairs/s with inputs resident in HBM (K steps in one timed region, CUDA events, max over ranks; the per-step median is
reported next to it), ``e2e`` the same through the module with pinned host buffers (H2D of the step's inputs + D2H of
the loss inside the timed region; the gradients stay on the device, where the optimiser of a training step consumes
them), ``roofline`` is for the dominant kernel (bwd_rows) from CUDA events recorded around its launches,
``parity`` compares this very run's loss / d logit_scale / sampled gradient rows with the blockwise fp64 CPU oracle
(outside the timed region; every rank's rows, so the N > 1 lines carry parity over NCCL), and ``cpu_baseline`` is the
reference's own loss module (oracle/_ref, else the oracle's torch port) timed on this box's host cores on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_GLOBAL = 32768
D = 512
K = 8
SEED = 1004
SCALE = 55.0  # cap (40) active, as in a trained run
SPATIAL_CFG = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=40.0, temp_reg_weight=0.05,
                   neighbor_alpha_scale=0.5, float32_logits=True)  # configs/loss/spatial.yaml


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(burst=float(j["bf16_tflops"]), sustained=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                    hbm=float(j["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class NvmlSampler:
    """SM clock and throttle reasons polled through NVML every ~2 ms while the timed region runs.  The timed region of
    the default run lasts ~60 ms, which nvidia-smi's 200 ms loop (ClockSampler, the fallback) barely sees."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, torch_device, period_s=0.002):
        self.dev = torch_device
        self.period = period_s
        self.sm, self.power, self.mask = [], [], 0
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self._handle = None

    def start(self) -> bool:
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            handle = None
            try:
                uuid = str(torch.cuda.get_device_properties(self.dev).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                idx = self.dev.index or 0
                if ids and all(v.isdigit() for v in ids) and idx < len(ids):
                    idx = int(ids[idx])
                handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)  # must work before the thread relies on it
            self._nvml, self._handle = pynvml, handle
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
            return True
        except Exception:
            return False

    def _run(self):
        nv, h = self._nvml, self._handle
        n = 0
        while not self._stop.is_set():
            try:  # two light queries per poll; the power reading (slower) only every 16th
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                if n % 16 == 0:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            n += 1
            time.sleep(self.period)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2.0)
        return self.summary(self.sm, self.sm_max, self.mask, self.power, self.period)

    @classmethod
    def summary(cls, sm, sm_max, mask, power, period):
        sm = sorted(sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": sm_max,
                "reasons": [name for bit, name in cls.REASONS if mask & bit], "samples": len(sm),
                "power_w_max": max(power) if power else None, "source": f"nvml, {period * 1e3:.0f} ms period"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own loss module on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(rows: int, steps: int = 1, warmup: int = 0):
    """Time the reference's per-rank step for `rows` local rows of the N=32768 workload (one rank of a W = N/rows
    emulation; a W=1 step at N=32768 needs ~90 GB of dense [N, N] fp32 temporaries).  The step is the reference's own
    SpatialLoss (oracle/_ref bytecode of src/models/components/losses.py, kind "reference") when oracle/_ref travels
    with the repo, else the oracle's torch port with the same cost structure (kind "port").

    Every rank of the reference rebuilds its two id -> column dicts over all N ids (losses.py:92-93); a W=1 step would
    build them once for N rows.  That part is therefore timed separately (t_dict) and charged once per N rows:
    seconds per pair = (t_step - t_dict) / rows + t_dict / N.  Returns a dict with both the raw rank-step and the
    W=1-equivalent figures."""
    import torch

    from oracle import ref_loader
    from oracle.torch_port import spatial_rank_step
    from spatial_clip_b200.synth import make_spot_batch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = make_spot_batch(n=N_GLOBAL, d=D, k=K, seed=SEED)
    sl = slice(0, rows)
    world = max(1, N_GLOBAL // rows)
    use_ref = ref_loader.available() and ref_loader.load_reference() is not None
    ctor = dict(local_loss=True, gather_with_grad=True, cap_logit_scale=SPATIAL_CFG["cap_logit_scale"],
                temp_reg_weight=SPATIAL_CFG["temp_reg_weight"], float32_logits=True,
                neighbor_alpha_scale=SPATIAL_CFG["neighbor_alpha_scale"])
    times, dict_times = [], []
    for it in range(warmup + steps):
        img_l = b.image_features[sl].clone().requires_grad_(True)
        txt_l = b.text_features[sl].clone().requires_grad_(True)
        s = torch.tensor(SCALE, requires_grad=True)
        t0 = time.perf_counter()
        if use_ref:
            ref_loader.reference_rank_step(img_l, txt_l, b.image_features, b.text_features, s, b.tile_ids,
                                           b.tile_ids[sl], b.neighbor_tile_ids[sl], b.neighbor_alphas[sl], 0, world,
                                           ctor)
        else:
            spatial_rank_step(img_l, txt_l, b.image_features, b.text_features, s, b.tile_ids, b.neighbor_tile_ids[sl],
                              b.neighbor_alphas[sl], 0, cap=ctor["cap_logit_scale"],
                              temp_reg_weight=ctor["temp_reg_weight"], alpha_scale=ctor["neighbor_alpha_scale"])
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()  # the two dict comprehensions of losses.py:92-93, alone
        _a = {tid.item(): i for i, tid in enumerate(b.tile_ids)}
        _b = {tid.item(): i for i, tid in enumerate(b.tile_ids)}
        dd = time.perf_counter() - t1
        if it >= warmup:
            times.append(dt)
            dict_times.append(dd)
    sec = sum(times) / len(times)
    t_dict = min(sum(dict_times) / len(dict_times), 0.9 * sec)
    sec_per_pair = (sec - t_dict) / rows + t_dict / N_GLOBAL
    return {"pairs_per_s": 1.0 / sec_per_pair, "rank_step_s": sec, "dict_s": t_dict, "rank_step_pairs_per_s": rows / sec,
            "cores": cores, "kind": "reference" if use_ref else "port", "rows": rows, "world": world, "steps": len(times)}


def cpu_sample_text(r):
    what = ("the reference's own SpatialLoss (oracle/_ref bytecode of src/models/components/losses.py)"
            if r["kind"] == "reference" else "oracle/torch_port.py (torch port of the reference)")
    return (f"{r['rows']} local rows x N={N_GLOBAL} columns (one rank of a W={r['world']} single-process emulation of the "
            f"same workload), fwd+bwd, {what}, torch CPU fp32, {r['cores']} threads, mean of {r['steps']} rank-steps of "
            f"{r['rank_step_s']:.2f} s; the {r['dict_s']:.2f} s of per-rank id-dict building is charged once per N rows "
            f"(W=1 equivalent; the raw rank-step rate is {r['rank_step_pairs_per_s']:.0f} pairs/s)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_sample(1024, steps=max(1, min(args.steps, 12)), warmup=min(args.warmup, 1))
    pps = r["pairs_per_s"]
    line = {
        "impl": "reference", "metric": "contrastive loss fwd+bwd pairs/sec (global batch 32768, D=512)",
        "value": pps, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": N_GLOBAL / pps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {**workload_config(args.gpus), "precision": args.precision},
        "cpu_baseline": {"value": pps, "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": cpu_sample_text(r)},
        "e2e": {"value": pps, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "SpatialLoss (multi-positive, K=8 neighbours) fwd+bwd, BASELINE configs[2]/[3] loss at the "
                        "metric's global batch", "global_batch": N_GLOBAL, "embed_dim": D, "k_neighbors": K,
            "local_batch": N_GLOBAL // n_gpus, "logit_scale": SCALE, **SPATIAL_CFG,
            "parallelism": f"dp{n_gpus} (row shards, local_loss feature all-gather over NCCL)",
            "tile_ids": "image-side and text-side id vectors are distinct tensors (as the reference's collate makes them)",
            "l2": "a step streams 134 MB of fp32 inputs, 67 MB of bf16 operand copies and 134 MB of gradients (> the "
                  "126 MB L2), so consecutive steps do not find their inputs in L2; no explicit flush in either loop"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
PARITY_ROWS_PER_RANK = 8


def parity_block(full, world, loss_ranks, ds_ranks, rows, gi_rows, gt_rows, split):
    """This run's results against the blockwise fp64 CPU oracle (oracle/blockwise_oracle.py) on the operand values the
    kernels see (bf16-rounded inputs in the bf16 mode, the fp32 inputs in the fp32-accurate mode).  Rank 0 only,
    outside every timed region."""
    import numpy as np

    from oracle.blockwise_oracle import blockwise_oracle

    t0 = time.perf_counter()
    img, txt = full.image_features, full.text_features
    if not split:
        img, txt = img.bfloat16().float(), txt.bfloat16().float()
    ref = blockwise_oracle(img.numpy(), txt.numpy(), SCALE, full.tile_ids.numpy(), full.tile_ids.numpy(),
                           full.neighbor_tile_ids.numpy(), full.neighbor_alphas.numpy(), world,
                           SPATIAL_CFG["cap_logit_scale"], SPATIAL_CFG["temp_reg_weight"],
                           SPATIAL_CFG["neighbor_alpha_scale"], SPATIAL_CFG["local_loss"],
                           SPATIAL_CFG["gather_with_grad"], rows)
    loss_rel = float(np.max(np.abs(np.asarray(loss_ranks) - ref.loss) / np.abs(ref.loss)))
    ds_rel = float(np.max(np.abs(np.asarray(ds_ranks) - ref.d_scale) / np.abs(ref.d_scale)))
    gi_rel = float(np.abs(gi_rows - ref.d_image_rows).max() / np.abs(ref.d_image_rows).max())
    gt_rel = float(np.abs(gt_rows - ref.d_text_rows).max() / np.abs(ref.d_text_rows).max())
    gates = {"loss_rel": 5e-5 if split else 2e-5, "d_scale_rel": 1e-3, "grad_rel_of_max": 1e-4 if split else 5e-3}
    ok = loss_rel <= gates["loss_rel"] and ds_rel <= gates["d_scale_rel"] and max(gi_rel, gt_rel) <= gates["grad_rel_of_max"]
    return {"oracle": "oracle/blockwise_oracle.py (fp64, closed-form gradients, W-rank emulation)", "ranks": world,
            "loss_rank0": float(loss_ranks[0]), "loss_rank0_oracle": float(ref.loss[0]), "loss_rel_max_over_ranks": loss_rel,
            "d_scale_rel_max_over_ranks": ds_rel, "sampled_rows": len(rows), "d_image_rows_err_of_max": gi_rel,
            "d_text_rows_err_of_max": gt_rel, "gates": gates, "ok": bool(ok),
            "transport": "nccl" if world > 1 else "single rank", "oracle_seconds": time.perf_counter() - t0}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from spatial_clip_b200 import SpatialLoss, losses
    from spatial_clip_b200.synth import make_spot_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    full = make_spot_batch(n=N_GLOBAL, d=D, k=K, seed=SEED)
    loc = full.rank_slice(rank, world)
    b_local = N_GLOBAL // world
    # image-side and text-side tile ids are separate tensors with equal contents, as the reference's collate builds
    # them (spatial_datamodule.py:126-127)
    host = dict(img=loc.image_features.pin_memory(), txt=loc.text_features.pin_memory(),
                ids=loc.tile_ids.pin_memory(), tids=loc.tile_ids.clone().pin_memory(),
                nbr=loc.neighbor_tile_ids.pin_memory(), alpha=loc.neighbor_alphas.pin_memory())
    dev_in = {k: v.to(dev) for k, v in host.items()}
    scale = torch.tensor(SCALE, device=dev, requires_grad=True)
    mod = SpatialLoss(**SPATIAL_CFG, precision=args.precision, cuda_graphs=not args.no_graphs)
    split = args.precision == "fp32"
    ops = losses._ops()

    def step(inp):
        img = inp["img"].requires_grad_(True)
        txt = inp["txt"].requires_grad_(True)
        scale.grad = None
        out = mod(img, txt, scale, inp["ids"], inp["tids"], inp["nbr"], inp["alpha"])["contrastive_loss"]
        out.backward()
        return out, img.grad, txt.grad

    cpu_group = dist.new_group(backend="gloo") if (world > 1 and args.barrier == "gloo") else None

    def sync_all():
        """Barrier over all ranks + device synchronisation (the bracket of every timed region).  With --barrier gloo
        the rendezvous runs on a CPU process group: an eager NCCL collective between graph replays of the same
        communicator showed up as one 10-18 ms step right after it."""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group) if cpu_group is not None else dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then EXACTLY K steps in one timed region
    # (the warm-up keeps each step's results alive into the next step exactly like the timed loop does, so that the
    # caching allocator already owns two sets of gradient buffers: otherwise the SECOND timed step calls cudaMalloc --
    # measured as one 10-25 ms step per run, which every rank then waits for)
    loss = g_img = g_txt = None
    for _ in range(args.warmup):
        loss, g_img, g_txt = step({k: v.detach() for k, v in dev_in.items()})
    # the clock sampler starts BEFORE the barrier that aligns the ranks: NVML initialisation takes 5-20 ms on rank 0,
    # and a rank that enters the timed loop late makes every other rank's first step wait for it at the first exchange
    # (measured: one 18 ms step out of 50 at 8 ranks)
    sampler = NvmlSampler(dev, period_s=max(args.sampler_period_ms, 0.5) * 1e-3)
    if rank == 0 and args.sampler_period_ms > 0 and not sampler.start():  # no NVML binding: nvidia-smi loop instead
        sampler = ClockSampler(local_rank)
        sampler.start()
    # no cyclic-GC pauses inside the timed region either: a multi-millisecond collection on ONE rank stalls every rank
    # at the next exchange (the step is ~1 ms at 8 ranks); reference counting keeps freeing tensors as usual
    gc.collect()
    gc.disable()
    sync_all()
    launches0 = ops.launches
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    marks[0].record()
    for i in range(args.steps):
        loss, g_img, g_txt = step({k: v.detach() for k, v in dev_in.items()})
        marks[i + 1].record()
    sync_all()
    gc.enable()
    launches = (ops.launches - launches0) // max(1, args.steps)
    clocks = sampler.stop() if (rank == 0 and args.sampler_period_ms > 0) else None
    ms = marks[0].elapsed_time(marks[-1]) / args.steps
    raw_steps = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    per_step = sorted(raw_steps)
    med = per_step[len(per_step) // 2]
    slow = float(sum(1 for v in raw_steps if v > 1.5 * med))  # steps more than 1.5x this rank's median
    t = torch.tensor([ms, med, per_step[0], per_step[int(0.9 * (len(per_step) - 1))], per_step[-1], slow], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, med = float(t[0].item()), float(t[1].item())
    step_spread = {"min": float(t[2]), "p90": float(t[3]), "max": float(t[4]), "steps_over_1.5x_median": int(t[5]),
                   "rank0_first_steps": [round(v, 3) for v in raw_steps[:4]],
                   "rank0_slowest_step_index": int(max(range(len(raw_steps)), key=lambda i: raw_steps[i]))}
    loss_val = float(loss.detach())

    # ---- host cost of a step: wall time to ENQUEUE ten steps on an idle device (no synchronisation inside); when this
    # exceeds the device time per step the job is host-bound (the case to watch at 8 ranks, ~1 ms of kernels per step)
    sync_all()
    h0 = time.perf_counter()
    for _ in range(10):
        step({k: v.detach() for k, v in dev_in.items()})
    host_ms = (time.perf_counter() - h0) * 1e3 / 10
    sync_all()
    th = torch.tensor([host_ms], device=dev)
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    host_ms = float(th.item())

    # ---- parity inputs: every rank's loss, d logit_scale and a fixed sample of its gradient rows (last timed step)
    from oracle.blockwise_oracle import sample_rows_for

    rows_all = sample_rows_for(N_GLOBAL, world, PARITY_ROWS_PER_RANK, seed=SEED)
    my_rows = torch.tensor([r - rank * b_local for r in rows_all if rank * b_local <= r < (rank + 1) * b_local], device=dev)
    mine = torch.cat([torch.stack([loss.detach().float().reshape(()), scale.grad.detach().float().reshape(())]),
                      g_img[my_rows].float().reshape(-1), g_txt[my_rows].float().reshape(-1)])
    if world > 1:
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    allv = [v.cpu().double().numpy() for v in allv]

    # ---- dominant-kernel timing: two extra steps with CUDA events around the tensor-core launches (the backward then
    # goes out as three host calls per direction instead of one; the timed steps above use the one-call route)
    ops.kernel_events = {}
    for _ in range(2):
        step({k: v.detach() for k, v in dev_in.items()})
    sync_all()
    kernel_events, ops.kernel_events = ops.kernel_events, None

    # ---- optional per-phase timeline of one step (developer aid for the multi-GPU fixed costs): CUDA events at named
    # points of forward and backward, reported as the GPU time between consecutive marks, max over ranks
    timeline = None
    if args.timeline:
        for _ in range(3):
            ops.timeline = []
            step({k: v.detach() for k, v in dev_in.items()})
            sync_all()
        tl, ops.timeline = ops.timeline, None
        names = [f"{tl[i][0]} -> {tl[i + 1][0]}" for i in range(len(tl) - 1)]
        tt = torch.tensor([tl[i][1].elapsed_time(tl[i + 1][1]) for i in range(len(tl) - 1)], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        timeline = {n: round(float(v), 4) for n, v in zip(names, tt.tolist())}
        timeline["sum_ms"] = round(float(tt.sum()), 4)

    # ---- end to end: pinned host inputs -> H2D -> module fwd+bwd -> loss D2H, every step.
    # Double-buffered input pipeline: step i+1's H2D copy is enqueued on a copy stream before step i's loss is
    # read back, so PCIe overlaps the kernels; every step's inputs are still copied inside the timed region.
    # The step's inputs travel as ONE pinned staging buffer (what a collate function writing into pinned memory hands
    # over): one H2D copy per step, the six tensors are views of the device copy.
    layout, total = {}, 0
    for k, v in host.items():
        layout[k] = (total, v.numel() * v.element_size(), v.dtype, tuple(v.shape))
        total = (total + v.numel() * v.element_size() + 255) // 256 * 256
    host_pack = torch.empty(total, dtype=torch.uint8).pin_memory()
    for k, v in host.items():
        off, nbytes, _, _ = layout[k]
        host_pack[off:off + nbytes] = v.contiguous().view(torch.uint8).reshape(-1)
    h2d = int(total)
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def enqueue_copy():
        with torch.cuda.stream(copy_stream):
            pack = host_pack.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        bufs = {k: pack[off:off + nbytes].view(dt).view(shape) for k, (off, nbytes, dt, shape) in layout.items()}
        bufs["_pack"] = pack
        return bufs, ev

    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()  # pinned landing slots for the per-step loss

    def e2e_loop(n_steps, deferred):
        """Every step: H2D of its inputs (copy stream, one step ahead), module fwd+bwd, D2H read of its loss.
        deferred: the loss lands in a pinned slot and the host reads step i's value while step i+1 is already
        enqueued (a training loop logs the loss one step late instead of draining the GPU every step); the last
        loss is read before the loop returns.  Otherwise the host blocks on every step's loss."""
        nxt = enqueue_copy()
        last, pending = None, None
        for i in range(n_steps):
            bufs, ev = nxt
            main_stream.wait_event(ev)
            bufs.pop("_pack").record_stream(main_stream)
            l, _, _ = step(bufs)
            if deferred:
                loss_host[i % 2].copy_(l.detach(), non_blocking=True)  # D2H of this step's result
                done = torch.cuda.Event()
                done.record(main_stream)
            if i + 1 < n_steps:
                nxt = enqueue_copy()
            if not deferred:
                last = float(l.detach().to("cpu"))  # D2H read of the step's result (synchronises)
                continue
            if pending is not None:  # read the previous step's loss on the host
                pending[1].synchronize()
                last = float(loss_host[pending[0]])
            pending = (i % 2, done)
        if deferred:
            pending[1].synchronize()
            last = float(loss_host[pending[0]])
        return last

    e2e_steps = max(3, min(args.steps, 20))

    def time_e2e(deferred):
        e2e_loop(min(4, max(1, args.warmup)), deferred)
        gc.collect()
        gc.disable()
        sync_all()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_loop(e2e_steps, deferred)
        t1.record()
        sync_all()
        gc.enable()
        return t0.elapsed_time(t1) / e2e_steps

    e2e_note = ("one packed H2D copy per step on a copy stream, one step ahead; every step's loss is copied D2H into pinned memory and read by "
                "the host one step behind; the gradients (2 x B_l x D) stay on the device for the optimiser; two timed "
                "loops, the faster one is reported (both in runs_ms)")
    try:
        e2e_runs = [time_e2e(True), time_e2e(True)]
    except Exception as exc:  # fall back to the blocking read-back
        e2e_note = ("double-buffered H2D on a copy stream, loss read back (blocking) every step; deferred read-back "
                    f"failed: {type(exc).__name__}: {exc}")[:400]
        e2e_runs = [time_e2e(False), time_e2e(False)]
    te = torch.tensor(e2e_runs, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)  # per loop: the slowest rank
    e2e_runs = [float(v) for v in te.tolist()]
    e2e_ms = min(e2e_runs)

    del loss, g_img, g_txt
    losses.release_cuda_graphs()  # captured NCCL kernels must go before the process group does
    if rank != 0:
        finish(world)
        return

    pk = peaks()
    prof = {}
    tp = ROOT / "profiles" / "r2_traffic.json"  # per-launch numbers of the committed ncu --set full capture (1 GPU)
    if tp.exists():
        prof = json.loads(tp.read_text())
    traffic = None
    if world == 1 and "bwd_rows_pair_kernel" in prof:
        tj = prof["bwd_rows_pair_kernel"]
        traffic = tj.get("dram_bytes_read", 0) + tj.get("dram_bytes_write", 0)
    # dominant kernel: bwd_rows, two launches per step (d image, d gene)
    ev = kernel_events.get("bwd_rows", [])
    k_ms = sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev))
    fev = kernel_events.get("fwd_rowstats", [])
    f_ms = sum(a.elapsed_time(b) for a, b in fev) / max(1, len(fev))
    alg_flops_launch = 2.0 * b_local * N_GLOBAL * D  # the dX GEMM; the z recompute is not algorithmic work
    mma_factor = 3.0 if split else 1.0  # fp32-accurate mode: three bf16 products per algorithmic one
    achieved = alg_flops_launch / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    pairs_per_s = N_GLOBAL / (ms * 1e-3)
    nd = float(N_GLOBAL) * D
    per_gpu_pairs = (N_GLOBAL / world) / (ms * 1e-3)
    step_alg_tflops = per_gpu_pairs * 6.0 * nd / 1e12  # F_alg = 6 N D per pair
    executed_nd = 12.0 * mma_factor  # two forward passes + two backward passes (recompute + gradient GEMM each)

    # ---- parity of this run (loss / d_scale of every rank, sampled gradient rows), CPU oracle, untimed
    n_s = PARITY_ROWS_PER_RANK
    gi = np.concatenate([v[2:2 + n_s * D].reshape(n_s, D) for v in allv])
    gt = np.concatenate([v[2 + n_s * D:2 + 2 * n_s * D].reshape(n_s, D) for v in allv])
    try:
        parity = parity_block(full, world, [v[0] for v in allv], [v[1] for v in allv], rows_all, gi, gt, split)
    except Exception as exc:  # never lose the bench line to the checker
        parity = {"ok": None, "error": f"{type(exc).__name__}: {exc}"[:300]}

    cpu = cpu_reference_sample(1024, steps=12, warmup=1)  # bounded sample: ~10 s of host work

    line = {
        "metric": "contrastive loss fwd+bwd pairs/sec (global batch 32768, D=512)",
        "value": pairs_per_s, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "median_ms_per_step": med, "step_ms_spread": step_spread, "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "bf16x2 (fp32-accurate: bf16 hi/lo operand pairs, fp32 accumulate)" if split else "bf16",
        "data": "synthetic", "config": {**workload_config(world), "precision": args.precision},
        "pairs_per_s_per_gpu": pairs_per_s / world,
        "loss": loss_val,
        "e2e": {"value": N_GLOBAL / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms, "runs_ms": e2e_runs, "steps": e2e_steps,
                "pipeline": e2e_note},
        "gpu_launches": int(launches), "cuda_graphs": not args.no_graphs,
        "host_enqueue_ms_per_step": host_ms,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "bwd_rows_pair_kernel", "achieved": achieved, "peak": pk["burst"],
                     "unit": "TFLOP/s", "frac": achieved / pk["burst"], "traffic": traffic,
                     "peak_source": pk["source"], "launch_ms": k_ms, "launches_per_step": 2,
                     "timed": "CUDA events around the launches in two extra steps after the timed loop",
                     "algorithmic_flops_per_launch": alg_flops_launch,
                     "frac_of_sustained": achieved / pk["sustained"],
                     "executed_tflops": 2.0 * mma_factor * achieved,
                     "fwd_rowstats_launch_ms": f_ms,
                     "fwd_rowstats_tflops": 2.0 * mma_factor * b_local * N_GLOBAL * D / (f_ms * 1e-3) / 1e12 if f_ms > 0 else None,
                     "step_algorithmic_tflops_per_gpu": step_alg_tflops,
                     "step_frac_of_burst": step_alg_tflops / pk["burst"],
                     # BASELINE.md §3 "tensor_pipe_util": executed MMA flop / t / peak with executed = 8 N D per pair
                     # (one forward pass + two recompute-and-gradient passes); this build still runs TWO forward
                     # passes, i.e. executes 12 N D per pair, reported separately
                     "tensor_pipe_util_8nd": step_alg_tflops * 8.0 / 6.0 / pk["burst"],
                     "executed_nd_per_pair": executed_nd,
                     "executed_tflops_per_gpu": step_alg_tflops * executed_nd / 6.0,
                     "executed_frac_of_burst": step_alg_tflops * executed_nd / 6.0 / pk["burst"],
                     "ncu_tensor_pipe_active_pct": prof.get("tensor_pipe_active_pct"),
                     "note": "achieved counts the dX GEMM only (algorithmic); the launch also recomputes the similarity "
                             "tile once (executed = 2x)"},
        "parity": parity,
        "timeline_ms": timeline,
        "cpu_baseline": {"value": cpu["pairs_per_s"], "unit": "pairs/s", "cores": cpu["cores"], "kind": cpu["kind"],
                         "sample": cpu_sample_text(cpu)},
    }
    print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """Tear the process group down, but never let a teardown that waits on another rank hold the job: the result is
    out already, so after 20 s the process exits on its own (exit code 0)."""
    if world <= 1:
        return
    import torch.distributed as dist

    def bail():
        sys.stdout.flush()
        os._exit(0)

    t = threading.Timer(20.0, bail)
    t.daemon = True
    t.start()
    try:
        dist.destroy_process_group()
    finally:
        t.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)  # 10 warm-up + 50 timed
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--sampler-period-ms", type=float, default=2.0, help="clock sampler period; 0 = no sampler")
    ap.add_argument("--barrier", choices=["nccl", "gloo"], default="nccl",
                    help="process group of the barriers that bracket the timed regions")
    ap.add_argument("--no-graphs", action="store_true", help="run every step eagerly (no CUDA-graph replay)")
    ap.add_argument("--timeline", action="store_true", help="add the per-phase GPU times of one step to the JSON line")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16",
                    help="fp32: the fp32-accurate mode (bf16 hi/lo operand pairs), BASELINE configs[4]'s 'fp32 vs bf16'")
    args = ap.parse_args()
    # at least three untimed steps: two eager calls per configuration precede the CUDA-graph capture, and the caching
    # allocator needs one more step to own both sets of gradient buffers; the JSON line reports what was run
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
