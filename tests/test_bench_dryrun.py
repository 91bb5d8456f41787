"""Dry run of bench.py's own control flow on the CPU: the CUDA runtime objects it touches (streams, events, pinned
memory, the device) are replaced by inert stand-ins and the compute backend by the TEST-ONLY emulated ops, at a
tiny problem size.  Nothing is measured -- the point is that every line of the timed loop, the end-to-end loop (both
read-back modes), the kernel-event modes and the JSON assembly executes without a GPU, so that a typo cannot cost the
round its bench line.  Runs in a subprocess because it patches torch globally."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]

HARNESS = r'''
import json, sys, types
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import torch
import torch.distributed.nn  # imported by the reference's loss module; must load before torch.device is replaced
import bench
from emulated_ops import EmulatedOps
from spatial_clip_b200 import losses

real_device = torch.device
class FakeEvent:
    def __init__(self, enable_timing=False): pass
    def record(self, stream=None): pass
    def elapsed_time(self, other): return 1.0
    def synchronize(self): pass
class FakeStream:
    def __init__(self, device=None): self.cuda_stream = 0
    def wait_event(self, ev): pass
class FakeCtx:
    def __init__(self, s): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
import torch.distributed as dist
real_init = dist.init_process_group
dist.init_process_group = lambda backend, **k: real_init("gloo")  # the N > 1 flow over gloo (env:// rendezvous)
torch.device = lambda *a, **k: real_device("cpu")
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.current_stream = lambda *a, **k: FakeStream()
torch.cuda.Stream = FakeStream
torch.cuda.Event = FakeEvent
torch.cuda.stream = FakeCtx
torch.Tensor.pin_memory = lambda self: self
torch.Tensor.record_stream = lambda self, s: None

class Ops(EmulatedOps):
    launches = 0
    kernel_events = None
    def split_cast(self, x, want_rows=True, want_cols=True): return x.clone(), x.clone()
    def bwd_rows(self, *a, **k):
        if self.kernel_events is not None:
            self.kernel_events.setdefault("bwd_rows", []).append((FakeEvent(), FakeEvent()))
        k.pop("split", None)
        return super().bwd_rows(*a, **k)
    def fwd_rowstats(self, *a, **k):
        if self.kernel_events is not None:
            self.kernel_events.setdefault("fwd_rowstats", []).append((FakeEvent(), FakeEvent()))
        return super().fwd_rowstats(*a, **k)
losses._set_ops_for_testing(Ops(round_bf16=False))
bench.N_GLOBAL, bench.D, bench.K = 256, 64, 4
orig_sample = bench.cpu_reference_sample
bench.cpu_reference_sample = lambda rows, steps=1, warmup=0: orig_sample(64, steps=1, warmup=0)
bench.PARITY_ROWS_PER_RANK = 3
if %(break_deferred)r:  # the deferred read-back raises -> the blocking loop must take over
    real_copy = torch.Tensor.copy_
    def bad_copy(self, src, non_blocking=False):
        if non_blocking and self.dim() == 0: raise RuntimeError("simulated failure of the pinned copy")
        return real_copy(self, src)
    torch.Tensor.copy_ = bad_copy
sys.argv = ["bench.py", "--steps", "3", "--warmup", "1", "--gpus", str(%(world)d)] + %(extra)r
bench.main()
'''


def _run(extra, break_deferred=False):
    code = HARNESS % {"root": str(ROOT), "extra": extra, "break_deferred": break_deferred, "world": 1}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def _run_world(extra, world=2):
    """The torchrun flow: one process per rank with RANK / WORLD_SIZE / MASTER_* in the environment."""
    import os
    import socket

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    code = HARNESS % {"root": str(ROOT), "extra": extra, "break_deferred": False, "world": world}
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                      text=True, env=env))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (_, err) in zip(procs, outs):
        assert p.returncode == 0, err[-3000:]
    lines = [ln for ln in outs[0][0].splitlines() if ln.startswith("{")]
    assert len(lines) == 1, outs[0][0][-2000:]
    for out, _ in outs[1:]:  # only rank 0 prints
        assert not [ln for ln in out.splitlines() if ln.startswith("{")]
    return json.loads(lines[0])


def test_bench_control_flow_runs_end_to_end():
    j = _run([])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "median_ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks",
                "roofline", "cpu_baseline", "parity"):
        assert key in j, key
    assert j["steps"] == 3 and j["n_gpus"] == 1 and j["unit"] == "pairs/s" and j["vs_baseline"] is None
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "pipeline"} <= set(j["e2e"])
    assert "one step behind" in j["e2e"]["pipeline"] and j["e2e"]["h2d_bytes_per_step"] > 0
    r = j["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["launches_per_step"] == 2
    assert {"value", "unit", "cores", "kind", "sample"} <= set(j["cpu_baseline"])
    assert j["cpu_baseline"]["kind"] in ("reference", "port")
    assert j["config"]["workload"] and "model" not in j["config"]
    # the emulated backend runs in fp64 on unrounded inputs, the oracle sees bf16-rounded ones: only the plumbing
    # (rows, ranks, shapes) is checked here, the gates themselves are for the GPU
    p = j["parity"]
    assert "error" not in p, p
    assert p["ranks"] == 1 and p["sampled_rows"] == 3 and p["loss_rel_max_over_ranks"] < 1e-2


def test_bench_falls_back_to_blocking_readback():
    j = _run([], break_deferred=True)
    assert "blocking" in j["e2e"]["pipeline"] and "simulated failure" in j["e2e"]["pipeline"]
    assert j["e2e"]["value"] > 0


def test_bench_control_flow_two_ranks():
    """N > 1: process group, barriers, max-over-ranks reductions, every rank in the kernel-event steps and the parity
    gather, rank 0 alone in the tail (oracle, CPU baseline, the JSON line) while the other ranks leave."""
    j = _run_world([], 2)
    assert j["n_gpus"] == 2 and j["config"]["local_batch"] == 128 and j["scaling"] == "strong"
    assert j["roofline"]["launches_per_step"] == 2 and j["e2e"]["value"] > 0
    p = j["parity"]
    assert "error" not in p, p
    assert p["ranks"] == 2 and p["sampled_rows"] == 6 and p["loss_rel_max_over_ranks"] < 1e-2


def test_reference_arm_prints_the_contract_line():
    code = r'''
import sys
sys.path.insert(0, %r)
import bench
bench.N_GLOBAL, bench.D, bench.K = 256, 64, 4
orig = bench.cpu_reference_sample
bench.cpu_reference_sample = lambda rows, steps=1, warmup=0: orig(64, steps=1, warmup=0)
sys.argv = ["bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"]
bench.main()
''' % str(ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    j = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert j["impl"] == "reference" and j["value"] > 0 and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["value"] == j["value"]


def test_graft_entry_smoke_control_flow():
    """__graft_entry__.smoke() with the same stand-ins: build(), the module call, the oracle comparison."""
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import torch
from emulated_ops import EmulatedOps
from spatial_clip_b200 import losses
real_device = torch.device
torch.device = lambda *a, **k: real_device("cpu")
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
losses._set_ops_for_testing(EmulatedOps(round_bf16=True))
import __graft_entry__ as g
g.smoke()
''' % (str(ROOT), str(ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "smoke ok" in r.stdout


def test_nvml_sampler_with_a_stand_in_binding(monkeypatch):
    """bench.NvmlSampler against a fake pynvml: handle lookup, polling thread, median / reasons / power summary."""
    import time
    import types

    sys.path.insert(0, str(ROOT))
    import torch

    import bench

    clocks = iter([1965, 1750, 1740, 1755, 1620] + [1750] * 10000)
    fake = types.SimpleNamespace(
        NVML_CLOCK_SM=1, nvmlInit=lambda: None,
        nvmlDeviceGetHandleByUUID=lambda u: (_ for _ in ()).throw(RuntimeError("no such uuid")),
        nvmlDeviceGetHandleByIndex=lambda i: ("handle", i),
        nvmlDeviceGetMaxClockInfo=lambda h, c: 1965,
        nvmlDeviceGetClockInfo=lambda h, c: next(clocks),
        nvmlDeviceGetCurrentClocksThrottleReasons=lambda h: 0x4,
        nvmlDeviceGetPowerUsage=lambda h: 998000)
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "3,5")
    s = bench.NvmlSampler(torch.device("cpu"), period_s=0.001)
    assert s.start() and s._handle == ("handle", 3)  # UUID lookup failed -> index through CUDA_VISIBLE_DEVICES
    time.sleep(0.05)
    out = s.stop()
    assert out["samples"] >= 5 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    assert out["sm_mhz"] == 1750.0 and out["sm_min_mhz"] == 1620.0 and abs(out["power_w_max"] - 998.0) < 1e-9
