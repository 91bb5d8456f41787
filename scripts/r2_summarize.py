#!/usr/bin/env python
"""Turn what scripts/r2_first_call.sh and scripts/r2_multi_gpu.sh left in gpurun_out/ into one table:
parity verdict and step time per developer knob, so that the defaults can be flipped (or the knob deleted)
without re-reading every log.  Reads only; run it here after a gpurun call has merged its outputs back.

    python scripts/r2_summarize.py [gpurun_out]
"""
import json
import re
import sys
from pathlib import Path

out = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")


def verdict(log: Path) -> str:
    if not log.exists():
        return "not run"
    text = log.read_text(errors="replace")
    m = re.findall(r"(\d+) passed", text)
    f = re.findall(r"(\d+) failed", text)
    e = re.findall(r"(\d+) error", text)
    s = re.findall(r"(\d+) skipped", text)
    if not m and not f and not e:
        tail = text.strip().splitlines()[-1][:100] if text.strip() else "empty log"
        return f"no pytest summary ({tail})"
    parts = []
    if m:
        parts.append(f"{m[-1]} passed")
    if f:
        parts.append(f"{f[-1]} FAILED")
    if e:
        parts.append(f"{e[-1]} ERRORS")
    if s:
        parts.append(f"{s[-1]} skipped")
    return ", ".join(parts)


def first_failures(log: Path, n=3):
    if not log.exists():
        return []
    return [ln.strip()[:160] for ln in log.read_text(errors="replace").splitlines()
            if ln.startswith(("FAILED", "ERROR")) or re.match(r"E\s+(Assert|assert|\w+Error)", ln)][:n]


def bench(path: Path):
    if not path.exists():
        return None
    for ln in reversed(path.read_text(errors="replace").splitlines()):
        if ln.startswith("{"):
            try:
                return json.loads(ln)
            except json.JSONDecodeError:
                continue
    return None


def bench_row(name: str, j) -> str:
    if j is None:
        return f"  {name:44s} no JSON line"
    r = j.get("roofline", {})
    e2e = j.get("e2e", {}).get("value")
    return (f"  {name:44s} {j['ms_per_step']:7.3f} ms/step  {j['value'] / 1e6:6.2f} M pairs/s  "
            f"e2e {e2e / 1e6 if e2e else float('nan'):6.2f} M  bwd {r.get('launch_ms', float('nan')):6.3f} ms  "
            f"fwd {r.get('fwd_rowstats_launch_ms') or float('nan'):6.3f} ms  frac {r.get('frac', float('nan')):.3f}  "
            f"loss {j.get('loss', float('nan')):.6f}")


print("== parity")
tests = [("default (must be green)", "r2_gpu_tests.log")]
tests += [(f"SCL_BWD_TUNE={t}", f"r2_tune{t}_tests.log") for t in (1, 2, 3, 7)]
tests += [("SCL_BWD_MN=1", "r2_mn_tests.log"), ("SCL_AUX_V2=1", "r2_auxv2_tests.log"),
          ("SCL_BWD_STREAMS=1", "r2_streams_tests.log"), ("asymmetric tile ids", "r2_asym_tests.log"),
          ("in-pass retrieval ranks", "r2_ranks_tests.log"), ("SpatialLossFromColumns", "r2_columns_tests.log"),
          ("precision=fp32", "r2_fp32_tests.log"), ("widths 640/1152/1280/1536", "r2_shapes_tests.log")]
tests += [(f"multi-rank [{p.stem[6:]}]", p.name) for p in sorted(out.glob("r2_mr_*.log"))]
for name, f in tests:
    print(f"  {name:44s} {verdict(out / f)}")
    for ln in first_failures(out / f):
        print(f"      {ln}")

print("== bench, 1 GPU (ms/step is timed with the kernel events inside the steps unless the name says 'streams')")
rows = [(f"SCL_BWD_TUNE={t}", f"r2_bench_tune{t}.json") for t in (0, 1, 2, 3, 7)]
rows += [("SCL_BWD_MN=1", "r2_bench_mn0.json"), ("SCL_BWD_MN=1 SCL_BWD_TUNE=3", "r2_bench_mn1.json"),
         ("streams off (--kernel-events after)", "r2_bench_streams0.json"),
         ("SCL_BWD_STREAMS=1 (--kernel-events after)", "r2_bench_streams1.json"),
         ("TUNE=7 MN AUX_V2 STREAMS (--kernel-events after)", "r2_bench_all.json"),
         ("precision=fp32", "r2_bench_fp32.json")]
base = None
for name, f in rows:
    j = bench(out / f)
    if name == "SCL_BWD_TUNE=0" and j:
        base = j["ms_per_step"]
    print(bench_row(name, j) + (f"  ({base / j['ms_per_step']:.3f}x of TUNE=0)" if base and j else ""))
multi = sorted(out.glob("r2_bench_n*_*.json"))
if multi:
    print("== bench, several GPUs (--kernel-events after)")
    for p in multi:
        print(bench_row(p.stem[len("r2_bench_"):], bench(p)))
for name in ("r2_dsmem_probe.txt", "r2_dsmem_mma_probe.txt", "r2_kernel_timing.txt"):
    p = out / name
    if p.exists():
        print(f"== {name}")
        print("\n".join("  " + ln for ln in p.read_text(errors="replace").strip().splitlines()[-24:]))
print("\nDecision rules (profiles/ROUND2_PLAN.md): a knob whose parity is green and whose step is faster becomes the "
      "default and its switch is deleted; one that fails parity is fixed or deleted, not kept.")
