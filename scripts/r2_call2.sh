#!/bin/bash
# Round-2 second GPU call (1 GPU): full parity suite on the ABI v7 library, the bwd lab, one bench run of each arm.
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c2_gpu_tests.log 2>&1; echo "exit $?"; tail -5 gpurun_out/c2_gpu_tests.log
grep -E "baseline-size parity|^FAILED|^ERROR" gpurun_out/c2_gpu_tests.log | cut -c1-400
echo "=== bwd lab"; timeout 900 python tools/bwd_lab.py run 2>&1 | tee gpurun_out/c2_bwd_lab.txt
echo "=== bench (ours)"; timeout 600 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "exit $?"; tail -c 300 gpurun_out/c2_bench.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/c2_bench.json")); r = j["roofline"]
    print("ms/step", round(j["ms_per_step"], 3), "median", round(j["median_ms_per_step"], 3), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]),
          "bwd_ms", round(r["launch_ms"], 3), "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "launches", j["gpu_launches"], "loss", j["loss"])
    print("parity", j["parity"]); print("cpu", j["cpu_baseline"]); print("clocks", j["clocks"])
except Exception as e:
    print("no json", e)
PY
echo "=== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/c2_bench_ref.json 2> gpurun_out/c2_bench_ref.err; echo "exit $?"; cat gpurun_out/c2_bench_ref.json | cut -c1-1500; tail -c 300 gpurun_out/c2_bench_ref.err
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
nproc; free -g | head -2
