#!/bin/bash
# What the driver runs at round end, in one 1-GPU call: -m gpu tests, smoke(), both bench arms.
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/final_gpu_tests.log 2>&1; echo "exit $?"; tail -2 gpurun_out/final_gpu_tests.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "exit $?"; cut -c1-200 gpurun_out/final_bench_ref.json
echo "=== bench (ours)"; timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "exit $?"; tail -c 300 gpurun_out/final_bench.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/final_bench.json")); r = j["roofline"]
    print("ms/step", round(j["ms_per_step"], 3), "median", round(j["median_ms_per_step"], 3), j["step_ms_spread"], "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["e2e"]["runs_ms"],
          "bwd_ms", round(r["launch_ms"], 3), "frac", round(r["frac"], 3), "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "host_ms", round(j["host_enqueue_ms_per_step"], 3))
    print("parity", j["parity"].get("ok"), "cpu", j["cpu_baseline"]["kind"], round(j["cpu_baseline"]["value"]), "clocks", j["clocks"])
except Exception as e:
    print("no json", e)
PY
