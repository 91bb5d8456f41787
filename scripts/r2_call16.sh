#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c16_gpu_tests.log 2>&1; echo "exit $?"; tail -3 gpurun_out/c16_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/c16_gpu_tests.log | head -10 | cut -c1-300
echo "=== bench"; timeout 400 python bench.py --timeline > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; echo "exit $?"; tail -c 300 gpurun_out/c16_bench.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/c16_bench.json")); r = j["roofline"]
    print("ms/step", round(j["ms_per_step"], 3), "median", round(j["median_ms_per_step"], 3), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]),
          "bwd_ms", round(r["launch_ms"], 3), "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "host_ms", round(j["host_enqueue_ms_per_step"], 3))
    print("parity", j["parity"].get("ok")); print("timeline", j["timeline_ms"])
except Exception as e:
    print("no json", e)
PY
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c16_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1; python scripts/launch_summary.py gpurun_out/c16_launches.csv 2>/dev/null | head -22
