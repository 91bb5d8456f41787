#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c9_gpu_tests.log 2>&1; echo "exit $?"; tail -4 gpurun_out/c9_gpu_tests.log
grep -E "^FAILED|^ERROR|Error" gpurun_out/c9_gpu_tests.log | head -10 | cut -c1-300
echo "=== bench (ours)"; timeout 600 python bench.py > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; echo "exit $?"; tail -c 400 gpurun_out/c9_bench.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/c9_bench.json")); r = j["roofline"]
    print("ms/step", round(j["ms_per_step"], 3), "median", round(j["median_ms_per_step"], 3), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]),
          "bwd_ms", round(r["launch_ms"], 3), "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "launches", j["gpu_launches"], "host_ms", j["host_enqueue_ms_per_step"])
    print("parity", j["parity"].get("ok"), j["parity"].get("loss_rel_max_over_ranks"), j["parity"].get("d_image_rows_err_of_max"))
except Exception as e:
    print("no json", e)
PY
