#!/bin/bash
# Round-2 third GPU call (1 GPU): 7-slot ring parity + lab, launch list of one bench step, reference arm check.
mkdir -p gpurun_out
echo "=== gpu tests (bwd + modules + baseline sizes + wide + fp32)"
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c5_gpu_tests.log 2>&1; echo "exit $?"; tail -4 gpurun_out/c5_gpu_tests.log
grep -E "baseline-size parity|^FAILED|^ERROR" gpurun_out/c5_gpu_tests.log | cut -c1-400
echo "=== bwd lab"; timeout 900 python tools/bwd_lab.py run 2>&1 | tee gpurun_out/c5_bwd_lab.txt
echo "=== bench (ours)"; timeout 600 python bench.py > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "exit $?"; tail -c 300 gpurun_out/c5_bench.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/c5_bench.json")); r = j["roofline"]
    print("ms/step", round(j["ms_per_step"], 3), "median", round(j["median_ms_per_step"], 3), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]),
          "bwd_ms", round(r["launch_ms"], 3), "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "launches", j["gpu_launches"], "loss", j["loss"])
    print("parity ok", j["parity"].get("ok"), "cpu kind", j["cpu_baseline"]["kind"], j["cpu_baseline"]["value"]); print("clocks", j["clocks"])
except Exception as e:
    print("no json", e)
PY
echo "=== launch list (ncu, 3 steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c5_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/c5_ncu.log 2>&1; echo "exit $?"
python scripts/launch_summary.py gpurun_out/c5_launches.csv | head -40
ls -la oracle/_ref
