#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c13_gpu_tests.log 2>&1; echo "exit $?"; tail -3 gpurun_out/c13_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/c13_gpu_tests.log | head -10 | cut -c1-300
echo "=== bwd lab"; timeout 600 python tools/bwd_lab.py run 2>&1 | tee gpurun_out/c13_bwd_lab.txt | cut -c1-130
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c13_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1; python scripts/launch_summary.py gpurun_out/c13_launches.csv 2>/dev/null | head -8
