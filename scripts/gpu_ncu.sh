#!/bin/bash
# ncu captures of the default (CTA-pair) kernels; bench must have exited 0 without ncu first.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bwd_rows|fwd_rowstats" -s 4 -c 4 -o gpurun_out/prof_r1_v2 -f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"; tail -2 gpurun_out/ncu_full.log
