#!/bin/bash
# smoke + bench + ncu captures (1 GPU).  Results under gpurun_out/.
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench"; timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 | tail -c 1200
echo "=== ncu launch list"
timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_list.log
echo "=== ncu full (bwd_rows, fwd_rowstats)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bwd_rows|fwd_rowstats" -s 4 -c 4 -o gpurun_out/prof_r2 -f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
