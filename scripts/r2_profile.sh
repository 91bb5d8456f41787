#!/bin/bash
# ncu --set full capture of the two tensor-core kernels inside a bench step (1 GPU); summaries are made offline.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bwd_rows_pair|fwd_rowstats_pair" -s 8 -c 4 \
    -o gpurun_out/prof_r2 -f python bench.py --steps 2 --warmup 3 > gpurun_out/prof_r2.log 2>&1
echo "exit $?"; ls -la gpurun_out/prof_r2.ncu-rep
