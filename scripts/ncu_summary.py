#!/usr/bin/env python
"""Summarise an ncu report: key metrics per kernel launch + top stall-sampled SASS lines (runs where ncu is installed)."""
import csv, subprocess, sys, io

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_xu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__cluster_size']
ki = hdr.index('Kernel Name')
for r in data:
    print('----', r[ki][:60])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:70s} {r[i]:>18s} {units[i]}")
