#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`):
per kernel name the launch count, total / average duration and share of the total.  Pure CSV parsing (runs anywhere).

    python scripts/launch_summary.py gpurun_out/launches.csv [steps]   # steps: divide the sums by this many steps
"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path, newline="") as f:
    for r in csv.reader(f):
        if len(r) >= 15 and r[12] == "gpu__time_duration.sum":
            unit = r[13]
            v = float(r[14].replace(",", ""))
            us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
            rows.append((r[4].split("(")[0][:70], us))
agg = OrderedDict()
for name, us in rows:
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + us)
total = sum(t for _, t in agg.values())
print(f"# total {total:.1f} us over {len(rows)} launches" + (f" = {total / steps:.1f} us per step" if steps != 1 else ""))
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:70s} n={n:3d} sum={t:10.1f} avg={t / n:9.1f} share={t / total:6.3f}")
