#!/usr/bin/env python
"""Top stall-sampled SASS instructions of one kernel in an ncu report."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {}
for r in data:
    for i in stall:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][isamp] or 0))[:n]:
    st = {hdr[i][6:]: int(r[i] or 0) for i in stall if int(r[i] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{idx:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>9s} {r[isrc][:64]:64s} {st}")
