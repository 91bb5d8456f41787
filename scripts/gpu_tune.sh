#!/bin/bash
# developer experiment: SCL_TUNE knobs (bit0 lane-0 polling, bit1 single G buffer + 7 stages)
for t in 0 1 2 3; do
  echo "=== SCL_TUNE=$t"
  SCL_TUNE=$t timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_t$t.json 2>gpurun_out/bench_t$t.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_t$t.json")); r=j["roofline"]
    print("ms/step",round(j["ms_per_step"],3),"bwd_ms",round(r["launch_ms"],3),"fwd_ms",round(r["fwd_rowstats_launch_ms"],3),"loss",j["loss"])
except Exception as e:
    print("fail",e); print(open("gpurun_out/bench_t$t.err").read()[-800:])
PY
done
