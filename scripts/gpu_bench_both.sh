#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  echo "=== bench SCL_VARIANT=$v"
  SCL_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  echo "exit $?"; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_v$v.json"))
    r=j["roofline"]
    print("ms/step",round(j["ms_per_step"],3),"pairs/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"bwd_ms",round(r["launch_ms"],3),"fwd_ms",round(r["fwd_rowstats_launch_ms"],3),"loss",j["loss"],"clk",j["clocks"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_v$v.err").read()[-1500:])
PY
done
