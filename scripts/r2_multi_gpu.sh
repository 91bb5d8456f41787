#!/bin/bash
# Round-2 multi-GPU call (default 2 GPUs): what has not run on more than one B200 yet.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/r2_multi_gpu.sh 2'
# 1. multi-rank parity (ranks share GPU 0, gloo transport) for the default route and the two W > 1 developer knobs
# 2. bench.py under torchrun (NCCL), one-call-per-phase route timed (--kernel-events after), for each knob and all together
N=${1:-2}
mkdir -p gpurun_out
for knobs in "" "SCL_OVERLAP_GATHER=1" "SCL_BWD_MN=1" "SCL_BWD_STREAMS=1"; do
  tag=$(echo "${knobs:-default}" | tr '= ' '__')
  echo "=== multi-rank parity [$tag]"
  env $knobs timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k multi_rank \
      > gpurun_out/r2_mr_$tag.log 2>&1
  echo "exit $?"; tail -2 gpurun_out/r2_mr_$tag.log
done
port=29500
for knobs in "" "SCL_OVERLAP_GATHER=1" "SCL_BWD_MN=1" "SCL_BWD_STREAMS=1" "SCL_BWD_CHUNKS=2" \
             "SCL_OVERLAP_GATHER=1 SCL_BWD_MN=1 SCL_BWD_STREAMS=1 SCL_BWD_TUNE=3"; do
  tag=$(echo "${knobs:-default}" | tr '= ' '__')
  port=$((port + 1))
  echo "=== bench N=$N [$tag]"
  env $knobs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --kernel-events after > gpurun_out/r2_bench_n${N}_$tag.json 2> gpurun_out/r2_bench_n${N}_$tag.err
  echo "exit $?"
  python - <<PY
import json
try:
    j = [json.loads(l) for l in open("gpurun_out/r2_bench_n${N}_$tag.json") if l.startswith("{")][-1]; r = j["roofline"]
    print("ms/step", round(j["ms_per_step"], 3), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]),
          "bwd_ms", round(r["launch_ms"], 3), "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "loss", j["loss"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/r2_bench_n${N}_$tag.err").read()[-1500:])
PY
done
