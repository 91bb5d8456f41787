#!/bin/bash
# Multi-GPU call: bench.py under torchrun at N ranks (NCCL), with the parity block and the per-phase timeline.
#   gpurun --gpus N --timeout 900 -- 'bash scripts/r2_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
for tag in a; do
  timeout ${SCL_RUN_TIMEOUT:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 30 --warmup 5 --timeline > gpurun_out/multi_${N}gpu_$tag.json 2> gpurun_out/multi_${N}gpu_$tag.err
  echo "exit $?"; tail -c 600 gpurun_out/multi_${N}gpu_$tag.err | tail -5
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/multi_${N}gpu_$tag.json") if l.startswith("{")][-1]); r = j["roofline"]
    print("N=$N ms/step", round(j["ms_per_step"], 3), "median", round(j["median_ms_per_step"], 3), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]),
          "bwd_ms", round(r["launch_ms"], 3), "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "launches", j["gpu_launches"], "host_ms", round(j["host_enqueue_ms_per_step"], 3))
    p = j["parity"]; print("parity", {k: p.get(k) for k in ("ok", "ranks", "loss_rel_max_over_ranks", "d_scale_rel_max_over_ranks", "d_image_rows_err_of_max", "d_text_rows_err_of_max", "error")})
    print("timeline", j["timeline_ms"])
except Exception as e:
    print("no json", e)
PY
done
