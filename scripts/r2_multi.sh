#!/bin/bash
# Multi-GPU call: bench.py under torchrun at N ranks (NCCL), with the parity block and the per-phase timeline.
#   gpurun --gpus N --timeout 600 -- 'bash scripts/r2_multi.sh N [extra bench flags]'
N=${1:-2}; shift
mkdir -p gpurun_out
t0=$(date +%s)
timeout ${SCL_RUN_TIMEOUT:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 50 --warmup 10 "$@" > gpurun_out/multi_${N}gpu.json 2> gpurun_out/multi_${N}gpu.err
echo "exit $? after $(( $(date +%s) - t0 )) s"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/multi_${N}gpu.err | tail -5
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/multi_${N}gpu.json") if l.startswith("{")][-1]); r = j["roofline"]
    print("N=$N ms/step", round(j["ms_per_step"], 3), "median", round(j["median_ms_per_step"], 3), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]),
          "bwd_ms", round(r["launch_ms"], 3), "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "launches", j["gpu_launches"], "host_ms", round(j["host_enqueue_ms_per_step"], 3), "graphs", j["cuda_graphs"])
    p = j["parity"]; print("parity", {k: p.get(k) for k in ("ok", "ranks", "loss_rel_max_over_ranks", "d_scale_rel_max_over_ranks", "d_image_rows_err_of_max", "d_text_rows_err_of_max", "error")})
    print("timeline", j["timeline_ms"])
except Exception as e:
    print("no json", e)
PY
