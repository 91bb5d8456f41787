#!/bin/bash
# Run the GPU parity tests layer by layer in separate processes (a trapped kernel poisons its CUDA context),
# logs into gpurun_out/.  Usage: scripts/gpu_layers.sh [extra pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() {
  name=$1; shift
  echo "=== $name"
  timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?"; tail -n 25 gpurun_out/$name.log
}
run l1_similarity -k similarity
run l2_rowstats -k row_statistics
run l3_positives -k positive_lists
run l4_bwd -k bwd_rows
run l5_modules -k "modules_match or bf16_inputs or no_grad"
run l6_mid -k "mid_size or full_size"
run l7_multirank -k multi_rank
