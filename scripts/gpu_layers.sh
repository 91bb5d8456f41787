#!/bin/bash
# GPU parity tests layer by layer, each in its own process (a trapped kernel poisons its CUDA context).
# Usage: scripts/gpu_layers.sh [variant filter, e.g. cta1 | cta2 | "cta1 or cta2"]
mkdir -p gpurun_out
V=${1:-"cta1 or cta2"}
run() {
  name=$1; shift
  echo "=== $name [$V]"
  timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider -s -k "($1) and ($V)" > gpurun_out/$name.log 2>&1
  echo "exit $?"; grep -E "passed|failed" gpurun_out/$name.log | tail -2; grep -E "^(FAILED|ERROR|E  +Assert|E  +assert|E  +.*Error)" gpurun_out/$name.log | cut -c1-260 | head -14
}
run l1_similarity "similarity"
run l2_rowstats "row_statistics"
run l3_positives "positive_lists"
run l4_bwd "bwd_rows"
run l5_modules "modules_match or bf16_inputs or no_grad"
run l6_mid "mid_size or full_size"
run l7_multirank "multi_rank"
