#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  echo "=== $name"
  timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider -s "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?"; grep -E "passed|failed" gpurun_out/$name.log | tail -2; grep -E "^(FAILED|E  +Assert|E  +assert)" gpurun_out/$name.log | cut -c1-300 | head -12
}
run l5_modules -k "modules_match or bf16_inputs or no_grad"
run l6_mid -k "mid_size or full_size"
run l7_multirank -k multi_rank
