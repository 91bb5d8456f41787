#!/bin/bash
# Round-2 first GPU call (1 GPU, ~30-35 min; every === block is independent): parity of the default kernels, then the opt-in SCL_BWD_TUNE
# instantiations (parity first, then speed), then the per-role wait-cycle counters -- everything lands in gpurun_out/.
#   gpurun --timeout 2700 -- 'bash scripts/r2_first_call.sh'
mkdir -p gpurun_out
echo "=== default parity"; timeout 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_gpu_tests.log 2>&1; echo "exit $?"; tail -3 gpurun_out/r2_gpu_tests.log
for t in 1 2 3 7; do
  echo "=== SCL_BWD_TUNE=$t parity (bwd kernels, modules, mid/full size)"
  SCL_BWD_TUNE=$t timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
      -k "bwd_rows or modules_match or mid_size or full_size or wide" > gpurun_out/r2_tune${t}_tests.log 2>&1
  echo "exit $?"; tail -2 gpurun_out/r2_tune${t}_tests.log
done
echo "=== SCL_BWD_MN=1 parity (no transposed copies: MN-major B operand in the gradient GEMM)"
SCL_BWD_MN=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
    -k "cta2 and (bwd_rows or modules_match or mid_size or full_size or wide or multi_rank)" > gpurun_out/r2_mn_tests.log 2>&1
echo "exit $?"; tail -2 gpurun_out/r2_mn_tests.log
for mn in 0 1; do
  echo "=== bench SCL_BWD_MN=1 SCL_BWD_TUNE=$((mn * 3))"
  SCL_BWD_MN=1 SCL_BWD_TUNE=$((mn * 3)) timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench_mn$mn.json 2> gpurun_out/r2_bench_mn$mn.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r2_bench_mn$mn.json")); r = j["roofline"]
    print("ms/step", round(j["ms_per_step"], 3), "pairs/s", round(j["value"]), "bwd_ms", round(r["launch_ms"], 3), "loss", j["loss"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/r2_bench_mn$mn.err").read()[-1500:])
PY
done
for t in 0 1 2 3 7; do
  echo "=== bench SCL_BWD_TUNE=$t"
  SCL_BWD_TUNE=$t timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench_tune$t.json 2> gpurun_out/r2_bench_tune$t.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r2_bench_tune$t.json")); r = j["roofline"]
    print("ms/step", round(j["ms_per_step"], 3), "pairs/s", round(j["value"]), "bwd_ms", round(r["launch_ms"], 3),
          "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "loss", j["loss"], "clk", j["clocks"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/r2_bench_tune$t.err").read()[-1500:])
PY
done
echo "=== SCL_AUX_V2=1 (row_finalize with 16-byte loads): parity, then the kernel's time in the launch list"
SCL_AUX_V2=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
    -k "row_statistics or modules_match or mid_size or wide" > gpurun_out/r2_auxv2_tests.log 2>&1
echo "exit $?"; tail -2 gpurun_out/r2_auxv2_tests.log
SCL_AUX_V2=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:row_finalize -c 8 --csv \
    --log-file gpurun_out/r2_auxv2_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
grep row_finalize gpurun_out/r2_auxv2_launches.csv | tail -4 | cut -c1-200
echo "=== SCL_BWD_STREAMS=1 (gene-side backward chain on a second stream): parity, then speed with the one-call-per-phase route timed"
SCL_BWD_STREAMS=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
    -k "modules_match or bf16_inputs or mid_size or full_size or multi_rank" > gpurun_out/r2_streams_tests.log 2>&1
echo "exit $?"; tail -2 gpurun_out/r2_streams_tests.log
for st in 0 1; do
  SCL_BWD_STREAMS=$st timeout 300 python bench.py --steps 8 --warmup 3 --kernel-events after > gpurun_out/r2_bench_streams$st.json 2> gpurun_out/r2_bench_streams$st.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r2_bench_streams$st.json"))
    print("SCL_BWD_STREAMS=$st ms/step", round(j["ms_per_step"], 3), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]), "loss", j["loss"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/r2_bench_streams$st.err").read()[-1500:])
PY
done
echo "=== everything at once (TUNE=7, MN, AUX_V2, STREAMS), one-call-per-phase route timed"
SCL_BWD_TUNE=7 SCL_BWD_MN=1 SCL_AUX_V2=1 SCL_BWD_STREAMS=1 timeout 300 python bench.py --steps 8 --warmup 3 --kernel-events after > gpurun_out/r2_bench_all.json 2> gpurun_out/r2_bench_all.err
echo "exit $?"; tail -c 400 gpurun_out/r2_bench_all.json | head -c 400; echo
echo "=== fixture with different image / text tile ids (added after the last B200 run)"
SCL_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k asym > gpurun_out/r2_asym_tests.log 2>&1
echo "exit $?"; tail -2 gpurun_out/r2_asym_tests.log
echo "=== in-pass retrieval ranks"
SCL_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_retrieval_ranks.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_ranks_tests.log 2>&1
echo "exit $?"; tail -2 gpurun_out/r2_ranks_tests.log
echo "=== data-side soft targets (SpatialLossFromColumns, phases = 6)"
SCL_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_positive_columns.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_columns_tests.log 2>&1
echo "exit $?"; tail -2 gpurun_out/r2_columns_tests.log
echo "=== fp32-accurate mode (precision=fp32): parity, then speed"
SCL_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_fp32_mode.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_fp32_tests.log 2>&1
echo "exit $?"; grep -E "passed|failed" gpurun_out/r2_fp32_tests.log | tail -2; grep -E "^(FAILED|ERROR|E  +Assert|E  +assert|E  +.*Error)" gpurun_out/r2_fp32_tests.log | cut -c1-240 | head -20
timeout 300 python bench.py --steps 5 --warmup 3 --precision fp32 > gpurun_out/r2_bench_fp32.json 2> gpurun_out/r2_bench_fp32.err; echo "exit $?"; tail -c 600 gpurun_out/r2_bench_fp32.json; tail -3 gpurun_out/r2_bench_fp32.err
echo "=== experimental widths 640 / 1152 / 1280 / 1536"
SCL_EXPERIMENTAL_SHAPES=1 timeout 300 python -m pytest tests/test_gpu_experimental_shapes.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_shapes_tests.log 2>&1
echo "exit $?"; grep -E "passed|failed" gpurun_out/r2_shapes_tests.log | tail -2
echo "=== DSMEM hand-over probe (feasibility of the split-role backward)"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spatial_clip_b200/csrc tools/dsmem_probe.cu -o /tmp/dsmem_probe 2>/dev/null || cp tools/dsmem_probe.bin /tmp/dsmem_probe
timeout 60 /tmp/dsmem_probe > gpurun_out/r2_dsmem_probe.txt 2>&1; cat gpurun_out/r2_dsmem_probe.txt
echo "=== may tcgen05.mma consume an operand another SM stored with st.shared::cluster, and with which fences?"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spatial_clip_b200/csrc tools/dsmem_mma_probe.cu -o /tmp/dsmem_mma_probe 2>/dev/null || cp tools/dsmem_mma_probe.bin /tmp/dsmem_mma_probe
timeout 120 /tmp/dsmem_mma_probe > gpurun_out/r2_dsmem_mma_probe.txt 2>&1; cat gpurun_out/r2_dsmem_mma_probe.txt
echo "=== wait-cycle counters (default kernels)"
timeout 300 python tools/kernel_timing.py > gpurun_out/r2_kernel_timing.txt 2>&1; tail -40 gpurun_out/r2_kernel_timing.txt
