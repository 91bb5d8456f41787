#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c12_gpu_tests.log 2>&1; echo "exit $?"; tail -3 gpurun_out/c12_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/c12_gpu_tests.log | head -10 | cut -c1-300
echo "=== bench bf16"; timeout 400 python bench.py > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; echo "exit $?"
echo "=== bench fp32"; timeout 400 python bench.py --precision fp32 --steps 20 --warmup 5 > gpurun_out/c12_bench_fp32.json 2> gpurun_out/c12_bench_fp32.err; echo "exit $?"; tail -c 300 gpurun_out/c12_bench_fp32.err
python - <<'PY'
import json
for f in ("c12_bench", "c12_bench_fp32"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json")); r = j["roofline"]
        print(f, "ms/step", round(j["ms_per_step"], 3), "median", round(j["median_ms_per_step"], 3), "pairs/s", round(j["value"]), "e2e", round(j["e2e"]["value"]),
              "bwd_ms", round(r["launch_ms"], 3), "fwd_ms", round(r["fwd_rowstats_launch_ms"], 3), "host_ms", round(j["host_enqueue_ms_per_step"], 3))
        p = j["parity"]; print("   parity", {k: p.get(k) for k in ("ok", "loss_rel_max_over_ranks", "d_scale_rel_max_over_ranks", "d_image_rows_err_of_max", "d_text_rows_err_of_max", "error")})
    except Exception as e:
        print(f, "no json", e)
PY
echo "=== sweep (configs[4])"; timeout 1200 python tools/sweep.py > gpurun_out/c12_sweep.jsonl 2> gpurun_out/c12_sweep.err; echo "exit $?"; tail -3 gpurun_out/c12_sweep.err
python - <<'PY'
import json
for l in open("gpurun_out/c12_sweep.jsonl"):
    j = json.loads(l); p = j.get("parity", {})
    print(j["n"], j["d"], j["precision"], "ms", round(j["ms_per_step"], 2), "Mpairs/s", round(j["pairs_per_s"] / 1e6, 3), "alg", round(j["algorithmic_frac"], 3), "exec", round(j["executed_frac"], 3),
          "euler", f'{j["euler_residual"]:.1e}', "loss_rel", p.get("loss_rel"), "grad", p.get("d_image_rows_err_of_max"))
PY
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c12_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1; python scripts/launch_summary.py gpurun_out/c12_launches.csv | head -16
