#!/bin/bash
# Round-2 GPU call 1: every hw_pending test (opt-in ones included, no -x so one failure does not hide the rest), the whole
# suite, then the A/B bench lines of the built-but-unmeasured switches.  Outputs: gpurun_out/r2_*.
set -x
mkdir -p gpurun_out
RALF_TEST_OPTIN=1 timeout 1200 python -m pytest tests -q -m "gpu and hw_pending" -p no:cacheprovider -o timeout=300 > gpurun_out/r2_pending_tests.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_gpu_tests.log 2>&1
B="--steps 5 --warmup 3 --no-cpu-baseline"
timeout 400 python bench.py $B > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err
RALF_GEMM_MINB=2 timeout 400 python bench.py $B > gpurun_out/r2_bench_minb2.json 2> gpurun_out/r2_bench_minb2.err
timeout 400 python bench.py $B --micro-batch 256 > gpurun_out/r2_bench_mb256.json 2> gpurun_out/r2_bench_mb256.err
RALF_SAMPLE_GRAPH=1 timeout 400 python bench.py $B > gpurun_out/r2_bench_samplegraph.json 2> gpurun_out/r2_bench_samplegraph.err
RALF_KNN_WAYS=2 timeout 400 python bench.py $B > gpurun_out/r2_bench_knnways2.json 2> gpurun_out/r2_bench_knnways2.err
RALF_ATTN_TC=2 timeout 400 python bench.py $B > gpurun_out/r2_bench_attn2.json 2> gpurun_out/r2_bench_attn2.err
timeout 400 python bench.py $B --decode-ways 2 > gpurun_out/r2_bench_ways2.json 2> gpurun_out/r2_bench_ways2.err
timeout 400 python bench.py $B --decode-ways 4 > gpurun_out/r2_bench_ways4.json 2> gpurun_out/r2_bench_ways4.err
timeout 400 python bench.py $B --overlap > gpurun_out/r2_bench_overlap.json 2> gpurun_out/r2_bench_overlap.err
tail -30 gpurun_out/r2_pending_tests.log; tail -5 gpurun_out/r2_gpu_tests.log
for f in gpurun_out/r2_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "knn frac", d["roofline"]["frac"],
          "model api", d.get("e2e_model_api", {}).get("value"))
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
done
