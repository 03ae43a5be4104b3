#!/bin/bash
# First GPU call of round 2 (run under gpurun from the repo root, ~25 min of box time: 12 bench lines + the tests; trim the list if the budget is tight):
#   1. the hw_pending tests (written after round 1's GPU budget was spent), then the whole -m gpu suite;
#   2. A/B bench lines for the opt-in variants that are built but unmeasured:
#        baseline | RALF_GEMM_MINB=2 (two GEMM CTAs per SM for the short-K shapes) | RALF_KNN_WAYS=2 (k-NN passes on parallel streams) | RALF_ATTN_TC=2 (8-warp encoder attention) | --decode-ways 2/4 (parallel decode
#        chains inside a batch) | --overlap (two batches in flight) | RALF_GEMM_RESERVE_SMS
# Outputs: gpurun_out/r2_*.{log,json}.  Nothing here is a bench value by itself: copy what is kept into profiles/.
set -x
mkdir -p gpurun_out
RALF_TEST_OPTIN=1 timeout 1500 python -m pytest tests -q -m "gpu and hw_pending" -p no:cacheprovider > gpurun_out/r2_pending_tests.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_gpu_tests.log 2>&1
B="--steps 5 --warmup 3 --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err
RALF_GEMM_MINB=2 timeout 600 python bench.py $B > gpurun_out/r2_bench_minb2.json 2> gpurun_out/r2_bench_minb2.err
timeout 600 python bench.py $B --micro-batch 256 > gpurun_out/r2_bench_mb256.json 2> gpurun_out/r2_bench_mb256.err
timeout 600 python bench.py $B --micro-batch 64 > gpurun_out/r2_bench_mb64.json 2> gpurun_out/r2_bench_mb64.err
RALF_SAMPLE_GRAPH=1 timeout 600 python bench.py $B > gpurun_out/r2_bench_samplegraph.json 2> gpurun_out/r2_bench_samplegraph.err  # moves e2e_model_api only
RALF_KNN_WAYS=2 timeout 600 python bench.py $B > gpurun_out/r2_bench_knnways2.json 2> gpurun_out/r2_bench_knnways2.err
RALF_ATTN_TC=2 timeout 600 python bench.py $B > gpurun_out/r2_bench_attn2.json 2> gpurun_out/r2_bench_attn2.err
timeout 600 python bench.py $B --decode-ways 2 > gpurun_out/r2_bench_ways2.json 2> gpurun_out/r2_bench_ways2.err
timeout 600 python bench.py $B --decode-ways 4 > gpurun_out/r2_bench_ways4.json 2> gpurun_out/r2_bench_ways4.err
timeout 600 python bench.py $B --overlap > gpurun_out/r2_bench_overlap.json 2> gpurun_out/r2_bench_overlap.err
RALF_GEMM_RESERVE_SMS=16 timeout 600 python bench.py $B --overlap > gpurun_out/r2_bench_overlap_reserve16.json 2> gpurun_out/r2_bench_overlap_reserve16.err
RALF_GEMM_MINB=2 timeout 600 python bench.py $B --overlap > gpurun_out/r2_bench_overlap_minb2.json 2> gpurun_out/r2_bench_overlap_minb2.err
tail -3 gpurun_out/r2_pending_tests.log gpurun_out/r2_gpu_tests.log
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
