#!/bin/bash
# round 2: half-TMEM encoder attention (two computing CTAs per SM) -- parity, micro-benchmark, bench step A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2a3_tests.log 2>&1
tail -5 gpurun_out/r2a3_tests.log
for v in 1 3; do
  RALF_ATTN_TC=$v timeout 300 python profiles/attn_bench.py 2>&1 | tail -4 | sed "s/^/ATTN_TC=$v /"
done
timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2a3_tests_model.log 2>&1
tail -3 gpurun_out/r2a3_tests_model.log
for rep in 1 2; do
for v in 1 3; do
  RALF_ATTN_TC=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a3_bench_tc$v.$rep.json 2> gpurun_out/r2a3_bench_tc$v.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2a3_bench_tc$v.$rep.json").read().strip().splitlines()[-1])
print("ATTN_TC=$v rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"])
PY
done
done
