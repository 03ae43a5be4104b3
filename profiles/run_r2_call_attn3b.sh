#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py tests/test_model_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2a4_tests.log 2>&1
tail -5 gpurun_out/r2a4_tests.log
RALF_ATTN_TC=3 timeout 300 python profiles/attn_bench.py 2>&1 | tail -1
for sh in l1.c3 l2.c3 enc.l1; do timeout 300 python profiles/gemm_bench.py $sh 2>&1 | grep "M="; done
for rep in 1 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a4_bench.$rep.json 2> gpurun_out/r2a4_bench.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2a4_bench.$rep.json").read().strip().splitlines()[-1])
print("rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"])
PY
done
