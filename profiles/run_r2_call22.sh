#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_attention_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2r_tests.log 2>&1
tail -4 gpurun_out/r2r_tests.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e"
timeout 600 python bench.py $B > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python -c "
import json,sys
d = json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
timeout 300 python profiles/gemm_bench.py > gpurun_out/r2r_gemm_bench.json 2> gpurun_out/r2r_gemm_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_gemm_bench.json'))
for r in d:
    if isinstance(r,dict): print({k:(round(v,1) if isinstance(v,float) else v) for k,v in r.items()})
PY
