#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py tests/test_pipeline_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2m_tests.log 2>&1
tail -6 gpurun_out/r2m_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
RALF_ATTN_FEWKEYS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e > gpurun_out/r2m_bench_fk1.json 2> gpurun_out/r2m_bench_fk1.err
for f in gpurun_out/r2m_bench*.json; do python -c "
import json,sys
d = json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'])"; done
