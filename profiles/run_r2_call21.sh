#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider -x -k "fused_layernorm" > gpurun_out/r2q_tests.log 2>&1
tail -4 gpurun_out/r2q_tests.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e"
timeout 600 python bench.py $B > gpurun_out/r2q_bench_base.json 2> gpurun_out/r2q_bench_base.err
RALF_FUSE_LN=1 timeout 600 python bench.py $B > gpurun_out/r2q_bench_fuseln.json 2> gpurun_out/r2q_bench_fuseln.err
tail -3 gpurun_out/r2q_bench_fuseln.err
for f in gpurun_out/r2q_bench_*.json; do python -c "
import json,sys
d = json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['gpu_launches'])"; done
RALF_FUSE_LN=1 timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x -k "reference_golden" > gpurun_out/r2q_tests_fuseln.log 2>&1
tail -3 gpurun_out/r2q_tests_fuseln.log
