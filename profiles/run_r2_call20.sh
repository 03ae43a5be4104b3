#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py tests/test_pipeline_gpu.py -q -p no:cacheprovider -x -s > gpurun_out/r2p_tests.log 2>&1
grep -h "greedy-step\|passed\|failed\|Error" gpurun_out/r2p_tests.log | tail -8
B="--steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 python bench.py $B > gpurun_out/r2p_bench_v2.json 2> gpurun_out/r2p_bench_v2.err
RALF_KV16_KERNEL=1 timeout 600 python bench.py $B --skip-e2e > gpurun_out/r2p_bench_v1.json 2> gpurun_out/r2p_bench_v1.err
for f in gpurun_out/r2p_bench_v*.json; do python -c "
import json,sys
d = json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'])"; done
