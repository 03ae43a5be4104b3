#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py -q -p no:cacheprovider -x -s > gpurun_out/r2w_tests.log 2>&1
grep -h "greedy-step\|passed\|failed\|Error" gpurun_out/r2w_tests.log | tail -8
B="--steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e"
for lanes in 4 8; do
  RALF_KV16_LANES=$lanes timeout 600 python bench.py $B > gpurun_out/r2w_bench_l$lanes.json 2> gpurun_out/r2w_bench_l$lanes.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2w_bench_l$lanes.json').read().strip().splitlines()[-1]); print('lanes$lanes', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])" || tail -3 gpurun_out/r2w_bench_l$lanes.err
done
