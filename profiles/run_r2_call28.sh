#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_attention_gpu.py -q -p no:cacheprovider -x -s > gpurun_out/r2v_tests.log 2>&1
grep -h "greedy-step\|passed\|failed\|Error" gpurun_out/r2v_tests.log | tail -8
B="--steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e"
for fold in 1 0; do
  RALF_GEMM_FOLD=$fold timeout 600 python bench.py $B > gpurun_out/r2v_bench_fold$fold.json 2> gpurun_out/r2v_bench_fold$fold.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2v_bench_fold$fold.json').read().strip().splitlines()[-1]); print('fold$fold', d['value'], d['ms_per_step'], [ (o['kernel'][:20], o['ms_per_launch'], o['frac']) for o in d['roofline_other']])" || tail -3 gpurun_out/r2v_bench_fold$fold.err
done
for fold in 1 0; do
  echo "fold=$fold"
  RALF_GEMM_FOLD=$fold timeout 300 python profiles/gemm_bench.py 2>&1 | grep "M="
done
