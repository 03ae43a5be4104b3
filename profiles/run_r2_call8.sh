#!/bin/bash
set -x
mkdir -p gpurun_out
for cfg in "1 1 0" "1 1 1" "1 2 1" "3 1 1" "3 2 1" "3 2 0"; do
  set -- $cfg
  export RALF_CHAIN_ACC=$1 RALF_CHAIN_PAIR=$2 RALF_CHAIN_PREFETCH=$3
  tag="acc$1_pair$2_pf$3"
  timeout 300 python profiles/chain_bench.py > gpurun_out/r2h_chain_$tag.json 2> gpurun_out/r2h_chain_$tag.err
  cat gpurun_out/r2h_chain_$tag.json
done
export RALF_CHAIN_ACC=3 RALF_CHAIN_PAIR=2 RALF_CHAIN_PREFETCH=1
timeout 300 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x -s -k "fused_decode_chain or reference_golden" > gpurun_out/r2h_tests.log 2>&1
grep -h "fused-vs-per-op\|passed\|failed\|Error" gpurun_out/r2h_tests.log | tail -6
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
