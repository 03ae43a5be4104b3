#!/bin/bash
set -x
mkdir -p gpurun_out
for cfg in "1 1 0" "1 1 1" "1 2 1" "3 1 1" "3 2 1"; do
  set -- $cfg
  export RALF_CHAIN_ACC=$1 RALF_CHAIN_PAIR=$2 RALF_CHAIN_PREFETCH=$3
  tag="acc$1_pair$2_pf$3"
  timeout 300 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x -s -k "fused_decode_chain or reference_golden" > gpurun_out/r2g_tests_$tag.log 2>&1
  grep -h "fused-vs-per-op\|passed\|failed\|Error" gpurun_out/r2g_tests_$tag.log | tail -6
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --skip-e2e > gpurun_out/r2g_bench_$tag.json 2> gpurun_out/r2g_bench_$tag.err
  python - "gpurun_out/r2g_bench_$tag.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], "ms/step", d["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
done
