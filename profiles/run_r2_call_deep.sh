#!/bin/bash
# round 2: TMA-epilogue GEMM without residual -> three operand stages + one chunk buffer per warp; K limit of that variant
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2d2_tests.log 2>&1
tail -3 gpurun_out/r2d2_tests.log
RALF_TEPI_KMAX_NORES=1024 timeout 600 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider -x -k "tma_epilogue" > gpurun_out/r2d2_tests_k1024.log 2>&1
tail -2 gpurun_out/r2d2_tests_k1024.log
for cfg in "0 256" "1 256" "1 1024"; do
  set -- $cfg
  for sh in enc.qkv enc.l1 l2.c1 l3.c1 l3.ds l4.ds; do
    RALF_TEPI_DEEP=$1 RALF_TEPI_KMAX_NORES=$2 timeout 300 python profiles/gemm_bench.py $sh 2>&1 | grep "M=" | sed "s/^/DEEP=$1 KNORES=$2 /"
  done
done
for rep in 1 2; do
for cfg in "0 256" "1 256" "1 1024"; do
  set -- $cfg
  RALF_TEPI_DEEP=$1 RALF_TEPI_KMAX_NORES=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2d2_bench_$1_$2.$rep.json 2> gpurun_out/r2d2_bench_$1_$2.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2d2_bench_$1_$2.$rep.json").read().strip().splitlines()[-1])
print("DEEP=$1 KNORES=$2 rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"], l["phases"]["encode_ms"])
PY
done
done
