#!/bin/bash
# round 2: strided implicit convolutions + TMA-epilogue coverage (K limit, K = 64 ring depth)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2t3_tests.log 2>&1
tail -5 gpurun_out/r2t3_tests.log
for cfg in "256 14" "256 23" "1024 14"; do
  set -- $cfg
  for sh in l1.c3 l1.ds l2.c1 l3.c1 l3.ds l4.c3 l4.ds; do
    RALF_TEPI_KMAX=$1 RALF_TEPI_K64=$2 timeout 300 python profiles/gemm_bench.py $sh 2>&1 | grep "M=" | sed "s/^/KMAX=$1 K64=$2 /"
  done
done
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2t3_tests_model.log 2>&1
tail -3 gpurun_out/r2t3_tests_model.log
for rep in 1 2; do
for cfg in "0 256" "1 256" "1 1024"; do
  set -- $cfg
  RALF_STRIDED_CONV=$1 RALF_TEPI_KMAX=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2t3_bench_s$1_k$2.$rep.json 2> gpurun_out/r2t3_bench_s$1_k$2.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2t3_bench_s$1_k$2.$rep.json").read().strip().splitlines()[-1])
print("STRIDED=$1 KMAX=$2 rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"])
PY
done
done
