#!/bin/bash
# round 2: TMA-epilogue GEMM for the bottleneck-tail shapes -- parity, micro-benchmark A/B, bench step A/B.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2t2_tests.log 2>&1
tail -5 gpurun_out/r2t2_tests.log
for v in 0 1; do
  for sh in l1. l2.c3 l3.c3; do
    RALF_GEMM_TEPI=$v timeout 300 python profiles/gemm_bench.py $sh 2>&1 | grep -v "^\[" | sed "s/^/TEPI=$v /"
  done
done
timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2t2_tests_model.log 2>&1
tail -3 gpurun_out/r2t2_tests_model.log
for rep in 1 2; do
for v in 0 1; do
  RALF_GEMM_TEPI=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2t2_bench_tepi$v.$rep.json 2> gpurun_out/r2t2_bench_tepi$v.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2t2_bench_tepi$v.$rep.json").read().strip().splitlines()[-1])
print("TEPI=$v rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"])
PY
done
done
