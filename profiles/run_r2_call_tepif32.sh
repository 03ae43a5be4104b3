#!/bin/bash
# round 2: fp32-output flavour of the TMA-epilogue GEMM (transformer encoder qkv / out-projection, FIDNetV3)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2q_tests.log 2>&1
tail -3 gpurun_out/r2q_tests.log
for v in 0 1; do
  for sh in enc.qkv enc.o; do
    RALF_GEMM_TEPI_F32=$v timeout 300 python profiles/gemm_bench.py $sh 2>&1 | grep "M=" | sed "s/^/F32=$v /"
  done
done
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2q_tests_model.log 2>&1
tail -3 gpurun_out/r2q_tests_model.log
for rep in 1 2; do
for v in 0 1; do
  RALF_GEMM_TEPI_F32=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2q_bench_f32_$v.$rep.json 2> gpurun_out/r2q_bench_f32_$v.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2q_bench_f32_$v.$rep.json").read().strip().splitlines()[-1])
print("F32=$v rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"], l["phases"]["encode_ms"])
PY
done
done
