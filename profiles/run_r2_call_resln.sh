#!/bin/bash
# round 2: LayerNorm in the epilogue of the decode loop's residual GEMMs (cluster of 8 CTAs, DSMEM statistics)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider -x -k "residual_layernorm" > gpurun_out/r2l_tests.log 2>&1
tail -3 gpurun_out/r2l_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_tasks_gpu.py tests/test_pipeline_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2l_tests_model.log 2>&1
tail -3 gpurun_out/r2l_tests_model.log
for rep in 1 2; do
for v in 0 1; do
  RALF_DECODE_RESLN=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_bench_resln$v.$rep.json 2> gpurun_out/r2l_bench_resln$v.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2l_bench_resln$v.$rep.json").read().strip().splitlines()[-1])
print("RESLN=$v rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"], l["phases"]["decode_ms"], l["gpu_launches"])
PY
done
done
