#!/bin/bash
# round 2: programmatic dependent launch for single-wave grids only (the decode loop) -- parity under PDL, A/B
set -x
mkdir -p gpurun_out
RALF_PDL=2 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py tests/test_tasks_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2p_tests.log 2>&1
tail -3 gpurun_out/r2p_tests.log
for rep in 1 2; do
for v in 0 2; do
  RALF_PDL=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2p_bench_pdl$v.$rep.json 2> gpurun_out/r2p_bench_pdl$v.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2p_bench_pdl$v.$rep.json").read().strip().splitlines()[-1])
print("PDL=$v rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"])
PY
done
done
