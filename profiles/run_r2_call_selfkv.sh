#!/bin/bash
# round 2: 16-bit self-attention cache -- parity tests, then same-box A/B of the default bench step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py tests/test_tasks_gpu.py tests/test_pipeline_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2z_tests.log 2>&1
tail -5 gpurun_out/r2z_tests.log
for rep in 1 2; do
for v in 0 1; do
  RALF_SELF_KV16=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2z_bench_self$v.$rep.json 2> gpurun_out/r2z_bench_self$v.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2z_bench_self$v.$rep.json").read().strip().splitlines()[-1])
print("SELF_KV16=$v rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"])
PY
done
done
