#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x -s -k "fused_decode_chain" > gpurun_out/r2e_chain_tests.log 2>&1
tail -25 gpurun_out/r2e_chain_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py tests/test_tasks_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2e_model_tests.log 2>&1
tail -8 gpurun_out/r2e_model_tests.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout 400 python bench.py $B > gpurun_out/r2e_bench_chain.json 2> gpurun_out/r2e_bench_chain.err
RALF_DECODE_CHAIN=0 timeout 400 python bench.py $B > gpurun_out/r2e_bench_nochain.json 2> gpurun_out/r2e_bench_nochain.err
timeout 400 python bench.py $B --micro-batch 256 > gpurun_out/r2e_bench_chain_mb256.json 2> gpurun_out/r2e_bench_chain_mb256.err
tail -3 gpurun_out/r2e_bench_chain.err
for f in gpurun_out/r2e_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "api", d.get("e2e_model_api", {}).get("value"))
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
done
