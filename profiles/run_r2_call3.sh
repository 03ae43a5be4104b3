#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_dist_gpu.py tests/test_train_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2c_tests.log 2>&1
tail -15 gpurun_out/r2c_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -5 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print(json.dumps(d.get("extras"), indent=1)[:6000])
PY
