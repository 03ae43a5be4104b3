#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_pipeline_gpu.py tests/test_data_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2j_tests.log 2>&1
tail -6 gpurun_out/r2j_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "knn", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], "api", d.get("e2e_model_api"))
for r in d["extras"]["knn_sweep"]["rows"]: print(r)
print({k: v for k, v in d["extras"].items() if k != "knn_sweep"})
PY
bash profiles/run_sanitizer.sh > /dev/null 2>&1
grep -H "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitizer_*.log
