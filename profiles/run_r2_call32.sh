#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_attention_gpu.py tests/test_pipeline_gpu.py tests/test_knn_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2y_tests.log 2>&1
tail -4 gpurun_out/r2y_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
tail -3 gpurun_out/r2y_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2y_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], "knn", d["roofline_knn"]["frac"], "api", d.get("e2e_model_api",{}).get("value"))
print({k: (v if k != "knn_sweep" else "...") for k, v in d["extras"].items()})
PY
