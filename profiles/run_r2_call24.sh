#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -rs > gpurun_out/r2s_gpu_tests.log 2>&1
tail -6 gpurun_out/r2s_gpu_tests.log
python -c "
import __graft_entry__ as g
g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
tail -2 gpurun_out/r2s_smoke.log
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
tail -3 gpurun_out/r2s_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_ref.json 2> gpurun_out/r2s_bench_ref.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], "knn", d["roofline_knn"]["frac"], "api", d.get("e2e_model_api",{}).get("value"))
print({k: (v if k != "knn_sweep" else "...") for k, v in d["extras"].items()})
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
r = json.loads(open("gpurun_out/r2s_bench_ref.json").read().strip().splitlines()[-1])
print("ref arm", r["value"], r["unit"], r["cpu_baseline"]["cores"])
PY
