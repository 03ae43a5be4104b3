#!/bin/bash
# round 2, final tree: full GPU suite + default bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2z_tests_final.log 2>&1
tail -4 gpurun_out/r2z_tests_final.log
timeout 900 python bench.py > gpurun_out/r2z_bench_final.json 2> gpurun_out/r2z_bench_final.err
tail -c 300 gpurun_out/r2z_bench_final.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2z_bench_final.json").read().strip().splitlines()[-1])
print("default", l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"], l["roofline_knn"]["frac"], l["phases"], l["gpu_launches"])
print([(o["kernel"][:30], o["frac"]) for o in l["roofline_other"]])
print({k:(v.get("value"),v.get("ms_per_step")) for k,v in l.get("extras",{}).items() if isinstance(v,dict)})
print(l.get("e2e_model_api"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
tail -c 500 gpurun_out/r2z_bench_ref.json
