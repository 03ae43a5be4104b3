#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -c 300 gpurun_out/r2h_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["e2e"]["value"]); print(l["phases"]); print(l["roofline"]["frac"], [ (o["kernel"][:40], o["frac"], o["ms_per_launch"]) for o in l["roofline_other"]])
PY
