#!/bin/bash
# round 2, final build: the default bench line at 8 GPUs (weak scaling, extras: configs[2] training, strong-scaled configs[4]) + smoke()
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2g_bench_8gpu.json 2> gpurun_out/r2g_bench_8gpu.err
tail -c 400 gpurun_out/r2g_bench_8gpu.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2g_bench_8gpu.json").read().strip().splitlines()[-1])
print("8gpu", l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"], l["roofline_knn"]["frac"], l["clocks"])
for k,v in l.get("extras",{}).items():
    if isinstance(v,dict): print(k, v.get("value"), v.get("ms_per_step"), json.dumps(v.get("allreduce"))[:300] if "allreduce" in v else "")
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
