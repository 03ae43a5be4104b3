#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_8gpu.json 2> gpurun_out/r2n_bench_8gpu.err
tail -5 gpurun_out/r2n_bench_8gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2n_bench_8gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_knn"]["frac"], d["roofline_knn"]["ms_per_launch"], d["roofline"]["frac"])
ex = d.get("extras", {})
print(json.dumps({k: v for k, v in ex.items() if k != "knn_sweep"}, indent=1)[:4000])
for r in ex.get("knn_sweep", {}).get("rows", []): print(r)
PY
