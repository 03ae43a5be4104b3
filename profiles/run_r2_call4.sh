#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_dist_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2d_tests_2gpu.log 2>&1
tail -5 gpurun_out/r2d_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err
tail -5 gpurun_out/r2d_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2d_bench_2gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
ex = d.get("extras", {})
print(json.dumps({k: v for k, v in ex.items() if k != "knn_sweep"}, indent=1)[:5000])
for r in ex.get("knn_sweep", {}).get("rows", []): print(r)
PY
