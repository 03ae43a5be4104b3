#!/bin/bash
# final tree at 8 GPUs, bench line without the extras (those are in r2_bench_z_8gpu_final.json of the build before the last kernels)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --no-cpu-baseline --no-extras > gpurun_out/r2z_bench_8gpu_noextras.json 2> gpurun_out/r2z_bench_8gpu_noextras.err
tail -c 300 gpurun_out/r2z_bench_8gpu_noextras.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2z_bench_8gpu_noextras.json").read().strip().splitlines()[-1])
print("8gpu", l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"], l["roofline_knn"]["frac"], l["phases"], l["clocks"])
PY
