#!/bin/bash
# final tree at 2 GPUs: NCCL training test + default bench line under torchrun
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_dist_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2z_tests_2gpu.log 2>&1
tail -2 gpurun_out/r2z_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2z_bench_2gpu.json 2> gpurun_out/r2z_bench_2gpu.err
tail -c 300 gpurun_out/r2z_bench_2gpu.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2z_bench_2gpu.json").read().strip().splitlines()[-1])
print("2gpu", l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline_knn"]["frac"], l["phases"])
print({k:(v.get("value"),v.get("ms_per_step")) for k,v in l.get("extras",{}).items() if isinstance(v,dict)})
PY
