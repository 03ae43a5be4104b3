#!/bin/bash
mkdir -p gpurun_out
B="--steps 5 --warmup 3 --no-cpu-baseline --no-extras"
for mb in 256 512 1024; do
  timeout 600 python bench.py $B --micro-batch $mb > gpurun_out/r2t_bench_mb$mb.json 2> gpurun_out/r2t_bench_mb$mb.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2t_bench_mb$mb.json').read().strip().splitlines()[-1]); print('mb$mb', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -3 gpurun_out/r2t_bench_mb$mb.err
done
