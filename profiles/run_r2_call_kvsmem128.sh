#!/bin/bash
# round 2: few-keys attention with 128-thread CTAs (three CTAs per SM) vs 256 (one)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2k_tests.log 2>&1
tail -3 gpurun_out/r2k_tests.log
for rep in 1 2; do
for v in 256 128; do
  RALF_KVSMEM_THREADS=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2k_bench_$v.$rep.json 2> gpurun_out/r2k_bench_$v.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2k_bench_$v.$rep.json").read().strip().splitlines()[-1])
print("KVSMEM_THREADS=$v rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"], l["phases"]["encode_ms"])
PY
done
done
timeout 600 ncu --profile-from-start off --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none -k regex:kvsmem --csv --log-file gpurun_out/r2k_kvsmem.csv python profiles/launch_slice.py > /dev/null 2>&1
grep kvsmem gpurun_out/r2k_kvsmem.csv | awk -F'","' '{print $5, $(NF)}' | head -6
