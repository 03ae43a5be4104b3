#!/bin/bash
# round 2: BN = 64 TMA-epilogue kernel (stem, layer-1 conv1 / conv2) -- parity, A/B, then a fresh launch-list slice
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2t4_tests.log 2>&1
tail -5 gpurun_out/r2t4_tests.log
for v in 0 1; do
  for sh in l1.c1 l1.c2; do
    RALF_TEPI_BN64=$v timeout 300 python profiles/gemm_bench.py $sh 2>&1 | grep "M=" | sed "s/^/BN64=$v /"
  done
done
timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2t4_tests_model.log 2>&1
tail -3 gpurun_out/r2t4_tests_model.log
for rep in 1 2; do
for v in 0 1; do
  RALF_TEPI_BN64=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2t4_bench_bn64_$v.$rep.json 2> gpurun_out/r2t4_bench_bn64_$v.$rep.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2t4_bench_bn64_$v.$rep.json").read().strip().splitlines()[-1])
print("BN64=$v rep $rep", l["value"], l["ms_per_step"], l["e2e"]["value"])
PY
done
done
timeout 900 ncu --profile-from-start off --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t4_slice.csv python profiles/launch_slice.py > gpurun_out/r2t4_slice.log 2>&1
python profiles/summarize_slice.py gpurun_out/r2t4_slice.csv > gpurun_out/r2t4_slice_summary.md 2>&1
head -30 gpurun_out/r2t4_slice_summary.md
