#!/bin/bash
# Run on the GPU box (under gpurun) from the repo root: launch list + full captures of the two dominant kernels.
# Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by hand afterwards.
set -x
mkdir -p gpurun_out
timeout 600 python profiles/breakdown.py 128 > gpurun_out/breakdown_128.json 2> gpurun_out/breakdown.err
ARGS="--steps 1 --warmup 0 --no-graph --no-cpu-baseline --skip-e2e"
timeout 1200 ncu --kernel-name-base demangled -k regex:ralf:: --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:knn_scan_kernel -c 1 -o gpurun_out/knn_scan \
    python bench.py $ARGS > gpurun_out/ncu_knn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gemm_bf16_kernel -s 30 -c 4 -o gpurun_out/gemm \
    python bench.py $ARGS > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
