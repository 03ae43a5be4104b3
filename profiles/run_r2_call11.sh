#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_train_slice.csv python profiles/train_slice.py > gpurun_out/r2k_train_slice.log 2>&1
python profiles/train_slice.py --summarize gpurun_out/r2k_train_slice.csv > gpurun_out/r2k_train_slice_summary.md
head -45 gpurun_out/r2k_train_slice_summary.md
for k in attn_bwd_dkv attn_bwd_dq "attention_kernel" bn_colstats col2im; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -c 2 -o gpurun_out/r2k_$k python profiles/train_slice.py > gpurun_out/r2k_ncu_$k.log 2>&1
  ncu -i gpurun_out/r2k_$k.ncu-rep --page raw --csv > gpurun_out/r2k_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep r2k
