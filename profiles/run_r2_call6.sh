#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_slice.csv python profiles/launch_slice.py > gpurun_out/r2f_slice.log 2>&1
python profiles/summarize_slice.py gpurun_out/r2f_slice.csv > gpurun_out/r2f_slice_summary.md 2>&1
cat gpurun_out/r2f_slice_summary.md | head -40
grep "decode_chain" gpurun_out/r2f_slice.csv | head -12
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:decode_chain -s 6 -c 3 -o gpurun_out/r2f_chain python profiles/launch_slice.py > gpurun_out/r2f_ncu_chain.log 2>&1
ncu -i gpurun_out/r2f_chain.ncu-rep --page raw --csv > gpurun_out/r2f_chain_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
