#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --profile-from-start off --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2zz_slice.csv python profiles/launch_slice.py > gpurun_out/r2zz_slice.log 2>&1
python profiles/summarize_slice.py gpurun_out/r2zz_slice.csv > gpurun_out/r2zz_slice_summary.md 2>&1
head -22 gpurun_out/r2zz_slice_summary.md
