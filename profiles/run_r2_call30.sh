#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2x_slice.csv python profiles/launch_slice.py > gpurun_out/r2x_slice.log 2>&1
python profiles/summarize_slice.py gpurun_out/r2x_slice.csv > gpurun_out/r2x_slice_summary.md 2>&1
cat gpurun_out/r2x_slice_summary.md | head -40
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:attention_decode_kv16x4 -c 1 -o gpurun_out/r2x_kv16x4 python profiles/launch_slice.py > gpurun_out/r2x_ncu_kv16x4.log 2>&1
ncu -i gpurun_out/r2x_kv16x4.ncu-rep --page raw --csv > gpurun_out/r2x_kv16x4_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2x_kv16x4_raw.csv')))
hdr,units=rows[0],rows[1]; idx={h:i for i,h in enumerate(hdr)}
for w in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active']:
    print(w, rows[2][idx[w]], units[idx[w]])
PY
