#!/bin/bash
# round 2: ncu --set full of the new kernels inside the bench-step slice (TMA-epilogue GEMMs, half-TMEM attention)
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tepi|attention_tc3" -c 16 -o gpurun_out/r2n2_new python profiles/launch_slice.py > gpurun_out/r2n2_ncu.log 2>&1
ncu -i gpurun_out/r2n2_new.ncu-rep --page raw --csv > gpurun_out/r2n2_new_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2n2_new_raw.csv')))
hdr,units=rows[0],rows[1]; idx={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_tensor.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sector_hit_rate.pct','launch__grid_size']
for r in rows[2:]:
    print(r[idx['Kernel Name']][:60], '|', ' | '.join(f"{w.split('.')[0].replace('__','_')[-28:]}={r[idx[w]]}{units[idx[w]]}" for w in want if w in idx))
PY
