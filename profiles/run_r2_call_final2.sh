#!/bin/bash
# round 2, final build: full GPU suite, default bench line (extras + cpu baseline), micro-batch A/B, launch-list slice,
# ncu --set full of the half-TMEM attention kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2f_tests.log 2>&1
tail -4 gpurun_out/r2f_tests.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 600 gpurun_out/r2f_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
print("default", l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"], l.get("roofline_knn",{}).get("frac"), l.get("roofline_other"))
print({k:(v.get("value"),v.get("ms_per_step")) for k,v in l.get("extras",{}).items() if isinstance(v,dict)})
PY
for mb in 128 512; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --micro-batch $mb > gpurun_out/r2f_bench_mb$mb.json 2> gpurun_out/r2f_bench_mb$mb.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2f_bench_mb$mb.json").read().strip().splitlines()[-1])
print("MB=$mb", l["value"], l["ms_per_step"], l["e2e"]["value"])
PY
done
timeout 900 ncu --profile-from-start off --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_slice.csv python profiles/launch_slice.py > gpurun_out/r2f_slice.log 2>&1
python profiles/summarize_slice.py gpurun_out/r2f_slice.csv > gpurun_out/r2f_slice_summary.md 2>&1
head -16 gpurun_out/r2f_slice_summary.md
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:attention_tc3 -c 1 -o gpurun_out/r2f_attn_tc3 python profiles/launch_slice.py > gpurun_out/r2f_ncu_attn.log 2>&1
ncu -i gpurun_out/r2f_attn_tc3.ncu-rep --page raw --csv > gpurun_out/r2f_attn_tc3_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2f_attn_tc3_raw.csv')))
hdr,units=rows[0],rows[1]; idx={h:i for i,h in enumerate(hdr)}
for w in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem']:
    if w in idx: print(w, rows[2][idx[w]], units[idx[w]])
PY
