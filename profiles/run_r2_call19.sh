#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py tests/test_pipeline_gpu.py tests/test_tasks_gpu.py -q -p no:cacheprovider -x -s > gpurun_out/r2o_tests.log 2>&1
grep -h "logits\|passed\|failed\|Error" gpurun_out/r2o_tests.log | tail -14
B="--steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 python bench.py $B > gpurun_out/r2o_bench_kv16.json 2> gpurun_out/r2o_bench_kv16.err
RALF_KVFMT=24 timeout 600 python bench.py $B > gpurun_out/r2o_bench_kv24.json 2> gpurun_out/r2o_bench_kv24.err
for f in gpurun_out/r2o_bench_kv*.json; do python -c "
import json,sys
d = json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'])"; done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:attention_decode_kv16 -c 1 -o gpurun_out/r2o_kv16 python profiles/launch_slice.py > gpurun_out/r2o_ncu_kv16.log 2>&1
ncu -i gpurun_out/r2o_kv16.ncu-rep --page raw --csv > gpurun_out/r2o_kv16_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2o_kv16_raw.csv')))
hdr,units=rows[0],rows[1]; idx={h:i for i,h in enumerate(hdr)}
for w in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active']:
    print(w, rows[2][idx[w]], units[idx[w]])
PY
