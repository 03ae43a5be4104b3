#!/bin/bash
# Round-2 GPU call 2: whole suite (0 skipped expected) + compute-sanitizer (memcheck / initcheck / racecheck / synccheck).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -rs > gpurun_out/r2b_gpu_tests.log 2>&1
tail -15 gpurun_out/r2b_gpu_tests.log
bash profiles/run_sanitizer.sh
