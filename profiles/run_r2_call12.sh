#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_autograd_ops_gpu.py tests/test_dropout_gpu.py tests/test_train_dist_gpu.py tests/test_data_gpu.py -q -p no:cacheprovider -x > gpurun_out/r2l_tests.log 2>&1
tail -6 gpurun_out/r2l_tests.log
timeout 300 python profiles/train_bench.py 32 10 > gpurun_out/r2l_train.json 2> gpurun_out/r2l_train.err
cat gpurun_out/r2l_train.json
