#!/bin/bash
# which round-2 change moved the (chaotic) median gradient error of the 2-sample BatchNorm test?  (profiles/r2_train.md)
mkdir -p gpurun_out
for cfg in "RALF_GEMM_FOLD=1" "RALF_GEMM_FOLD=0" "RALF_GEMM_FOLD=0 RALF_BN_SLAB=2048" "RALF_GEMM_FOLD=0 RALF_BN_SLAB=256" "RALF_GEMM_FOLD=0 RALF_ATTN_TC=1"; do
  env $cfg timeout 300 python -m pytest "tests/test_train_gpu.py::test_training_gradients_match_oracle" -q -p no:cacheprovider -s 2>&1 | grep "^\.\?train_trunk=" | sed "s/^/[$cfg] /"
done
