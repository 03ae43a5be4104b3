#!/bin/bash
mkdir -p gpurun_out
for bn in 0 256; do
  echo "block_n=$bn"
  for sh in l2.c3 l3.c3 l3.c2 l4.c2 enc.qkv enc.o enc.l1 enc.l2 ckv; do
    RALF_BENCH_BN=$bn timeout 120 python profiles/gemm_bench.py $sh 2>&1 | grep "M="
  done
done
