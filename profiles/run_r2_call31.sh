#!/bin/bash
mkdir -p gpurun_out
for big in 0 1; do RALF_ATTN_TC_BIG=$big timeout 300 python profiles/encode_bench.py 350 240 2>&1 | tail -1; done
