#!/bin/bash
mkdir -p gpurun_out
export RALF_CHAIN_ACC=1 RALF_CHAIN_PAIR=1 RALF_CHAIN_PREFETCH=0
for d in 0 1 2 3; do
  RALF_CHAIN_DEBUG=$d timeout 300 python profiles/chain_bench.py > gpurun_out/r2i_chain_dbg$d.json 2> gpurun_out/r2i_chain_dbg$d.err
  echo "debug=$d"; cat gpurun_out/r2i_chain_dbg$d.json
done
