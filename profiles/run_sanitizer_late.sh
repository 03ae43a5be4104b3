#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: TMA-epilogue GEMM (bulk tensor loads / stores, per-warp chunk
# rings), strided implicit convolutions (TMA element strides), half-TMEM attention.  Small shapes, each tool bounded.
set -x
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -x -q -p no:cacheprovider -o timeout=0"
T="tests/test_gemm_gpu.py::test_gemm_tma_epilogue_is_bit_identical_to_register_epilogue[19072-256-64-True-None] tests/test_gemm_gpu.py::test_gemm_tma_epilogue_is_bit_identical_to_register_epilogue[19072-256-64-False-None] tests/test_gemm_gpu.py::test_gemm_tma_epilogue_is_bit_identical_to_register_epilogue[20000-512-128-True-None] tests/test_gemm_gpu.py::test_conv_gemm_stride2_matches_conv2d_fp64_and_im2col[2-64-64-128-128-3] tests/test_gemm_gpu.py::test_conv_gemm_stride2_matches_conv2d_fp64_and_im2col[3-16-16-1024-128-1] tests/test_gemm_gpu.py::test_conv_gemm_stride2_matches_conv2d_fp64_and_im2col[2-22-15-256-256-3] tests/test_attention_gpu.py::test_encoder_attention_tcgen05_matches_fp64[5-128-64-8] tests/test_attention_gpu.py::test_encoder_attention_tcgen05_matches_fp64[2-200-200-8]"
for tool in memcheck initcheck synccheck; do
  timeout 420 $S --tool $tool --print-limit 20 --error-exitcode 9 $PY $T > gpurun_out/sanitizer_late_$tool.log 2>&1
done
timeout 420 $S --tool racecheck --racecheck-report analysis --print-limit 20 --error-exitcode 9 $PY $T > gpurun_out/sanitizer_late_racecheck.log 2>&1
grep -H "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitizer_late_*.log
