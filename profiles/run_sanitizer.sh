#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests (run under gpurun from the repo root; SURVEY.md 5 asks for
# memcheck / racecheck targets).  The sanitizer slows kernels 10-100x: only the small-shape parametrisations are selected,
# every run is bounded by `timeout`, and the logs land in gpurun_out/.  A clean run ends with "ERROR SUMMARY: 0 errors".
set -x
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -x -q -p no:cacheprovider -o timeout=0"
# memcheck: out-of-bounds / misaligned accesses (global + shared), leak of device allocations at exit
timeout 1500 $S --tool memcheck --print-limit 20 --error-exitcode 9 \
    $PY "tests/test_knn_gpu.py::test_knn_topk_matches_oracle[257-512-2-16]" "tests/test_knn_gpu.py::test_knn_topk_matches_oracle[20-512-2-16]" \
        "tests/test_knn_gpu.py::test_knn_topk_matches_oracle[1000-100-3-8]" \
        "tests/test_attention_gpu.py::test_encoder_attention_tcgen05_matches_fp64[5-128-64-8]" \
        "tests/test_attention_gpu.py::test_cross_attention_decode_matches_fp64" \
        "tests/test_attention_gpu.py::test_self_attention_decode_append_with_padding_mask" \
        "tests/test_attention_gpu.py::test_kv24_cache_gemm_and_cross_attention[2-40]" \
        tests/test_gemm_gpu.py > gpurun_out/sanitizer_memcheck.log 2>&1
# racecheck: shared-memory hazards between the warps of a CTA (the hand-rolled smem staging / candidate lists)
timeout 1500 $S --tool racecheck --racecheck-report analysis --print-limit 20 --error-exitcode 9 \
    $PY "tests/test_knn_gpu.py::test_knn_topk_matches_oracle[1000-100-3-8]" \
        "tests/test_attention_gpu.py::test_fusion_attention_fewkeys_matches_fp64" \
        "tests/test_attention_gpu.py::test_layernorm_matches_fp64" > gpurun_out/sanitizer_racecheck.log 2>&1
# synccheck: barrier misuse (divergent __syncthreads / mbarrier) in the warp-specialised kernels
timeout 1500 $S --tool synccheck --print-limit 20 --error-exitcode 9 \
    $PY "tests/test_attention_gpu.py::test_encoder_attention_tcgen05_matches_fp64[5-128-64-8]" \
        "tests/test_knn_gpu.py::test_knn_topk_matches_oracle[257-512-2-16]" > gpurun_out/sanitizer_synccheck.log 2>&1
grep -H "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_*.log
