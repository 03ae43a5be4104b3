#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests (run under gpurun from the repo root).  The sanitizer slows kernels
# 10-100x: only small-shape parametrisations are selected, every run is bounded by `timeout`, logs land in gpurun_out/.
# A clean run ends with "ERROR SUMMARY: 0 errors".  Tools: memcheck (OOB / misaligned), initcheck (reads of uninitialised
# global memory -- the suspicion VERDICT r1 raised about the graph-replayed training step), racecheck (shared-memory hazards),
# synccheck (barrier misuse).
set -x
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -x -q -p no:cacheprovider -o timeout=0"
KNN="tests/test_knn_gpu.py::test_knn_topk_matches_oracle[257-512-2-16] tests/test_knn_gpu.py::test_knn_topk_matches_oracle[20-512-2-16] tests/test_knn_gpu.py::test_knn_topk_matches_oracle[1000-100-3-8] tests/test_knn_gpu.py::test_knn_uncertified_queries_are_fixed_on_the_device"
ATT="tests/test_attention_gpu.py::test_encoder_attention_tcgen05_matches_fp64[5-128-64-8] tests/test_attention_gpu.py::test_cross_attention_decode_matches_fp64 tests/test_attention_gpu.py::test_self_attention_decode_append_with_padding_mask tests/test_attention_gpu.py::test_kv24_cache_gemm_and_cross_attention[2-40]"
timeout 900 $S --tool memcheck --print-limit 20 --error-exitcode 9 $PY $KNN $ATT tests/test_gemm_gpu.py ${EXTRA_MEMCHECK} > gpurun_out/sanitizer_memcheck.log 2>&1
# initcheck: the model path end to end at a small shape (engine vs golden), one training step, the k-NN fix-up
timeout 900 $S --tool initcheck --print-limit 20 --error-exitcode 9 $PY $KNN \
    "tests/test_model_gpu.py::test_engine_matches_reference_golden_pku" \
    "tests/test_train_gpu.py::test_train_steps_reduce_loss_and_update_state_dict" ${EXTRA_INITCHECK} > gpurun_out/sanitizer_initcheck.log 2>&1
timeout 900 $S --tool racecheck --racecheck-report analysis --print-limit 20 --error-exitcode 9 \
    $PY "tests/test_knn_gpu.py::test_knn_topk_matches_oracle[1000-100-3-8]" \
        "tests/test_attention_gpu.py::test_fusion_attention_fewkeys_matches_fp64" \
        "tests/test_attention_gpu.py::test_layernorm_matches_fp64" ${EXTRA_RACECHECK} > gpurun_out/sanitizer_racecheck.log 2>&1
timeout 600 $S --tool synccheck --print-limit 20 --error-exitcode 9 \
    $PY "tests/test_attention_gpu.py::test_encoder_attention_tcgen05_matches_fp64[5-128-64-8]" \
        "tests/test_knn_gpu.py::test_knn_topk_matches_oracle[257-512-2-16]" ${EXTRA_SYNCCHECK} > gpurun_out/sanitizer_synccheck.log 2>&1
grep -H "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_*.log
