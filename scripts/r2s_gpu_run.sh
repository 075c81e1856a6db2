#!/bin/bash
# round 2, call s (1 GPU): compute-sanitizer memcheck over the kernels added / changed this round
set -u
mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x -m gpu \
  "tests/test_gpu_parity.py::test_nearfar" "tests/test_gpu_parity.py::test_sample_mask_and_compaction" \
  "tests/test_gpu_parity.py::test_render_rays_end_to_end[render_fast]" "tests/test_gpu_parity.py::test_pair_list_overflow_is_detected_not_silent" \
  "tests/test_gpu_anerf.py::test_anerf_training_step_gradients" > gpurun_out/r2s_sanitizer_memcheck.log 2>&1
echo "sanitizer rc $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2s_sanitizer_memcheck.log | head -10
