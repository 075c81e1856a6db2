#!/bin/bash
# round 2, call d (2 GPUs): multi-GPU test with surviving logs, then the rest of the suite
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -x > gpurun_out/r2d_multi.log 2>&1
cp gpurun_out/multi_worker.log gpurun_out/r2d_multi_worker.log 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_multi.py > gpurun_out/r2d_gpu_all.log 2>&1
tail -3 gpurun_out/r2d_multi.log
grep -E "passed|failed" gpurun_out/r2d_gpu_all.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r2d_gpu_all.log | head -30
