#!/bin/bash
# round 2, call c (2 GPUs): everything -m gpu, with hard limits
set -u
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -x > gpurun_out/r2c_multi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_multi.py > gpurun_out/r2c_gpu_all.log 2>&1
tail -3 gpurun_out/r2c_multi.log
grep -E "passed|failed" gpurun_out/r2c_gpu_all.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r2c_gpu_all.log | head -30
