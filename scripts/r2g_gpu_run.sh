#!/bin/bash
# round 2, call g (1 GPU): suite after the constant empty-sample path + 3-stream backward; bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -x > gpurun_out/r2g_gpu_all.log 2>&1
grep -E "passed|failed" gpurun_out/r2g_gpu_all.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/r2g_gpu_all.log | head -30
DANBO_BENCH_SKIP_CONFIGS=1 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err
tail -c 1200 gpurun_out/r2g_bench_1gpu.json
