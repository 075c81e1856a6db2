#!/bin/bash
# round 2, call j (1 GPU): the full -m gpu suite as the driver runs it, then smoke()
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_gpu_all.log 2>&1
tail -4 gpurun_out/r2j_gpu_all.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2j_smoke.log 2>&1
echo "smoke rc $?"; tail -8 gpurun_out/r2j_smoke.log
