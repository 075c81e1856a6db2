#!/bin/bash
# round 2, call f (2 GPUs): full suite (mma default, backward streams), tc-vs-simt backward, bench at N = 1 and 2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2f_gpu_all.log 2>&1
cp gpurun_out/multi_worker.log gpurun_out/r2f_multi_worker.log 2>/dev/null
grep -E "passed|failed" gpurun_out/r2f_gpu_all.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/r2f_gpu_all.log | head -30
timeout 300 python scripts/bwd_compare.py 16 192 > gpurun_out/r2f_bwd_compare.log 2>&1
CUDA_VISIBLE_DEVICES=0 DANBO_BENCH_SKIP_CONFIGS=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err
CUDA_VISIBLE_DEVICES=0 DANBO_BENCH_SKIP_CONFIGS=1 DANBO_BWD_STREAMS=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_1gpu_1stream.json 2> gpurun_out/r2f_bench_1gpu_1stream.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err
tail -c 400 gpurun_out/r2f_bench_1gpu.json; echo; tail -c 1500 gpurun_out/r2f_bench_2gpu.json
