#!/bin/bash
# round 2, call b (2 GPUs): multi-GPU training equivalence, mma aggregation net, pixel-error audit, 2-GPU bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pair_logits_mma.py -m gpu -q -s > gpurun_out/r2b_multi_mma.log 2>&1
timeout 600 python scripts/pixel_error_audit.py > gpurun_out/r2b_pixel_audit.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err
tail -5 gpurun_out/r2b_multi_mma.log
tail -c 1500 gpurun_out/r2b_bench_2gpu.json
