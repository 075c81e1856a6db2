#!/bin/bash
# round 2, call h (8 GPUs): the driver's own N = 8 launch of bench.py, with a hard limit
set -u
mkdir -p gpurun_out
DANBO_BENCH_VERBOSE=1 timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2h_bench_8gpu.json 2> gpurun_out/r2h_bench_8gpu.err
echo "rc $?"
tail -c 2500 gpurun_out/r2h_bench_8gpu.json
grep "bench rank 0" gpurun_out/r2h_bench_8gpu.err | tail -12
