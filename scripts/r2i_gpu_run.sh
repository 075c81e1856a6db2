#!/bin/bash
# round 2, call i (2 GPUs): final validation - full suite, then the driver's N = 2 launch must END by itself
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -s > gpurun_out/r2i_gpu_all.log 2>&1
grep -E "passed|failed" gpurun_out/r2i_gpu_all.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/r2i_gpu_all.log | head -30
t0=$(date +%s)
DANBO_BENCH_SKIP_CONFIGS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
echo "bench N=2 rc $? in $(( $(date +%s) - t0 )) s"
CUDA_VISIBLE_DEVICES=0 DANBO_BENCH_SKIP_CONFIGS=1 timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench_1gpu.json 2> gpurun_out/r2i_bench_1gpu.err
echo "bench N=1 rc $?"
python -c "
import json
for f in ('r2i_bench_1gpu','r2i_bench_2gpu'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['train'].get('value'), d['train'].get('ms_per_iter'))
"
