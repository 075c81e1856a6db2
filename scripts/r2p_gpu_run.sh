#!/bin/bash
# round 2, call p (2 GPUs): the driver's N = 2 launch after the last bench changes (must end by itself)
set -u
mkdir -p gpurun_out
t0=$(date +%s)
DANBO_BENCH_SKIP_CONFIGS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2p_bench_2gpu.json 2> gpurun_out/r2p_bench_2gpu.err
echo "bench N=2 rc $? in $(( $(date +%s) - t0 )) s"
python -c "
import json; d=json.loads(open('gpurun_out/r2p_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['train'].get('value'), d['train'].get('loss'), d['train'].get('weak',{}).get('rays_per_s'))"
tail -3 gpurun_out/r2p_bench_2gpu.err
