#!/bin/bash
# round 2, call l (1 GPU): suite + bench after the eval-graph pack removal and the faster empty trunk
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r2l_gpu_all.log 2>&1; tail -3 gpurun_out/r2l_gpu_all.log
grep -E "^FAILED|^ERROR" gpurun_out/r2l_gpu_all.log | head
DANBO_BENCH_SKIP_CONFIGS=1 timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2l_bench_1gpu.json 2> gpurun_out/r2l_bench_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2l_bench_1gpu.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['train']['value'], d['train']['ms_per_iter'], d['roofline']['frac'])"
