#!/bin/bash
# round 2, call k (1 GPU): occupancy A/B of the tensor-core aggregation net, fresh launch list
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_pair_logits_mma.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2k_tests.log 2>&1; tail -2 gpurun_out/r2k_tests.log
for occ in 3 4 5; do
  DANBO_PAIR_LOGITS_BLOCKS=$occ DANBO_BENCH_SKIP_CONFIGS=1 DANBO_BENCH_SKIP_TRAIN=1 timeout 150 python bench.py --steps 20 --warmup 3 > gpurun_out/r2k_bench_occ$occ.json 2> gpurun_out/r2k_bench_occ$occ.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_occ$occ.json').read().strip().splitlines()[-1]); print('occ $occ', d['ms_per_step'], d['value'])"
done
DANBO_PAIR_LOGITS_BLOCKS=5 DANBO_BENCH_SKIP_CONFIGS=1 DANBO_BENCH_SKIP_TRAIN=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2k_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2k_ncu_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2k_launches_bench.csv | head -14
