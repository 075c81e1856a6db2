#!/bin/bash
# round 2, call e (1 GPU): full -m gpu suite, both bench arms, the mma aggregation net, training profile, ncu
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2e_gpu_all.log 2>&1
grep -E "passed|failed" gpurun_out/r2e_gpu_all.log | tail -2
grep -E "^FAILED|^ERROR" gpurun_out/r2e_gpu_all.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_ffma.json 2> gpurun_out/r2e_bench_ffma.err
DANBO_PAIR_LOGITS=mma DANBO_BENCH_SKIP_CONFIGS=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_mma.json 2> gpurun_out/r2e_bench_mma.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2e_bench_reference.json 2> gpurun_out/r2e_bench_reference.err
timeout 300 python scripts/train_profile.py > gpurun_out/r2e_train_profile.txt 2>&1
DANBO_PAIR_LOGITS=mma DANBO_BENCH_SKIP_CONFIGS=1 DANBO_BENCH_SKIP_TRAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2e_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2e_ncu_bench.log 2>&1
DANBO_PAIR_LOGITS=mma DANBO_BENCH_SKIP_CONFIGS=1 DANBO_BENCH_SKIP_TRAIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sample_mask|pair_logits|mlp_kernel|field_rows|composite|nearfar|ray_bias" -s 60 -c 14 -o gpurun_out/r2e_ncu_full python bench.py --steps 2 --warmup 1 > gpurun_out/r2e_ncu_full.log 2>&1
tail -c 600 gpurun_out/r2e_bench_ffma.json; echo; tail -c 300 gpurun_out/r2e_bench_reference.json
