#!/bin/bash
# round 2, call o (1 GPU): final launch list + ncu --set full of the render kernels (never bench values)
set -u
mkdir -p gpurun_out
DANBO_BENCH_SKIP_CONFIGS=1 DANBO_BENCH_SKIP_TRAIN=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2o_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2o_ncu_bench.log 2>&1
DANBO_BENCH_SKIP_CONFIGS=1 DANBO_BENCH_SKIP_TRAIN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sample_mask|pair_logits|mlp_kernel|field_rows|composite|nearfar_finish|ray_bias_kernel|empty_rows" -s 50 -c 12 -o gpurun_out/r2o_ncu_full python bench.py --steps 2 --warmup 1 > gpurun_out/r2o_ncu_full.log 2>&1
ncu -i gpurun_out/r2o_ncu_full.ncu-rep --page raw --csv > gpurun_out/r2o_ncu_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2o_*
