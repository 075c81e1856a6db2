#!/bin/bash
# round 2, call n (1 GPU): the default bench line (with configs), A-NeRF training iteration, smoke()
set -u
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err; echo "bench rc $?"
python -c "
import json; d=json.loads(open('gpurun_out/r2n_bench_default.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['train'].get('value'), {k:(v.get('ms_per_image'), v.get('ms_images'), v.get('ms_lattice'), v.get('error')) for k,v in d['configs'].items()})"
timeout 300 python scripts/train_anerf_bench.py > gpurun_out/r2n_train_anerf.log 2>&1; tail -3 gpurun_out/r2n_train_anerf.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2n_smoke.log
