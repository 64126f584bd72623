#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-4}
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_robust.py -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/robust_n$N.log 2>&1
( timeout 300 python bench.py --gpus $N --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/bench_cfg4_n${N}_strong.json; python -c "
import sys,json
d=json.loads(open('gpurun_out/bench_cfg4_n${N}_strong.json').read()); print(round(d['value']), d['ms_per_step'], d['per_rank'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'pageable', round(d['e2e_pageable']['value']))" ) > gpurun_out/scale_dbg_n$N.log 2>&1
cat gpurun_out/robust_n$N.log gpurun_out/scale_dbg_n$N.log
