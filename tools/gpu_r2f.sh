#!/bin/bash
# round 2: rank-counting rewrite -- GPU suite, cfg3 bench line, ncu of score_select_kernel<.., AUC>.  $1 = tag
cd "$(dirname "$0")/.."
TAG=${1:-r02c}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_${TAG}.txt 2>&1
( time timeout 600 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_cfg3_${TAG}.log 2>&1
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_select -c 1 -f -o gpurun_out/ncu_auc_${TAG} \
    python tools/run_once.py --config 3 --users 37888 --reps 1 > gpurun_out/ncu_auc_${TAG}.log 2>&1
fi
tail -12 gpurun_out/pytest_gpu_${TAG}.txt; tail -4 gpurun_out/bench_cfg3_${TAG}.log | cut -c1-1500
