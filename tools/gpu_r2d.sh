#!/bin/bash
# round 2, baseline call: GPU suite, cfg4 bench line, launch list, ncu of the filter kernel.  $1 = tag
cd "$(dirname "$0")/.."
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv,noheader > gpurun_out/env_${TAG}.txt
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_${TAG}.txt 2>&1
( time timeout 900 python bench.py ) > gpurun_out/bench_cfg4_${TAG}.log 2>&1
( time timeout 600 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_cfg3_${TAG}.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_${TAG}.csv \
    python bench.py --users 151552 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench_${TAG}.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:filter_select -c 1 -f -o gpurun_out/ncu_filter_${TAG} \
    python tools/run_once.py --config 4 --users 151552 --reps 1 > gpurun_out/ncu_filter_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.txt; tail -4 gpurun_out/bench_cfg4_${TAG}.log | cut -c1-1800; tail -3 gpurun_out/bench_cfg3_${TAG}.log | cut -c1-800
