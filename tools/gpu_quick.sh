#!/bin/bash
# tests + cfg4 launch list + default bench.  $1 = tag
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_${TAG}.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_filter_${TAG}.csv \
    python tools/run_once.py --config 4 --users 151552 --reps 2 > gpurun_out/launches_filter_${TAG}.log 2>&1
( time timeout 1700 python bench.py ) > gpurun_out/bench_cfg4_${TAG}.log 2>&1
( time timeout 900 python bench.py --config 5 ) > gpurun_out/bench_cfg5_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log
