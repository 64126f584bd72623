#!/bin/bash
# session-4 first visit: parity tests, default bench, ncu full capture of the tensor-path kernels.  $1 = tag
TAG=${1:-r01m}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 ) > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_${TAG}.log
( time timeout 900 python bench.py ) > gpurun_out/bench_cfg4_${TAG}.log 2>&1
for K in filter_select exact_topk; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${K}_cfg4_${TAG} \
    python tools/run_once.py --config 4 --users 18944 --items 400000 --reps 2 > gpurun_out/prof_${K}_cfg4_${TAG}.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_filter_${TAG}.csv \
    python tools/run_once.py --config 4 --users 151552 --reps 2 > gpurun_out/launches_filter_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log; tail -2 gpurun_out/bench_cfg4_${TAG}.log | cut -c1-2500
