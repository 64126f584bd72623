#!/bin/bash
# Round-end style measurements.  $1 = tag
TAG=${1:-r01}
mkdir -p gpurun_out
( time timeout 1700 python bench.py ) > gpurun_out/bench_cfg4_${TAG}.log 2>&1; echo "rc=$?" >> gpurun_out/bench_cfg4_${TAG}.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_${TAG}.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_${TAG}.csv \
    python bench.py --users 151552 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench_${TAG}.log 2>&1
for cfg in 1 2 3 5; do
  ( time timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 ) > gpurun_out/bench_cfg${cfg}_${TAG}.log 2>&1
done
tail -4 gpurun_out/bench_cfg4_${TAG}.log | cut -c1-1500; tail -4 gpurun_out/bench_ref_${TAG}.log | cut -c1-600
