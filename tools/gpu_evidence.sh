#!/bin/bash
# Round evidence: parity tests, bench lines of every configuration, reference arm, launch list, ncu captures.  $1 = tag
TAG=${1:-r01}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 ) > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_${TAG}.log
( time timeout 900 python bench.py ) > gpurun_out/bench_cfg4_${TAG}.log 2>&1; echo "rc=$?" >> gpurun_out/bench_cfg4_${TAG}.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_${TAG}.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_${TAG}.csv \
    python bench.py --users 151552 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench_${TAG}.log 2>&1
for K in filter_select exact_topk user_metrics; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${K}_cfg4_${TAG} \
    python tools/run_once.py --config 4 --users 151552 --reps 2 > gpurun_out/prof_${K}_cfg4_${TAG}.log 2>&1
done
for cfg in 1 2 3 5; do
  ( time timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 ) > gpurun_out/bench_cfg${cfg}_${TAG}.log 2>&1
done
tail -3 gpurun_out/pytest_gpu_${TAG}.log; for c in 1 2 3 4 5; do grep -h '^{' gpurun_out/bench_cfg${c}_${TAG}.log | cut -c1-160; done
