#!/bin/bash
# round-end style evidence.  $1 = tag
cd "$(dirname "$0")/.."
TAG=${1:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw,power.limit --format=csv,noheader > gpurun_out/env_${TAG}.txt; lscpu | grep "Model name\|^CPU(s)" >> gpurun_out/env_${TAG}.txt
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_${TAG}.txt 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_cfg4_${TAG}.log 2>&1
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_${TAG}.log 2>&1
for cfg in 3 5 2 1; do
  ( time timeout 400 python bench.py --config $cfg --steps 5 --warmup 3 ) > gpurun_out/bench_cfg${cfg}_${TAG}.log 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_${TAG}.csv \
    python bench.py --users 151552 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench_${TAG}.log 2>&1
( time timeout 300 python bench.py --zero-users 0.01 --no-cpu-baseline ) > gpurun_out/bench_cfg4_zero1pct_${TAG}.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.txt 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.txt; for f in gpurun_out/bench_*_${TAG}.log; do grep '^{' $f | cut -c1-400; done; tail -3 gpurun_out/smoke_${TAG}.txt
