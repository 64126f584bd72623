#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-s}
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_ties.py tests/test_gpu_full_order.py -x -q -m gpu 2>&1 | tail -3
  for sl in 1 0; do echo "== RMB200_SLICES=$sl (0 = automatic)"; RMB200_SLICES=$sl timeout 100 python tools/run_once.py --config 1 --users 6040 --reps 3 2>&1 | tail -1 | cut -c1-330; done
  ( time timeout 900 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
  timeout 200 python bench.py --config 1 --steps 10 --warmup 3 2>&1 | grep '^{' > gpurun_out/bench_cfg1_${TAG}.json; cut -c1-300 gpurun_out/bench_cfg1_${TAG}.json
) > gpurun_out/slices_${TAG}.log 2>&1
cat gpurun_out/slices_${TAG}.log
