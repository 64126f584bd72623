#!/bin/bash
# round 2: multi-GPU call + strong-scaling bench on N GPUs.  $1 = tag, $2 = N
cd "$(dirname "$0")/.."
TAG=${1:-r02b}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/env_${TAG}.txt
nvidia-smi topo -m >> gpurun_out/env_${TAG}.txt 2>&1
( time timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu ) > gpurun_out/pytest_multi_${TAG}.txt 2>&1
DEVS=$(seq -s, 0 $((N-1)))
( for dv in 0 $DEVS; do echo "== one call, devices $dv (pageable host memory, 1M users)"; timeout 300 python tools/run_once.py --config 4 --users 1000000 --reps 3 --devices $dv 2>&1 | tail -2 | cut -c1-400; done ) > gpurun_out/incall_${TAG}.log 2>&1
( time timeout 600 python bench.py --gpus $N --no-cpu-baseline ) > gpurun_out/bench_cfg4_n${N}_${TAG}.log 2>&1
if [ -n "$ALSO" ]; then for M in $ALSO; do ( time timeout 600 python bench.py --gpus $M --no-cpu-baseline ) > gpurun_out/bench_cfg4_n${M}_${TAG}.log 2>&1; done; fi
tail -5 gpurun_out/pytest_multi_${TAG}.txt; cat gpurun_out/incall_${TAG}.log; for f in gpurun_out/bench_cfg4_n*_${TAG}.log; do grep '^{' $f | cut -c1-330; done
