#!/bin/bash
# Parity tests + smoke + quick kernel timings.  $1 = tag
TAG=${1:-chk}
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_${TAG}.log
for cfg in ${CFGS:-4 3 5 2 1}; do
  timeout 600 python tools/run_once.py --config $cfg --users ${USERS:-37888} --reps 3 > gpurun_out/once_cfg${cfg}_${TAG}.log 2>&1
done
tail -3 gpurun_out/smoke_${TAG}.log; tail -15 gpurun_out/pytest_gpu_${TAG}.log; tail -qn 1 gpurun_out/once_cfg*_${TAG}.log | cut -c1-330
