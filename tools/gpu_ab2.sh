#!/bin/bash
# A/B of the build variants in build_variants/ (cfg4, one 151,552-user batch) against the round-1 build on the same box,
# with a correctness check of each (tensor == fma bit for bit)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== round-1 build"
( cd build_variants/r1 && timeout 300 python tools/run_once.py --config ${CFG:-4} --users ${USERS:-151552} --reps 3 2>&1 | tail -1 | cut -c1-60 )
for so in build_variants/v*.so; do
  echo "== $(grep "^$(basename $so .so):" build_variants/list.txt)"
  RMB200_LIB=$PWD/$so timeout 300 python tools/run_once.py --config ${CFG:-4} --users ${USERS:-151552} --reps 3 2>&1 | tail -2 | cut -c1-200 | sed 's/"kernel_ms.*//'
  RMB200_LIB=$PWD/$so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_filter_path_equals or sampled_threshold" 2>&1 | tail -1
done ) 2>&1 | tee gpurun_out/ab2.log
