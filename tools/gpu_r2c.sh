#!/bin/bash
# round 2, call C: A/B of build variants against the round-1 build on the same box + the robustness tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== round-1 build"
( cd build_variants/r1 && timeout 300 python tools/run_once.py --config 4 --users 151552 --reps 3 2>&1 | tail -1 | cut -c1-120 )
for so in build_variants/v*.so; do
  echo "== $(grep "^$(basename $so .so):" build_variants/list.txt)"
  RMB200_LIB=$PWD/$so timeout 300 python tools/run_once.py --config 4 --users 151552 --reps 3 2>&1 | tail -2 | cut -c1-200
done ) > gpurun_out/r2c_ab.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_robust.py tests/test_gpu_parity.py -x -q -m gpu -k "robust or heavy or wide or zero or bias_much or cfg2_full or unscalable or nan_payload or empty_user or tensor_filter_path or full_catalogue_cfg4 or sampled_threshold or golden or cfg4_shape or k_sweep or tie_breaking" 2>&1 | tail -25 > gpurun_out/r2c_tests.log 2>&1
cat gpurun_out/r2c_ab.log gpurun_out/r2c_tests.log
