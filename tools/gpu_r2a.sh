#!/bin/bash
# round 2, call A: robustness tests of the interval filter + A/B of the append path (build_variants/: v0 call-based, v1 predicated, v2 stats)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_env.txt
( timeout 1500 python -m pytest tests/test_gpu_robust.py tests/test_gpu_parity.py -x -q -m gpu -k "robust or heavy or wide or zero or bias_much or cfg2_full or unscalable or nan_payload or empty_user or tensor_filter_path or full_catalogue_cfg4 or sampled_threshold or golden or cfg4_shape or k_sweep or tie_breaking" 2>&1 | tail -25 ) > gpurun_out/r2a_tests.log 2>&1
for so in build_variants/v*.so; do
  echo "== $(grep "^$(basename $so .so):" build_variants/list.txt)"
  RMB200_LIB=$PWD/$so timeout 300 python tools/run_once.py --config 4 --users 151552 --reps 3 2>&1 | tail -2 | cut -c1-200
done > gpurun_out/r2a_ab.log 2>&1
cat gpurun_out/r2a_tests.log gpurun_out/r2a_ab.log
