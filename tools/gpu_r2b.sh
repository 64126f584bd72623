#!/bin/bash
# round 2, call B: where the new filter kernel's time goes -- round-1 build on the same box, DBG decomposition, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv,noheader
echo "== round-1 build"
( cd build_variants/r1 && timeout 300 python tools/run_once.py --config 4 --users 151552 --reps 3 2>&1 | tail -1 | cut -c1-120 )
for so in build_variants/v*.so; do
  for e in ${ENVS:-0}; do
    echo "== $(grep "^$(basename $so .so):" build_variants/list.txt) RMB200_DBG=$e"
    RMB200_DBG=$e RMB200_LIB=$PWD/$so timeout 300 python tools/run_once.py --config 4 --users 151552 --reps 3 2>&1 | tail -1 | cut -c1-120
  done
done
so=build_variants/v0.so
for e in 1 2 9; do
  echo "== $(grep "^v0:" build_variants/list.txt) RMB200_DBG=$e"
  RMB200_DBG=$e RMB200_LIB=$PWD/$so timeout 300 python tools/run_once.py --config 4 --users 151552 --reps 3 2>&1 | tail -1 | cut -c1-120
done
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader
) > gpurun_out/r2b_ab.log 2>&1
RMB200_LIB=$PWD/build_variants/v1.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:filter_select -c 1 -f -o gpurun_out/r2b_filter python tools/run_once.py --config 4 --users 151552 --reps 1 > gpurun_out/r2b_ncu.log 2>&1
timeout 900 python -m pytest tests/test_gpu_robust.py -x -q -m gpu -k "unscalable or nan_payload or empty_user" 2>&1 | tail -15 > gpurun_out/r2b_tests.log
cat gpurun_out/r2b_ab.log gpurun_out/r2b_tests.log; tail -3 gpurun_out/r2b_ncu.log
