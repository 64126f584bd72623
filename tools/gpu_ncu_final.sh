#!/bin/bash
# ncu --set full captures of the two dominant kernels (one launch each).  $1 = tag
cd "$(dirname "$0")/.."
TAG=${1:-x}
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:filter_select -c 1 -f -o gpurun_out/ncu_filter_${TAG} \
    python tools/run_once.py --config 4 --users 151552 --reps 1 > gpurun_out/ncu_filter_${TAG}.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:score_select -c 1 -f -o gpurun_out/ncu_auc_${TAG} \
    python tools/run_once.py --config 3 --users 75776 --reps 1 > gpurun_out/ncu_auc_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_filter_${TAG}.log; tail -1 gpurun_out/ncu_auc_${TAG}.log
