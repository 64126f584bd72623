#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:score_select -c 1 -f -o gpurun_out/ncu_auc_${1:-x} \
    python tools/run_once.py --config 3 --users 37888 --reps 1 > gpurun_out/ncu_auc_${1:-x}.log 2>&1
tail -2 gpurun_out/ncu_auc_${1:-x}.log
