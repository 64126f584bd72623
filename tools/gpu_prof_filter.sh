#!/bin/bash
# ncu full capture of filter_select_kernel (cfg4 shape, one wave of CTAs).  $1 = tag
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_select -s 1 -c 1 -f -o gpurun_out/prof_filter_select_cfg4_${TAG} \
    python tools/run_once.py --config 4 --users 18944 --items ${ITEMS:-1000000} --reps 2 > gpurun_out/prof_filter_select_cfg4_${TAG}.log 2>&1
