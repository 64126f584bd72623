#!/bin/bash
# launch list + ncu full capture of the tensor-core filter kernel (cfg4 shape).  $1 = tag
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_filter_${TAG}.csv \
    python tools/run_once.py --config 4 --users 37888 --reps 2 > gpurun_out/launches_filter_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_select -s 1 -c 1 -f -o gpurun_out/prof_filter_cfg4_${TAG} \
    python tools/run_once.py --config 4 --users 18944 --items ${ITEMS:-400000} --reps 2 > gpurun_out/prof_filter_cfg4_${TAG}.log 2>&1
tail -2 gpurun_out/prof_filter_cfg4_${TAG}.log | cut -c1-300
