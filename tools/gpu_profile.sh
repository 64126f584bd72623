#!/bin/bash
# ncu evidence for the dominant kernel.  $1 = tag (e.g. r01a)
TAG=${1:-r01}
mkdir -p gpurun_out
# launch list (per-launch device time, cold & serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python tools/run_once.py --users 18944 --items 200000 --reps 2 > gpurun_out/launches_${TAG}.log 2>&1
# full capture of the scoring kernel
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:score_select -s 1 -c 1 -f -o gpurun_out/prof_${TAG} \
    python tools/run_once.py --users 18944 --items 200000 --reps 2 > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
