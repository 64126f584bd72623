#!/bin/bash
# ncu full captures: cfg4 (plain top-K) and cfg3 (AUC) on one wave of users.  $1 = tag
TAG=${1:-r01}
mkdir -p gpurun_out
(cd tools/ubench && ./tile_ubench) > gpurun_out/tile_ubench_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_select -s 1 -c 1 -f -o gpurun_out/prof_cfg4_${TAG} \
    python tools/run_once.py --config 4 --users 18944 --items 200000 --reps 2 > gpurun_out/prof_cfg4_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_select -s 1 -c 1 -f -o gpurun_out/prof_cfg3_${TAG} \
    python tools/run_once.py --config 3 --users 18944 --items 40000 --reps 2 > gpurun_out/prof_cfg3_${TAG}.log 2>&1
cat gpurun_out/tile_ubench_${TAG}.log; ls -la gpurun_out | tail -5
