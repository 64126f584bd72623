#!/bin/bash
# small catalogues: where cfg2's filter kernel and cfg1's call spend their time
cd "$(dirname "$0")/.."
TAG=${1:-x}
mkdir -p gpurun_out
( for c in 1 2; do echo "== cfg$c"; timeout 120 python tools/run_once.py --config $c --users $([ $c = 1 ] && echo 6040 || echo 138493) --reps 3 2>&1 | tail -1 | cut -c1-420; done ) > gpurun_out/small_${TAG}.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:filter_select -c 1 -f -o gpurun_out/ncu_filter_cfg2_${TAG} \
    python tools/run_once.py --config 2 --users 138493 --reps 1 > gpurun_out/ncu_filter_cfg2_${TAG}.log 2>&1
cat gpurun_out/small_${TAG}.log; tail -2 gpurun_out/ncu_filter_cfg2_${TAG}.log
