#!/bin/bash
# launch list + ncu full capture of tensor-path kernels (cfg4 shape).  $1 = tag, $2 = kernel regex (default filter_select)
TAG=${1:-r01}
KREG=${2:-filter_select}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_filter_${TAG}.csv \
    python tools/run_once.py --config 4 --users 37888 --reps 2 > gpurun_out/launches_filter_${TAG}.log 2>&1
for K in $KREG; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${K}_cfg4_${TAG} \
    python tools/run_once.py --config 4 --users 18944 --items ${ITEMS:-400000} --reps 2 > gpurun_out/prof_${K}_cfg4_${TAG}.log 2>&1
done
