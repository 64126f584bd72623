#!/bin/bash
# developer: filter kernel time against the number of appended candidates (k_metrics sweep at the cfg4 shape; stats build = v0)
( for k in 1 10 30 100; do
  echo "== production k=$k"; timeout 180 python tools/run_once.py --config 4 --users 151552 --k $k --reps 3 2>&1 | tail -1 | cut -c1-60
  echo "== stats build k=$k"; RMB200_LIB=$PWD/build_variants/v0.so timeout 180 python tools/run_once.py --config 4 --users 151552 --k $k --reps 2 2>&1 | grep "rmb200 stats" | tail -1
done ) 2>&1 | tee gpurun_out/ksweep.log
