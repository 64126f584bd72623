#!/bin/bash
# developer: ring depth sensitivity of the filter kernel (cfg4, one batch)
( for so in build_variants/v0.so build_variants/v1.so; do for st in 3 4 5; do
  echo "== $(grep "^$(basename $so .so):" build_variants/list.txt) RMB200_STAGES=$st"
  RMB200_STAGES=$st RMB200_LIB=$PWD/$so timeout 180 python tools/run_once.py --config 4 --users 151552 --reps 3 2>&1 | tail -1 | cut -c1-60
done; done ) 2>&1 | tee gpurun_out/stages.log
