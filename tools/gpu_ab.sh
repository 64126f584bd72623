#!/bin/bash
# developer: A/B of library builds in build_variants/ on the same box (cfg4, one batch), plain and RMB200_DBG ceilings
( for so in build_variants/v*.so; do for e in ${ENVS:-RMB200_DBG=0 RMB200_DBG=1}; do
  echo "== $(grep "^$(basename $so .so):" build_variants/list.txt) $e"
  env $e RMB200_LIB=$PWD/$so python tools/run_once.py --config ${CFG:-4} --users ${USERS:-151552} --reps 3 2>&1 | tail -1 | cut -c1-60
done; done ) 2>&1 | tee gpurun_out/ab.log
