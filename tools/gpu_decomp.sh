#!/bin/bash
# developer: decompose the filter kernel's tile time with the RMB200_DBG bits of a -DRMB_F_DBG=1 build (build_variants/v0.so)
# 1 = no scan, 8 = no tcgen05.ld, 16 = no TMA after the first ring round, 32 = half of the k steps, 2 = fast path only, 4 = no meetings
( for e in 0 1 9 17 25 33 57 2 4; do
  echo "== RMB200_DBG=$e"
  RMB200_DBG=$e RMB200_LIB=$PWD/build_variants/v0.so python tools/run_once.py --config ${CFG:-4} --users ${USERS:-151552} --reps 3 2>&1 | tail -1 | cut -c1-60
done ) 2>&1 | tee gpurun_out/decomp.log
