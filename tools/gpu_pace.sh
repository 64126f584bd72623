#!/bin/bash
# developer: the epilogue's own pace -- cfg4 with fewer factors (cheaper tiles: TMA bytes and MMA k steps shrink, the scan does not)
( for p in 128 64 32; do
  echo "== production p=$p"; python tools/run_once.py --config 4 --users 151552 --factors $p --reps 3 2>&1 | tail -1 | cut -c1-60
  echo "== debug build RMB200_DBG=1 (no scan) p=$p"; RMB200_DBG=1 RMB200_LIB=$PWD/build_variants/v0.so python tools/run_once.py --config 4 --users 151552 --factors $p --reps 3 2>&1 | tail -1 | cut -c1-60
  echo "== debug build RMB200_DBG=2 (fast path only) p=$p"; RMB200_DBG=2 RMB200_LIB=$PWD/build_variants/v0.so python tools/run_once.py --config 4 --users 151552 --factors $p --reps 3 2>&1 | tail -1 | cut -c1-60
done ) 2>&1 | tee gpurun_out/pace.log
