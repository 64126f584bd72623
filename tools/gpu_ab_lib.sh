#!/bin/bash
# A/B of a build variant (build_variants/$1.so, loaded through RMB200_LIB) against the default build on one configuration
cd "$(dirname "$0")/.."
V=${1:-alu}; CFG=${2:-3}; USERS=${3:-75776}
for rep in 1 2; do
  echo "== default"; timeout 100 python tools/run_once.py --config $CFG --users $USERS --reps 3 2>&1 | tail -1 | cut -c1-110
  echo "== $V"; RMB200_LIB=$PWD/build_variants/$V.so timeout 100 python tools/run_once.py --config $CFG --users $USERS --reps 3 2>&1 | tail -1 | cut -c1-110
done
RMB200_LIB=$PWD/build_variants/$V.so timeout 200 python -m pytest tests/test_gpu_ties.py -x -q -m gpu 2>&1 | tail -1
