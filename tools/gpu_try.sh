#!/bin/bash
# developer: quick timing of the in-tree build (hang guard), then the GPU parity suite.  CFGS / USERS as in gpu_check.sh
( for cfg in ${CFGS:-4 5 2}; do
    echo "== cfg$cfg"; timeout 180 python tools/run_once.py --config $cfg --users ${USERS:-151552} --reps 3 2>&1 | tail -1 | cut -c1-${CUT:-100}
  done
  timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
) 2>&1 | tee gpurun_out/try_${1:-x}.log
