#!/bin/bash
# developer: sampling on/off + pipeline ceilings of the filter kernel (RMB200_DBG bits), cfg4 one batch
( timeout 600 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -3
for e in "RMB200_SAMPLE=0" "RMB200_SAMPLE=1" "RMB200_SAMPLE=1 RMB200_SAMPLE_STRIDE=8" "RMB200_SAMPLE=1 RMB200_SAMPLE_STRIDE=32" "RMB200_DBG=9" "RMB200_DBG=1"; do echo "== $e"; env $e python tools/run_once.py --config 4 --users ${USERS:-151552} --reps 3 2>&1 | tail -1 | cut -c1-90; env $e python tools/run_once.py --config 4 --users ${USERS:-151552} --reps 3 2>&1 | tail -1 | grep -o '"retry_rows.*'; done
for e in "RMB200_SAMPLE=0" "RMB200_SAMPLE=1"; do echo "== cfg5 $e"; env $e python tools/run_once.py --config 5 --users ${USERS:-151552} --reps 3 --f64 2>&1 | tail -1 | cut -c1-90; done ) 2>&1 | tee gpurun_out/dbg.log
