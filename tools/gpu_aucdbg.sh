#!/bin/bash
# where the rank-counting kernel's time goes: RMB200_AUC_DBG switches parts off (results wrong, times only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( for e in ${DBGS:-0 1 2 4}; do echo "== RMB200_AUC_DBG=$e"; RMB200_AUC_DBG=$e timeout 100 python tools/run_once.py --config 3 --users 75776 --reps 3 2>&1 | tail -3 | cut -c1-230; done
echo "== top-K only on the FMA tiles (no counting warps: the plain kernel)"; RMB200_PATH=fma timeout 100 python tools/run_once.py --config 3 --users 75776 --reps 3 --no-auc 2>&1 | tail -1 | cut -c1-200 ) > gpurun_out/aucdbg_${1:-x}.log 2>&1
cat gpurun_out/aucdbg_${1:-x}.log
