#!/bin/bash
# developer: sampled-guess parameters of the filter (cfg4, one batch): stride of the sample pass and rank of the guess
( for e in "RMB200_SAMPLE=1" "RMB200_SAMPLE_RANK=21" "RMB200_SAMPLE_RANK=19" "RMB200_SAMPLE_STRIDE=10" "RMB200_SAMPLE_STRIDE=20" "RMB200_SAMPLE_STRIDE=28" "RMB200_SAMPLE_STRIDE=20 RMB200_SAMPLE_RANK=16"; do
  echo "== $e"; env $e timeout 180 python tools/run_once.py --config 4 --users 151552 --reps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['dom_ms'], 'retry_rows', d['retry_rows'], 'fallback', d['fallback'])"
done ) 2>&1 | tee gpurun_out/sample_sweep.log
