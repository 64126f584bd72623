#!/bin/bash
# developer: host -> device upload of pageable numpy memory, pipeline on/off (cfg4 one batch; h2d_ms of tools/run_once.py)
( for t in 0 2 4 8; do echo "== RMB200_UPLOAD_THREADS=$t"
  RMB200_UPLOAD_THREADS=$t timeout 300 python tools/run_once.py --config 4 --users 151552 --reps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('h2d_ms', round(d['h2d_ms'],1), 'total_ms', round(d['total_ms'],1), 'kernel_ms', round(d['kernel_ms'],1))"
done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/upload.log
