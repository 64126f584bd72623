#!/bin/bash
# full GPU suite + e2e of two configurations (after a host-pipeline change)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
for cfg in 5 4; do timeout 300 python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cfg$cfg value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pageable', round(d['e2e_pageable']['value']), 'ms', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2))"; done
