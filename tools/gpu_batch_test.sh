#!/bin/bash
# users per batch (RMB200_BATCH_USERS): the default against the previous one, on the bench configurations
cd "$(dirname "$0")/.."
for cfg in ${CFGS:-4 3 5 2}; do for ub in 0 151552; do echo "== cfg$cfg RMB200_BATCH_USERS=$ub (0 = default)"; RMB200_BATCH_USERS=$ub timeout 200 python bench.py --config $cfg --no-cpu-baseline --steps 3 --warmup 2 2>&1 | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'kernel %.2f'%d['roofline']['kernel_ms_per_step'])"; done; done
