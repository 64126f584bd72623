#!/bin/bash
# cfg4 tuning sweep at the 4-wave batch: sample stride (environment) x cut margin of the main pass (build variants)
cd "$(dirname "$0")/.."
run() { timeout 100 python tools/run_once.py --config 4 --users 303104 --reps 2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('   dom_ms %.2f kernel_ms %.2f retry %s fallback %s' % (d['dom_ms'], d['kernel_ms'], d['retry_rows'], d['fallback']))"; }
for rep in 1 2; do
echo "== default"; run
for so in build_variants/*.so; do for s in 14 10; do echo "== $so stride $s"; RMB200_SAMPLE_STRIDE=$s RMB200_LIB=$PWD/$so run; done; done
done
