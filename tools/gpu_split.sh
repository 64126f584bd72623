#!/bin/bash
# GPU check of the splitters: parity tests, one memcheck pass on a small case, timings beside the reference, and smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_split.py -x -q -s 2>&1 | tail -25
echo "== memcheck (small separated split)"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import numpy as np
from recometrics_b200 import _capi
from tools import split_cases
for name in ('split_separated_f64', 'split_joined_unsorted_f32', 'split_all_unsorted_f64'):
    mk, kw = split_cases.CASES[name]; kw = dict(kw)
    p, i, v = split_cases.make_csr(**mk)
    r = _capi.split(kw.pop('split_type'), p, i, v, mk['m'], mk['n'], **kw)
    print(name, r['timing']['kernel_launches'])
" 2>&1 | tail -8
echo "== timings"
for kind in all separated joined; do timeout 200 python tools/split_once.py --kind $kind 2>&1 | tail -1 | tee -a gpurun_out/split_once.jsonl; done
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
