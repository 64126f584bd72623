#!/bin/bash
# last GPU pass of the round: ncu capture of the splitters' partition kernel, the full GPU suite, the default bench, smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:partition_kernel -c 3 -f -o gpurun_out/r02_split_partition \
    python tools/split_once.py --no-reference --reps 1 > gpurun_out/r02_split_ncu.log 2>&1
tail -2 gpurun_out/r02_split_ncu.log | cut -c1-300
( time timeout 400 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r02end.txt 2>&1
tail -4 gpurun_out/pytest_gpu_r02end.txt
( time timeout 200 python bench.py ) > gpurun_out/bench_cfg4_r02end.log 2>&1
grep '^{' gpurun_out/bench_cfg4_r02end.log | cut -c1-300; grep -o '"split": {.*' gpurun_out/bench_cfg4_r02end.log | cut -c1-700
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02end.txt 2>&1; tail -4 gpurun_out/smoke_r02end.txt
