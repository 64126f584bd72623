#!/bin/bash
# First GPU visit: environment facts, smoke, parity tests, a short bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
{
  echo "== host"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|Flags" | cut -c1-300; free -g | head -2
  echo "== gpu"; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,memory.total --format=csv
} > gpurun_out/env.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --users 151552 --steps 2 --warmup 1 > gpurun_out/bench_small.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_small.log
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_small.log
