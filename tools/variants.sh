#!/bin/bash
# Time build variants of the library on cfg4 / cfg5 (developer tool).  Variants are built on the CPU box into
# build_variants/ by `tools/variants.sh build`, then timed on the GPU box by `tools/variants.sh run`.
cd "$(dirname "$0")/.."
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3,-fvisibility=hidden -shared -cudart static"
if [ "$1" = build ]; then
  mkdir -p build_variants; rm -f build_variants/*.so
  shift
  i=0
  for v in "$@"; do
    $NVCC $FLAGS $v -o build_variants/v$i.so recometrics_b200/csrc/api.cu recometrics_b200/csrc/full_order.cu & 
    echo "v$i: $v" >> build_variants/list.txt.tmp
    i=$((i+1))
  done
  wait; mv build_variants/list.txt.tmp build_variants/list.txt; cat build_variants/list.txt
else
  mkdir -p gpurun_out
  for so in build_variants/v*.so; do
    for cfg in ${CFGS:-4 5}; do
      echo "== $so cfg$cfg $(grep "^$(basename $so .so):" build_variants/list.txt)"
      RMB200_LIB=$PWD/$so timeout 600 python tools/run_once.py --config $cfg --users ${USERS:-37888} --reps 3 2>&1 | tail -1 | cut -c1-120
    done
  done | tee gpurun_out/variants.log
fi
