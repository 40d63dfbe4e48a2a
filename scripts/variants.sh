#!/bin/bash
# Development aid: build libvkp variants with different tuning macros into build/variants/<name>/libvkp.so
# usage: scripts/variants.sh name "-DVKP_LB_SHADOW=8 ..." [name2 "flags2" ...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/vulkanhybridraytracer_b200/csrc
while [ $# -ge 2 ]; do
  NAME=$1; FLAGS=$2; shift 2
  OUT=$ROOT/variants/$NAME; mkdir -p $OUT
  for f in vkp_api bvh_build kernels_frame vkp_group; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -ccbin /usr/bin/g++ \
      --cudart static -Xptxas -v -diag-suppress 20012 $FLAGS -c $CS/$f.cu -o $OUT/$f.o 2> $OUT/$f.log &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -ccbin /usr/bin/g++ -o $OUT/libvkp.so $OUT/*.o -lpthread -ldl -lrt
  grep -E "Used [0-9]+ registers|spill" $OUT/kernels_frame.log | paste - - | grep -v " 0 bytes spill stores" | head -5 || true
  echo "built $NAME: $FLAGS"
done
