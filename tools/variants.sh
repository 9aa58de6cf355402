#!/bin/bash
# Builds one librt_b200 variant per experiment setting into build_variants/ (git-ignored, travels with gpurun).
# usage: tools/variants.sh name "-DRT_TRAV_MODE=0 ..." [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/../distributed-raytracer_b200/csrc"
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  d=../../build_variants/$name; mkdir -p $d
  for f in rt_api render bvh_build; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a $flags -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off -c $f.cu -o $d/$f.o &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $d/librt_b200.so $d/rt_api.o $d/render.o $d/bvh_build.o
  rm -f $d/*.o
  echo built $name
done
