#!/bin/sh
# Kernel-variant builds of the library for A/B runs (tools/ab_r02.py loads them through RT_B200_LIB).  Not part of the product.
set -e
cd "$(dirname "$0")/../distributed-raytracer_b200/csrc"
mkdir -p ../../build_variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off -shared -cudart static"
SRC="rt_api.cu rt_group.cu render.cu bvh_build.cu obj_ingest.cu wire.cu"
$NV -DRT_LD256=0 -o ../../build_variants/librt_b200_ld128.so $SRC &
$NV -DRT_LD256=0 -DRT_OCTANT=0 -o ../../build_variants/librt_b200_nooct.so $SRC &
wait
$NV -DRT_SMEM_STACK=1 -o ../../build_variants/librt_b200_smemstack.so $SRC &
$NV -DRT_TRI_LD256=0 -o ../../build_variants/librt_b200_tri48.so $SRC &
$NV -DRT_FUSED_THREADS=64 -o ../../build_variants/librt_b200_cta64.so $SRC &
wait
$NV -DRT_ANY_DEFER=0 -o ../../build_variants/librt_b200_nodefer.so $SRC &
$NV -DRT_PUSH_PREFETCH=1 -o ../../build_variants/librt_b200_pf1.so $SRC &
$NV -DRT_PUSH_PREFETCH=2 -o ../../build_variants/librt_b200_pf2.so $SRC &
wait
$NV -DRT_CHUNK_BLOCKS=16 -o ../../build_variants/librt_b200_chunk16.so $SRC &
$NV -DRT_CHUNK_BLOCKS=64 -o ../../build_variants/librt_b200_chunk64.so $SRC &
$NV -DRT_CHUNK_BLOCKS=128 -o ../../build_variants/librt_b200_chunk128.so $SRC &
wait
$NV -DRT_NODEQ=0 -o ../../build_variants/librt_b200_noq.so $SRC
$NV -DRT_TRI_NOALLOC=0 -o ../../build_variants/librt_b200_trialloc.so $SRC &
$NV -DRT_NODE_EVICT_LAST=1 -o ../../build_variants/librt_b200_nodeel.so $SRC &
$NV -DRT_HIT_STREAM=1 -o ../../build_variants/librt_b200_hitstream.so $SRC &
wait
# $NV -DRT_W=8 -o ../../build_variants/librt_b200_w8.so $SRC      # 8 children per node: measured 35-70 % slower (profiles/r02_kernel_ab.txt)
