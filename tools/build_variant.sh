#!/bin/bash
# A/B builds of the library: tools/build_variant.sh NAME -DFLAG=... -> variants/libmstcn_NAME.so
# (git-ignored like every .so, travels with gpurun; select with MSTCN_B200_LIB=variants/libmstcn_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -DMSTCN_WITH_TC "$@" \
  -o variants/libmstcn_$name.so pytorch_video_action_b200/csrc/mstcn_capi.cu
echo built variants/libmstcn_$name.so
