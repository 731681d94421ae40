#!/bin/bash
# Development aid: builds gpurun_variants/lib_<name>.so with extra -D flags (applied to every source)
#   scripts/build_variant.sh name "-DMMG_NCTR=32u"
set -e
cd "$(dirname "$0")/../monkey-moore_b200/csrc"
NAME=$1; shift
mkdir -p ../../gpurun_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $@ -shared \
  -o ../../gpurun_variants/lib_$NAME.so capi.cu scan_kernels.cu unique.cu comm.cu pattern.cpp -lcudart -ldl
echo built lib_$NAME.so
