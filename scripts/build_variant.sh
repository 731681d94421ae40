#!/bin/bash
# Development aid: builds gpurun_variants/lib_<name>.so with extra -D flags for scan_kernels.cu
#   scripts/build_variant.sh name "-DMMG_STAGE_BYTES=4096u -DMMG_NSTAGES=2"
set -e
cd "$(dirname "$0")/../monkey-moore_b200/csrc"
NAME=$1; shift
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
OBJ=../../gpurun_variants/obj
mkdir -p $OBJ
for f in capi.cu comm.cu unique.cu pattern.cpp; do
  o=$OBJ/${f%.*}.o
  if [ ! -f $o ] || [ $f -nt $o ] || [ scan_kernels.cuh -nt $o ] || [ program.h -nt $o ]; then $NVCC $FLAGS -c -o $o $f; fi
done
$NVCC $FLAGS $@ -c -o $OBJ/scan_kernels_$NAME.o scan_kernels.cu
$NVCC $FLAGS -shared -o ../../gpurun_variants/lib_$NAME.so $OBJ/capi.o $OBJ/comm.o $OBJ/unique.o $OBJ/pattern.o $OBJ/scan_kernels_$NAME.o -lcudart -ldl
echo built lib_$NAME.so
