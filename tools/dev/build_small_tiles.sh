#!/bin/bash
# Build-time experiment: the library with half-size BA tiles (MSFM_K2_SMALL_TILES=1: 256 observations / 24 local cameras per tile,
# two CTAs of the linearisation kernel per SM) as build/variants/libmsfm_b200_small.so; select it with MSFM_B200_LIB.
set -e
mkdir -p build/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
objs=""
for f in ba_api ba_kernels ba_band ba_solver; do
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -DMSFM_K2_SMALL_TILES=1 \
      -c monocularsfm_b200/csrc/$f.cu -o build/variants/${f}_small.o 2> build/variants/${f}_small.ptxas.log
  objs="$objs build/variants/${f}_small.o"
done
rest=$(ls build/obj/*.o | grep -v -E "/(ba_api|ba_kernels|ba_band|ba_solver)\.o")
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libmsfm_b200_small.so $objs $rest -lcusolver -lcublas -ldl -Xlinker -rpath=/usr/local/cuda/lib64
grep -A2 "fused_linearize_kernelILb0" build/variants/ba_kernels_small.ptxas.log | grep -E "registers|spill"
ls -la build/variants/libmsfm_b200_small.so
