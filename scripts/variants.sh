#!/bin/bash
# Build one library per kernel-experiment setting into gpurun_out/variants/ (run HERE, then bench on the GPU box
# with GJ_LIB_PATH=...).   usage: scripts/variants.sh name1:"-DGJ_X=0 -DGJ_Y=1" name2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 --extended-lambda \
      -Xcompiler -fPIC -I include -I gradabm-june_b200/csrc $flags -c -o variants/k_$name.o gradabm-june_b200/csrc/gj_kernels.cu && \
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o variants/lib_$name.so \
      variants/k_$name.o gradabm-june_b200/build/gj_world.o ) &
done
wait
ls -la variants
