#!/bin/bash
# build_variant.sh NAME "-DMARCH_WARPS2=12 ..." [src]  -> variants/NAME.so (src.cu rebuilt with the flags; default ipdg)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
L=stefanproblem_b200/lib
SRC=${3:-ipdg}
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas=-v $2 \
  -c stefanproblem_b200/csrc/$SRC.cu -o variants/$1.o 2> variants/$1.ptxas.log
objs=$(ls $L/*.o | grep -v "/$SRC.o")
/usr/local/cuda/bin/nvcc -shared -o variants/$1.so variants/$1.o $objs -lcudart 2>/dev/null
rm variants/$1.o
grep -A2 "apply_march" variants/$1.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
