#!/bin/bash
# tools/build_variant.sh NAME SRC.cu "-DFLAG ..." : rebuild one translation unit of libfdb200 with extra flags and link
# firedrake_b200/lib/variants/libfdb200_NAME.so (select it with FDB200_LIB=...; A/B timing on the GPU box)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; SRC=$2; FLAGS=$3
OUT=$ROOT/firedrake_b200/lib/variants; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas=-v $FLAGS \
  -c $ROOT/firedrake_b200/csrc/$SRC -o $OUT/${SRC%.cu}_$NAME.o 2> $OUT/${NAME}_ptxas.log
OBJS=$(ls $ROOT/firedrake_b200/lib/*.o | grep -v "/${SRC%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libfdb200_$NAME.so $OBJS $OUT/${SRC%.cu}_$NAME.o -ldl
grep -A2 "helmholtz_action_kernelILi4ELb0ELb1ELi3ELb0ELb0ELb0E" $OUT/${NAME}_ptxas.log | grep -E "registers|spill" | head -2
