#!/bin/bash
# tuning build of one translation unit: tools/build_variant.sh <name> <file.cu> "<-D flags>"  -> build/libeffimvs_<name>.so
# (select it with EFFIMVS_LIB=build/libeffimvs_<name>.so; build/ is git-ignored but travels to the GPU box)
set -e
name=$1; src=$2; flags=$3
cd "$(dirname "$0")/../effi-mvs-plus_b200/csrc"
mkdir -p ../../build
obj=../../build/${src%.cu}_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $flags -c $src -o $obj 2> ../../build/${src%.cu}_$name.ptxas.log
others=$(ls *.o | grep -v "^${src%.cu}.o$")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/libeffimvs_$name.so $obj $others -lcuda
grep -A2 "warp_corr_tile_kernelILi8ELi1\|warp_corr_tile_kernelILi16ELi1" ../../build/${src%.cu}_$name.ptxas.log | grep "Used\|spill" | tr '\n' ' '; echo
