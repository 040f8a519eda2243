#!/bin/bash
# A/B build of ONE translation unit with extra -D flags, linked with the product's other objects:
#   scratch/build_variant.sh setdata_tma.cu plainst '-DMT_TMA_ST(p,v)=(*(p)=(v))'   -> scratch/variants/fastcompute_plainst.so
set -e
cd "$(dirname "$0")/.."
src=$1; tag=$2; shift 2
mkdir -p scratch/variants
obj=scratch/variants/${src%.cu}_$tag.o
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC \
  -I include -I meteotools_b200/csrc "$@" -c meteotools_b200/csrc/$src -o $obj
others=$(ls meteotools_b200/lib/obj/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/variants/fastcompute_$tag.so $obj $others
echo scratch/variants/fastcompute_$tag.so
