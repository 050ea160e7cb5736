#!/bin/bash
# build_variant.sh NAME [-Dmacro=value ...] -> build/ab/libgoofy_NAME.so (experiments; same flags as goofy_b200/build.py)
name=$1; shift
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -cudart shared \
     -Xptxas -v "$@" -o build/ab/libgoofy_$name.so goofy_b200/csrc/capi.cu 2> build/ab/$name.ptxas.txt || { tail -20 build/ab/$name.ptxas.txt; exit 1; }
grep -A2 -E "encode_rows_kernelILi[012]ELb0|encode_direct_kernelILi0ELb0ELb0" build/ab/$name.ptxas.txt | grep -E "Used|spill" | tr '\n' ' '; echo
