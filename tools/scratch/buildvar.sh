#!/bin/bash
# usage: buildvar.sh name [-D...]   -> tools/scratch/libs/name.so from the working tree
cd /root/repo
n=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O2 --shared "$@" pantax_b200/csrc/ptx_kernels.cu pantax_b200/csrc/ptx_api.cu -o tools/scratch/libs/$n.so -lcudart -ldl
