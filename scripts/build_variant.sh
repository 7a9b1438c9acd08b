#!/bin/bash
# build_variant.sh NAME [-D...]: compile a variant of libsgc_b200.so into build/variants/NAME.so (kernel experiments)
NAME=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC -Xcompiler -pthread -lz "$@" \
  -o build/variants/$NAME.so spacegraphcats_b200/csrc/sgc_b200.cu
