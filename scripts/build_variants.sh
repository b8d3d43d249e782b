#!/bin/bash
# kernel experiments: builds variants of the library with different -D settings into genofix_b200/variants/
cd "$(dirname "$0")/.."
build() { # name flags...
  name=$1; shift
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared "$@" \
    -o genofix_b200/variants/lib_$name.so genofix_b200/csrc/gfx_kernels.cu genofix_b200/csrc/gfx_schedule.cpp &
}
build base
build m64 -DGFX_MULTI_CHUNK=64
build m64w4 -DGFX_MULTI_CHUNK=64 -DGFX_MULTI_WARPS=4
build m128w4 -DGFX_MULTI_CHUNK=128 -DGFX_MULTI_WARPS=4
build m256 -DGFX_MULTI_CHUNK=256
wait
ls -la genofix_b200/variants
