#!/bin/bash
# kernel experiments: builds variants of the library with different -D settings into genofix_b200/variants/
cd "$(dirname "$0")/.."
build() { # name flags...
  name=$1; shift
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared "$@" \
    -o genofix_b200/variants/lib_$name.so genofix_b200/csrc/gfx_kernels.cu genofix_b200/csrc/gfx_schedule.cpp &
}
build base
build t512 -DR2_THREADS=512
build t768 -DR2_THREADS=768
wait
ls -la genofix_b200/variants
