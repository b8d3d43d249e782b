#!/bin/bash
# kernel experiments: builds variants of the library with different -D settings into genofix_b200/variants/
cd "$(dirname "$0")/.."
build() { # name flags...
  name=$1; shift
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared "$@" \
    -o genofix_b200/variants/lib_$name.so genofix_b200/csrc/gfx_kernels.cu genofix_b200/csrc/gfx_schedule.cpp &
}
build base
build f4b4 -DGFX_FAM_BLOCKS=4
build f3b4 -DGFX_FAM_KIDS=3 -DGFX_FAM_BLOCKS=4
build f4b5 -DGFX_FAM_BLOCKS=5
wait
ls -la genofix_b200/variants
