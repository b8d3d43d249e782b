#!/bin/bash
# kernel experiments: builds variants of the library with different -D settings into genofix_b200/variants/
cd "$(dirname "$0")/.."
build() { # name flags...
  name=$1; shift
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared "$@" \
    -o genofix_b200/variants/lib_$name.so genofix_b200/csrc/gfx_kernels.cu genofix_b200/csrc/gfx_schedule.cpp &
}
build base
build g8i16 -DKG_BLOCKS=8 -DKG_ITEMS=16
build g6i16 -DKG_BLOCKS=6 -DKG_ITEMS=16
build g8i8 -DKG_BLOCKS=8 -DKG_ITEMS=8
build g4i16 -DKG_ITEMS=16
wait
ls -la genofix_b200/variants
