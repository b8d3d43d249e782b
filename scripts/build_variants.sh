#!/bin/bash
# kernel experiments: builds variants of the library with different -D settings into genofix_b200/variants/
cd "$(dirname "$0")/.."
build() { # name flags...
  name=$1; shift
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared "$@" \
    -o genofix_b200/variants/lib_$name.so genofix_b200/csrc/gfx_kernels.cu genofix_b200/csrc/gfx_schedule.cpp &
}
build base
build xfast -DR2_X_NO_FAST_ATOMS
build xkids -DR2_X_NO_KIDS_ATOMS
build xbig -DR2_X_NO_BIG
build xepi -DR2_X_NO_EPILOGUE
build xall -DR2_X_NO_FAST_ATOMS -DR2_X_NO_KIDS_ATOMS -DR2_X_NO_BIG -DR2_X_NO_EPILOGUE
wait
ls -la genofix_b200/variants
