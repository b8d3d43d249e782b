#!/bin/bash
# round 2 kernel experiments on the correction stage: occupancy of the deferred-job kernels, ticket sizes
cd "$(dirname "$0")/.."
mkdir -p genofix_b200/variants
rm -f genofix_b200/variants/lib_*.so
build() { # name flags...
  name=$1; shift
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -Xptxas -v "$@" \
    -o genofix_b200/variants/lib_$name.so genofix_b200/csrc/gfx_kernels.cu genofix_b200/csrc/gfx_schedule.cpp \
    > genofix_b200/variants/build_$name.log 2>&1 &
}
build base
build kj8 -DKJ_BLOCKS=8
build kj6 -DKJ_BLOCKS=6
build kg8 -DKG_BLOCKS=8
build kg6 -DKG_BLOCKS=6
build kc8 -DKC_BLOCKS=8
build kgi32 -DKG_ITEMS=32
build kgi8 -DKG_ITEMS=8
wait
grep -h -E "k_correct_jobs|k_correct_jobs_general|k_correctILi32" -A2 genofix_b200/variants/build_*.log | grep -E "Compiling|Used" | head -40
ls genofix_b200/variants
