#!/bin/bash
# kernel experiments: builds variants of the library with different -D settings into genofix_b200/variants/
# (git-ignored; they travel to the GPU box).  scripts/bench_rank_variants.py times the rank kernel of every variant in one
# process and compares scores + histogram bit for bit with lib_base.so; scripts/run_variants.sh runs any other script per variant.
cd "$(dirname "$0")/.."
mkdir -p genofix_b200/variants
build() { # name flags...
  name=$1; shift
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -Xptxas -v "$@" \
    -o genofix_b200/variants/lib_$name.so genofix_b200/csrc/gfx_kernels.cu genofix_b200/csrc/gfx_schedule.cpp \
    > genofix_b200/variants/build_$name.log 2>&1 &
}
build base
# threads per SM of the rank kernel (registers per thread = 65536 / threads, rounded down to a multiple of 8)
build t1024 -DR2_THREADS=1024
build t768 -DR2_THREADS=768
# task lengths in words: rows with >= 8 children / 1..7 children / childless
build H64 -DR2_WORDS_HEAVY=64
build L256 -DR2_WORDS_LIGHT=256
wait
grep -h -A2 "k_mendel_rankILb1" genofix_b200/variants/build_*.log | grep "Used"
ls -la genofix_b200/variants
