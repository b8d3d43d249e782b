#!/bin/bash
# round 2: occupancy variants of k_trio_scan
cd "$(dirname "$0")/.."
mkdir -p genofix_b200/variants
rm -f genofix_b200/variants/lib_*.so genofix_b200/variants/build_*.log
for v in "$@"; do
  name=${v%%=*}; flags=${v#*=}
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -Xptxas -v ${flags//,/ } \
    -o genofix_b200/variants/lib_$name.so genofix_b200/csrc/gfx_kernels.cu genofix_b200/csrc/gfx_schedule.cpp \
    > genofix_b200/variants/build_$name.log 2>&1 &
done
wait
for f in genofix_b200/variants/build_*.log; do echo $f; grep -h -A3 "k_trio_scanPKj" $f | grep -E "Used|spill"; done
