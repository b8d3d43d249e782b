#!/bin/bash
# round 2: thread counts / task sizes of k_mendel_rank3.  usage: build_variants_rank3.sh name=flag,flag name=flag ...
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
for f in genofix_b200/variants/build_*.log; do echo $f; grep -h -A3 "k_mendel_rank3ILb1" $f | grep -E "Used|spill"; done
ls genofix_b200/variants/*.so
