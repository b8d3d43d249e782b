#!/bin/bash
# Round 2, call 3: tests of the integer fast path, counters of a debug build, full ncu capture of the correction kernels
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c3_pytest.log 2>&1; tail -3 $O/r2c3_pytest.log
mkdir -p genofix_b200/variants
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -DGFX_DEBUG_COUNTERS \
   -o genofix_b200/variants/libdbg.so genofix_b200/csrc/gfx_kernels.cu genofix_b200/csrc/gfx_schedule.cpp > $O/r2c3_dbgbuild.log 2>&1
GFX_LIB_PATH=$PWD/genofix_b200/variants/libdbg.so python scripts/debug_counters.py > $O/r2c3_counters.log 2>&1; tail -12 $O/r2c3_counters.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_correct$|k_correct_jobs" -s 18 -c 18 -o $O/r2c3_full_correct -f \
    python scripts/bench_correct.py 4096 > $O/r2c3_ncu_correct.log 2>&1
ls -la $O/r2c3_full_correct.ncu-rep
