#!/bin/bash
# Round 2, call 52: deferred job items carry what is already known about them (no repeated failing walks)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "correct_matrix or exact_integer or huge_litter or very_wide or config4_deep or config1_example" > $O/r2c52_pytest.log 2>&1
tail -3 $O/r2c52_pytest.log
for v in prev main prev main; do
  echo "== $v"
  GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_$v.so timeout 300 python scripts/bench_correct.py 10000 2>&1 | tail -1
done | tee $O/r2c52_variants.log
