#!/bin/bash
# Round 2, call 49: one-pass table products in the general elimination; items per ticket / blocks per SM of the general kernel again
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "correct_matrix or exact_integer or huge_litter or very_wide or config4_deep" > $O/r2c49_pytest.log 2>&1
tail -3 $O/r2c49_pytest.log
for v in tab3v1 main kgi8 kgi32 kg6 kg3 kji16 main; do
  echo "== $v"
  GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_$v.so timeout 300 python scripts/bench_correct.py 10000 2>&1 | tail -1
done | tee $O/r2c49_variants.log
