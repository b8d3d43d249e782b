#!/bin/bash
# Round 2, call 48: ablation of the correction-kernel changes (one switch at a time), correction stage per 10 000 x 10 000 chunk
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "correct_matrix or exact_integer or huge_litter or very_wide or config4_deep" > $O/r2c48_pytest.log 2>&1
tail -3 $O/r2c48_pytest.log
for v in base all0 tab3 jt tk agg ku mrow all1 base all1; do
  echo "== $v"
  GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_$v.so timeout 300 python scripts/bench_correct.py 10000 2>&1 | tail -1
done | tee $O/r2c48_variants.log
