#!/bin/bash
# Round 2, call 15: k_mendel_rank3 against k_mendel_rank (bit for bit), timings, thread-count variants
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python scripts/bench_rank_forms.py > $O/r2c15_forms.log 2>&1
echo "rc $?" >> $O/r2c15_forms.log
cat $O/r2c15_forms.log
for f in genofix_b200/variants/lib_t*.so; do
  echo "== $f"
  GFX_FORMS_TIME_ONLY=1 GFX_LIB_PATH=$PWD/$f timeout 300 python scripts/bench_rank_forms.py 2>&1 | tail -2
done > $O/r2c15_variants.log 2>&1
cat $O/r2c15_variants.log
