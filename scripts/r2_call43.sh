#!/bin/bash
# Round 2, call 43: k_mendel_rank3 thread-count / prefetch variants (timing only)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for f in genofix_b200/variants/lib_*.so; do
  echo "== $f"
  GFX_FORMS_TIME_ONLY=1 GFX_LIB_PATH=$PWD/$f timeout 120 python scripts/bench_rank_forms.py 2>&1 | tail -2
done > $O/r2c43_variants.log 2>&1
cat $O/r2c43_variants.log
