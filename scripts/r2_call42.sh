#!/bin/bash
# Round 2, call 42: task pipeline without the proxy fence (MEMBAR.ALL.CTA per task): bit-for-bit check, A/B timing
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/bench_rank_forms.py > $O/r2c42_forms.log 2>&1
echo "rc $?" >> $O/r2c42_forms.log
grep -v "True  hist True\|True  hist -" $O/r2c42_forms.log | tail -8
echo "== no flush"; GFX_NOFLUSH=1 GFX_FORMS_TIME_ONLY=1 timeout 120 python scripts/bench_rank_forms.py 2>&1 | tail -2
for f in genofix_b200/variants/lib_*.so; do
  echo "== $f"
  GFX_FORMS_TIME_ONLY=1 GFX_LIB_PATH=$PWD/$f timeout 120 python scripts/bench_rank_forms.py 2>&1 | tail -2
done
