#!/bin/bash
# Round 2, call 23: boolean-network k_mendel_rank3 against k_mendel_rank (bit for bit), timings, variants, full ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/bench_rank_forms.py > $O/r2c23_forms.log 2>&1
echo "rc $?" >> $O/r2c23_forms.log
grep -v "True  hist True\|True  hist -" $O/r2c23_forms.log | tail -20
for f in genofix_b200/variants/lib_*.so; do
  echo "== $f"
  GFX_FORMS_TIME_ONLY=1 GFX_LIB_PATH=$PWD/$f timeout 120 python scripts/bench_rank_forms.py 2>&1 | tail -2
done > $O/r2c23_variants.log 2>&1
cat $O/r2c23_variants.log
GFX_FORMS_TIME_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_mendel_rank3" -s 4 -c 1 -o $O/r2c23_full_rank3 -f \
  python scripts/bench_rank_forms.py > $O/r2c23_ncu.log 2>&1
tail -2 $O/r2c23_ncu.log
