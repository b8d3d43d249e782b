#!/bin/bash
# Round 2, call 16: full ncu capture of k_mendel_rank3<true> on the config 2 chunk
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
GFX_FORMS_TIME_ONLY=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_mendel_rank3" -s 4 -c 1 -o $O/r2c16_full_rank3 -f \
  python scripts/bench_rank_forms.py > $O/r2c16_ncu.log 2>&1
tail -3 $O/r2c16_ncu.log
ls -la $O/r2c16_full_rank3.ncu-rep
