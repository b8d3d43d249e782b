#!/bin/bash
# Round 2, call 36: k_mendel_rank3 on the childless rows alone and on the rows with children alone (what would two co-resident
# kernels with their own register budgets have to beat?)
cd "$(dirname "$0")/.."
for v in 0 1 2; do
  echo "== GFX_RANK_ONLY=$v"; GFX_RANK_ONLY=$v GFX_FORMS_TIME_ONLY=1 timeout 120 python scripts/bench_rank_forms.py 2>&1 | tail -2
done
