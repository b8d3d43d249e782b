#!/bin/bash
# Round 2, call 25: k_mendel_rank3 with the work list and (class, m < 64) counters: bit-for-bit against k_mendel_rank, timings,
# then the whole GPU test suite
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/bench_rank_forms.py > $O/r2c25_forms.log 2>&1
echo "rc $?" >> $O/r2c25_forms.log
grep -v "True  hist True\|True  hist -" $O/r2c25_forms.log | tail -12
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c25_pytest.log 2>&1
tail -5 $O/r2c25_pytest.log
