#!/bin/bash
# Round 2, call 46: stall samples per source line of the three correction kernels (generation 4 of a 10 000 x 10 000 chunk)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/bench_correct.py 10000 > $O/r2c46_correct.log 2>&1; tail -1 $O/r2c46_correct.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_correct$|^k_correct_jobs" -s 54 -c 3 -o $O/r2c46_full_correct -f \
    python scripts/bench_correct.py 10000 > $O/r2c46_ncu.log 2>&1
ls -la $O/r2c46_full_correct.ncu-rep
for k in 'k_correct$' 'k_correct_jobs$' 'k_correct_jobs_general'; do
  python scripts/ncu_stalls.py $O/r2c46_full_correct.ncu-rep "$k" 70 > "$O/r2c46_stalls_${k%\$}.txt" 2>&1
done
python scripts/ncu_summary.py $O/r2c46_full_correct.ncu-rep > $O/r2c46_summary.txt 2>&1
head -30 $O/r2c46_summary.txt
