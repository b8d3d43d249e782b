#!/bin/bash
# Round 2, call 54: --set full + stall samples per source line of the three correction kernels (generation 4, shipped build)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_correct$|^k_correct_jobs" -s 66 -c 3 -o $O/r2c54_full_correct_gen4 -f \
    python scripts/bench_correct.py 10000 > $O/r2c54_ncu.log 2>&1
for k in 'k_correct$' 'k_correct_jobs$' 'k_correct_jobs_general'; do
  python scripts/ncu_stalls.py $O/r2c54_full_correct_gen4.ncu-rep "$k" 50 > "$O/r2c54_stalls_gen4_${k%\$}.txt" 2>&1
done
python scripts/ncu_summary.py $O/r2c54_full_correct_gen4.ncu-rep > $O/r2c54_summary_gen4.txt 2>&1
grep -E "Kernel Name|time_duration|inst_executed.sum|issue_active|warps_active|registers" $O/r2c54_summary_gen4.txt
