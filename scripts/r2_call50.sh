#!/bin/bash
# Round 2, call 50: fast kernel variants (grand-parents requested with the parents, blocks per SM, 64-SNP tasks); launch list of the stage
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in main eager kc5 kc8 lean64 main eager; do
  echo "== $v"
  GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_$v.so timeout 300 python scripts/bench_correct.py 10000 2>&1 | tail -1
done | tee $O/r2c50_variants.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 400 --csv --log-file $O/r2c50_launches.csv \
   python scripts/bench_correct.py 10000 > $O/r2c50_ncu.log 2>&1
python scripts/launch_agg.py $O/r2c50_launches.csv > $O/r2c50_launches.txt 2>&1
head -14 $O/r2c50_launches.txt
