#!/bin/bash
# Round 2, call 47: correction kernels with prefetched / batched tickets, aggregated slot draws, batched mate scores:
# parity subset, correction stage per chunk for the draw sizes, stall profile of generation 4
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "correct_matrix or exact_integer or edge_shapes or huge_litter or very_wide or config4_deep or multilane" > $O/r2c47_pytest.log 2>&1
tail -4 $O/r2c47_pytest.log
for v in base draw1 draw2 draw4; do
  echo "== $v"
  GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_$v.so timeout 300 python scripts/bench_correct.py 10000 2>&1 | tail -1
done | tee $O/r2c47_variants.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_correct$|^k_correct_jobs" -s 66 -c 3 -o $O/r2c47_full_correct_gen4 -f \
    python scripts/bench_correct.py 10000 > $O/r2c47_ncu.log 2>&1
for k in 'k_correct$' 'k_correct_jobs$' 'k_correct_jobs_general'; do
  python scripts/ncu_stalls.py $O/r2c47_full_correct_gen4.ncu-rep "$k" 60 > "$O/r2c47_stalls_gen4_${k%\$}.txt" 2>&1
done
python scripts/ncu_summary.py $O/r2c47_full_correct_gen4.ncu-rep > $O/r2c47_summary_gen4.txt 2>&1
grep -E "Kernel Name|time_duration|inst_executed.sum|issue_active|warps_active" $O/r2c47_summary_gen4.txt
