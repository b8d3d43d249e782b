#!/bin/bash
# Round 2, call 1: parity of the build with the component-sorted deferred jobs, and its effect on the correction stage.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c1_pytest.log 2>&1; tail -3 $O/r2c1_pytest.log
GFX_COMP_SORT=0 python scripts/bench_correct.py 10000 > $O/r2c1_correct_sort0.log 2>&1; tail -1 $O/r2c1_correct_sort0.log
GFX_COMP_SORT=1 python scripts/bench_correct.py 10000 > $O/r2c1_correct_sort1.log 2>&1; tail -1 $O/r2c1_correct_sort1.log
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file $O/r2c1_launches_sort1.csv \
    python scripts/bench_correct.py 4096 > $O/r2c1_ncu1.log 2>&1
GFX_COMP_SORT=0 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file $O/r2c1_launches_sort0.csv \
    python scripts/bench_correct.py 4096 > $O/r2c1_ncu0.log 2>&1
