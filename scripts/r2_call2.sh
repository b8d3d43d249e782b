#!/bin/bash
# Round 2, call 2: parity of the integer fast path, correction-stage time, bench line
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c2_pytest.log 2>&1; tail -5 $O/r2c2_pytest.log
python scripts/bench_correct.py 10000 > $O/r2c2_correct.log 2>&1; tail -1 $O/r2c2_correct.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c2_bench.json 2> $O/r2c2_bench.err; tail -c 1500 $O/r2c2_bench.json
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file $O/r2c2_launches.csv \
    python scripts/bench_correct.py 4096 > $O/r2c2_ncu.log 2>&1
