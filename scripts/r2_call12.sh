#!/bin/bash
# Round 2, call 12: shared-memory first tier of the general elimination; fresh per-line profile of the rank kernel
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c12_pytest.log 2>&1; tail -4 $O/r2c12_pytest.log
python scripts/bench_correct.py 10000 > $O/r2c12_correct.log 2>&1; tail -1 $O/r2c12_correct.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c12_bench_c2.json 2> $O/r2c12_bench_c2.err; python -c "
import json;d=json.load(open('$O/r2c12_bench_c2.json'));print('c2',d['value'],d['ms_per_step'],d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -s 46 -c 46 --csv --log-file $O/r2c12_launches.csv \
    python scripts/bench_correct.py 4096 > $O/r2c12_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_mendel_rank" -s 6 -c 2 -o $O/r2c12_full_rank -f \
    python scripts/bench_rank.py > $O/r2c12_ncu_rank.log 2>&1
