#!/bin/bash
# Round 2, call 13: statuses of a blanket prefetched before the walks of the deferred-job kernels
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c13_pytest.log 2>&1; tail -4 $O/r2c13_pytest.log
python scripts/bench_correct.py 10000 > $O/r2c13_correct.log 2>&1; tail -1 $O/r2c13_correct.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c13_bench_c2.json 2> $O/r2c13_bench_c2.err; python -c "
import json;d=json.load(open('$O/r2c13_bench_c2.json'));print('c2',d['value'],d['ms_per_step'],d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -s 46 -c 46 --csv --log-file $O/r2c13_launches.csv \
    python scripts/bench_correct.py 4096 > $O/r2c13_ncu.log 2>&1
