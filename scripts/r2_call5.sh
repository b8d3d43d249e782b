#!/bin/bash
# Round 2, call 5: one-warp blocks in the deferred-job kernels; lanes experiment
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c5_pytest.log 2>&1; tail -3 $O/r2c5_pytest.log
python scripts/bench_correct.py 10000 > $O/r2c5_correct.log 2>&1; tail -1 $O/r2c5_correct.log
for L in 3 5; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lanes $L > $O/r2c5_bench_l$L.json 2> $O/r2c5_bench_l$L.err; python -c "
import json;d=json.load(open('$O/r2c5_bench_l$L.json'));print('lanes $L',d['value'],d['ms_per_step'],d['e2e']['value'])"
done
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -s 46 -c 46 --csv --log-file $O/r2c5_launches.csv \
    python scripts/bench_correct.py 4096 > $O/r2c5_ncu.log 2>&1
