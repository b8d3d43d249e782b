#!/bin/bash
# Round 2, call 7: parity at the BASELINE sizes, c1 launch list, c5 bench + ncu capture of k_trio_scan
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q -k "config1 or config4_named or config5 or multilane" > $O/r2c7_pytest.log 2>&1; tail -4 $O/r2c7_pytest.log
python bench.py --config c5 --steps 3 --warmup 3 > $O/r2c7_bench_c5.json 2> $O/r2c7_bench_c5.err; tail -c 900 $O/r2c7_bench_c5.json; tail -3 $O/r2c7_bench_c5.err
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -s 300 -c 320 --csv --log-file $O/r2c7_launches_c1.csv \
    python bench.py --config c1 --steps 1 --warmup 1 --no-cpu-baseline --lanes 1 > $O/r2c7_ncu_c1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trio_scan" -s 3 -c 1 -o $O/r2c7_full_trio -f \
    python bench.py --config c5 --animals 300000 --steps 1 --warmup 1 > $O/r2c7_ncu_trio.log 2>&1
ls -la $O/r2c7_full_trio.ncu-rep
