#!/bin/bash
# Round 2, call 6: new tests (entry points, reference index), bench lines of configs c2 / c1 / c4 with the new bench.py
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c6_pytest.log 2>&1; tail -4 $O/r2c6_pytest.log
python bench.py --steps 3 --warmup 3 > $O/r2c6_bench_c2.json 2> $O/r2c6_bench_c2.err; tail -c 600 $O/r2c6_bench_c2.json; tail -3 $O/r2c6_bench_c2.err
python bench.py --config c1 --steps 5 --warmup 3 > $O/r2c6_bench_c1.json 2> $O/r2c6_bench_c1.err; tail -c 300 $O/r2c6_bench_c1.json; tail -3 $O/r2c6_bench_c1.err
python bench.py --config c4 --steps 2 --warmup 3 > $O/r2c6_bench_c4.json 2> $O/r2c6_bench_c4.err; tail -c 300 $O/r2c6_bench_c4.json; tail -3 $O/r2c6_bench_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c6_bench_ref.json 2> $O/r2c6_bench_ref.err; tail -c 400 $O/r2c6_bench_ref.json
