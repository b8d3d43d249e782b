#!/bin/bash
# Round 2, call 8: tile-major trio scan, compact decision bytes + larger lists (c1), c3 on one GPU
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c8_pytest.log 2>&1; tail -4 $O/r2c8_pytest.log
python bench.py --config c5 --steps 3 --warmup 3 > $O/r2c8_bench_c5.json 2> $O/r2c8_bench_c5.err; python -c "
import json;d=json.load(open('$O/r2c8_bench_c5.json'));print('c5',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['chunked'])"; tail -3 $O/r2c8_bench_c5.err
python bench.py --config c1 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c8_bench_c1.json 2> $O/r2c8_bench_c1.err; python -c "
import json;d=json.load(open('$O/r2c8_bench_c1.json'));print('c1',d['value'],d['ms_per_step'],d['e2e']['value'])"; tail -3 $O/r2c8_bench_c1.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c8_bench_c2.json 2> $O/r2c8_bench_c2.err; python -c "
import json;d=json.load(open('$O/r2c8_bench_c2.json'));print('c2',d['value'],d['ms_per_step'],d['e2e']['value'])"; tail -3 $O/r2c8_bench_c2.err
python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline > $O/r2c8_bench_c3_n1.json 2> $O/r2c8_bench_c3_n1.err; python -c "
import json;d=json.load(open('$O/r2c8_bench_c3_n1.json'));print('c3',d['value'],d['ms_per_step'],d['e2e'],d['roofline']['frac'])"; tail -3 $O/r2c8_bench_c3_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trio_scan" -s 3 -c 1 -o $O/r2c8_full_trio -f \
    python bench.py --config c5 --animals 300000 --steps 1 --warmup 1 > $O/r2c8_ncu_trio.log 2>&1
