#!/bin/bash
# Round 2, call 10: TMA-staged trio scan (GFX_TRIO_TMA=1) -- parity tests and the c5 bench line against the LDG kernel
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
GFX_TRIO_TMA=1 timeout 600 python -m pytest tests -m gpu -x -q -k "trio_scan or config5 or config3_scale" > $O/r2c10_pytest_tma.log 2>&1; tail -3 $O/r2c10_pytest_tma.log
GFX_TRIO_TMA=1 timeout 600 python bench.py --config c5 --steps 3 --warmup 3 > $O/r2c10_bench_c5_tma.json 2> $O/r2c10_bench_c5_tma.err; python -c "
import json;d=json.load(open('$O/r2c10_bench_c5_tma.json'));print('c5 tma',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['chunked']['frac'])"; tail -3 $O/r2c10_bench_c5_tma.err
GFX_TRIO_TMA=0 timeout 600 python bench.py --config c5 --steps 3 --warmup 3 > $O/r2c10_bench_c5_ldg.json 2> $O/r2c10_bench_c5_ldg.err; python -c "
import json;d=json.load(open('$O/r2c10_bench_c5_ldg.json'));print('c5 ldg',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['chunked']['frac'])"; tail -3 $O/r2c10_bench_c5_ldg.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c10_bench_c2.json 2> $O/r2c10_bench_c2.err; python -c "
import json;d=json.load(open('$O/r2c10_bench_c2.json'));print('c2',d['value'],d['ms_per_step'],d['e2e']['value'])"
GFX_TRIO_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trio_scan" -s 3 -c 1 -o $O/r2c10_full_trio_tma -f \
    python bench.py --config c5 --animals 300000 --steps 1 --warmup 1 > $O/r2c10_ncu_trio.log 2>&1
