#!/bin/bash
# Round 2, call 27: bench line of config 2 with k_mendel_rank3 (kernel timed by the library's own events), full ncu capture of
# the shipped build, launch list of one bench step
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python bench.py --steps 3 --warmup 3 > $O/r2c27_bench_c2.json 2> $O/r2c27_bench_c2.err; python -c "
import json;d=json.load(open('$O/r2c27_bench_c2.json'));print('c2',d['value'],d['ms_per_step'],d['e2e']['value']);k=d['roofline']['kernel'];print({x:k[x] for x in ('achieved','frac','us_per_launch','us_per_call','achieved_in_step','share_of_step')})"
tail -3 $O/r2c27_bench_c2.err
GFX_FORMS_TIME_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_mendel_rank3" -s 4 -c 1 -o $O/r2c27_full_rank3 -f \
  python scripts/bench_rank_forms.py > $O/r2c27_ncu.log 2>&1
tail -2 $O/r2c27_ncu.log
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file $O/r2c27_launches_c2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r2c27_ncu_bench.log 2>&1
tail -1 $O/r2c27_ncu_bench.log | head -c 300
