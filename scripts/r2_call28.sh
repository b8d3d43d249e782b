#!/bin/bash
# Round 2, call 28: bench line of config 2 with the cached work lists (cudaMallocAsync in the lanes doubled the step)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python bench.py --steps 3 --warmup 3 > $O/r2c28_bench_c2.json 2> $O/r2c28_bench_c2.err; python -c "
import json;d=json.load(open('$O/r2c28_bench_c2.json'));print('c2',d['value'],d['ms_per_step'],d['e2e']['value']);k=d['roofline']['kernel'];print({x:k[x] for x in ('achieved','frac','us_per_launch','us_per_call','achieved_in_step','share_of_step')})"
tail -3 $O/r2c28_bench_c2.err
GFX_RANK_FORM=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c28_bench_c2_form2.json 2> $O/r2c28_bench_c2_form2.err; python -c "
import json;d=json.load(open('$O/r2c28_bench_c2_form2.json'));print('c2 form2',d['value'],d['ms_per_step'],d['e2e']['value'])"
