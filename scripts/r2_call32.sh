#!/bin/bash
# Round 2, call 32: the whole GPU suite, smoke(), bench lines of every single-GPU config with the shipped build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2c32_pytest.log 2>&1
tail -4 $O/r2c32_pytest.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
for c in c2 c1 c4 c5; do
  python bench.py --config $c --steps 3 --warmup 3 > $O/r2c32_bench_$c.json 2> $O/r2c32_bench_$c.err
  python -c "
import json;d=json.load(open('$O/r2c32_bench_$c.json'));r=d.get('roofline',{});print('$c',d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),r.get('frac'), (d.get('cpu_baseline') or {}).get('value'))"
done
python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline > $O/r2c32_bench_c3_n1.json 2> $O/r2c32_bench_c3_n1.err
python -c "
import json;d=json.load(open('$O/r2c32_bench_c3_n1.json'));print('c3 n1',d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'))"
python bench.py --impl reference --steps 1 --warmup 0 > $O/r2c32_bench_reference_c2.json 2> $O/r2c32_bench_reference_c2.err; tail -c 400 $O/r2c32_bench_reference_c2.json
