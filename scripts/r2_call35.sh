#!/bin/bash
# Round 2, call 35: rank form chosen from the previous chunk's missing fraction: config 4 and config 2 step times, rank tests
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for c in c4 c2; do
  python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c35_$c.json 2> $O/r2c35_$c.err
  python -c "
import json;d=json.load(open('$O/r2c35_$c.json'));r=d['roofline'];print('$c',d['value'],d['ms_per_step'],d['e2e']['value'],'rank call us',r['kernel'].get('us_per_call'),'kernel',r['kernel'].get('us_per_launch'),r['kernel'].get('frac'))"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rank or golden or config4" > $O/r2c35_pytest.log 2>&1; tail -2 $O/r2c35_pytest.log
