#!/bin/bash
# Round 2, call 39: how many lanes?  bench step of config 2 with 1, 2, 3, 4, 5 lanes (5 chunks per step)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for l in 1 2 3 5; do
  python bench.py --lanes $l --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c39_l$l.json 2> $O/r2c39_l$l.err
  python -c "
import json;d=json.load(open('$O/r2c39_l$l.json'));print('lanes $l',d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'])"
done
