#!/bin/bash
# Round 2, call 56: bench lines of configs 1, 4 and 5 with the final build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for c in c4 c1 c5; do
  timeout 400 python bench.py --config $c > $O/r2c56_bench_$c.json 2> $O/r2c56_bench_$c.err
  python - <<PY
import json
d = json.loads(open("$O/r2c56_bench_$c.json").read().strip().splitlines()[-1])
print("$c", "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"].get("ms_per_step"), "frac", d["roofline"]["frac"])
PY
done
