#!/bin/bash
# Round 2, call 51: blocks per SM of the fast correction kernel, in the 3-lane bench step of config 2 and alone
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in main kc7 kc8; do
  echo "== $v"
  GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_$v.so timeout 300 python scripts/bench_correct.py 10000 2>&1 | tail -1
  GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_$v.so timeout 600 python bench.py > $O/r2c51_bench_$v.json 2> $O/r2c51_bench_$v.err
  python - <<PY
import json
d = json.loads(open("$O/r2c51_bench_$v.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "rank frac", d["roofline"]["frac"])
PY
done | tee $O/r2c51_variants.log
