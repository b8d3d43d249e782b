#!/bin/bash
# Round 2, call 58: 5 blocks per SM (96 registers) for the message-passing kernel in the 3-lane bench step; parity subset with it
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in kj5 main; do
  GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_$v.so timeout 300 python bench.py --steps 5 --warmup 3 > $O/r2c58_bench_$v.json 2> $O/r2c58_bench_$v.err
  python - <<PY
import json
d = json.loads(open("$O/r2c58_bench_$v.json").read().strip().splitlines()[-1])
print("$v", "ms_per_step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_kj5.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "correct_matrix or exact_integer or huge_litter or very_wide or config4_deep" > $O/r2c58_pytest_kj5.log 2>&1
tail -2 $O/r2c58_pytest_kj5.log
