#!/bin/bash
# times every library variant built by build_variants.sh on the rank microbench
cd "$(dirname "$0")/.."
for f in genofix_b200/variants/lib_*.so; do
  echo "== $f"
  GFX_LIB_PATH=$PWD/$f python ${GFX_SCRIPT:-scripts/bench_rank.py} "$@" 2>&1 | tail -2
done
