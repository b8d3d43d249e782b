#!/bin/bash
# Round 2, call 14: correction-stage variants (occupancy of the deferred-job kernels, ticket sizes), 10k x 10k chunk
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
GFX_SCRIPT=scripts/bench_correct.py bash scripts/run_variants.sh 10000 > $O/r2c14_variants.log 2>&1
cat $O/r2c14_variants.log
