#!/bin/bash
# Round 2, call 30: the empirical (LD-window) blend on the GPU against the reference fixtures and the oracle
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "empirical_blend or ld_index or entry_point_text" > $O/r2c30_pytest.log 2>&1
tail -30 $O/r2c30_pytest.log
