#!/bin/bash
# Round 2, call 55 (2 GPUs): the two-rank tests with the final build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_rank" > $O/r2c55_pytest_2gpu.log 2>&1
tail -3 $O/r2c55_pytest_2gpu.log
