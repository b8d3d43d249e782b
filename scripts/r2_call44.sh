#!/bin/bash
# Round 2, call 44: linkage-aware gene drop (quickgsim's meiosis) on the device: property test, simulate.py --device_sim
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "linked_gene_drop or simulate_and_founder or device_simulator" > $O/r2c44_pytest.log 2>&1
tail -25 $O/r2c44_pytest.log
