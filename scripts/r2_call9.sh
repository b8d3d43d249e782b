#!/bin/bash
# Round 2, call 9 (2 GPUs): the 2-rank product tests, c2 weak scaling and c3 strong scaling at N = 2
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2c9_gpus.txt
python -m pytest tests -m gpu -x -q -k "two_rank or trio_scan" > $O/r2c9_pytest.log 2>&1; tail -4 $O/r2c9_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 3 --warmup 3 \
   > $O/r2c9_bench_c2_n2.json 2> $O/r2c9_bench_c2_n2.err; python -c "
import json;d=json.load(open('$O/r2c9_bench_c2_n2.json'));print('c2 n2',d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_imbalance'))"; tail -2 $O/r2c9_bench_c2_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --config c3 --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline \
   > $O/r2c9_bench_c3_n2.json 2> $O/r2c9_bench_c3_n2.err; python -c "
import json;d=json.load(open('$O/r2c9_bench_c3_n2.json'));print('c3 n2',d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_imbalance'))"; tail -2 $O/r2c9_bench_c3_n2.err
