#!/bin/bash
# Round 2, call 40 (2 GPUs): the two-rank tests and a 2-GPU bench line of configs 2 and 3 with the shipped build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L | tee $O/r2c40_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q -k "two_rank or sharded" > $O/r2c40_pytest.log 2>&1; tail -3 $O/r2c40_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2c40_bench_c2_n2.json 2> $O/r2c40_bench_c2_n2.err
python -c "
import json;d=json.load(open('$O/r2c40_bench_c2_n2.json'));print('c2 n2',d['value'],d['ms_per_step'],d['e2e']['value'],d['scaling'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --config c3 --gpus 2 --steps 1 --warmup 1 > $O/r2c40_bench_c3_n2.json 2> $O/r2c40_bench_c3_n2.err
python -c "
import json;d=json.load(open('$O/r2c40_bench_c3_n2.json'));print('c3 n2',d['value'],d['ms_per_step'],d['e2e']['value'],d['scaling'])"
