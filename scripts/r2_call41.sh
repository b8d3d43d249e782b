#!/bin/bash
# Round 2, call 41 (8 GPUs): bench lines of config 3 (strong scaling) and config 2 (weak scaling) at N = 4 and 8, shipped build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --config c3 --gpus $n --steps 2 --warmup 1 > $O/r2c41_bench_c3_n$n.json 2> $O/r2c41_bench_c3_n$n.err
  python -c "
import json;d=json.load(open('$O/r2c41_bench_c3_n$n.json'));print('c3 n$n',d['value'],d['ms_per_step'],d['e2e']['value'],d['scaling'])"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 > $O/r2c41_bench_c2_n$n.json 2> $O/r2c41_bench_c2_n$n.err
  python -c "
import json;d=json.load(open('$O/r2c41_bench_c2_n$n.json'));print('c2 n$n',d['value'],d['ms_per_step'],d['e2e']['value'],d['scaling'])"
done
