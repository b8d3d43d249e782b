#!/bin/bash
# Round 2, call 11 (N GPUs): c3 strong scaling (one 100k x 600k matrix, 60 chunks dealt round-robin), c2 weak scaling
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
mkdir -p $O
nvidia-smi -L > $O/r2c11_gpus_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --config c3 --gpus $N --steps 2 --warmup 1 --no-cpu-baseline \
   > $O/r2c11_bench_c3_n$N.json 2> $O/r2c11_bench_c3_n$N.err; python -c "
import json;d=json.load(open('$O/r2c11_bench_c3_n$N.json'));print('c3 n$N',d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_imbalance'),d['roofline']['frac'])"; tail -2 $O/r2c11_bench_c3_n$N.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 3 --warmup 3 \
   > $O/r2c11_bench_c2_n$N.json 2> $O/r2c11_bench_c2_n$N.err; python -c "
import json;d=json.load(open('$O/r2c11_bench_c2_n$N.json'));print('c2 n$N',d['value'],d['ms_per_step'],d['e2e']['value'],d.get('rank_imbalance'))"; tail -2 $O/r2c11_bench_c2_n$N.err | cut -c1-300
