#!/bin/bash
# Round 2, call 57 (2 GPUs): bench line of config 2 at N = 2 with the final build (weak scaling)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2c57_bench_c2_n2.json 2> $O/r2c57_bench_c2_n2.err
python -c "
import json;d=json.loads(open('$O/r2c57_bench_c2_n2.json').read().strip().splitlines()[-1]);print('c2 n2',d['value'],d['ms_per_step'],d['e2e']['value'],d['scaling'])"
