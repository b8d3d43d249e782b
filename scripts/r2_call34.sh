#!/bin/bash
# Round 2, call 34: is it the shared-memory carve-out?  form 2 with its dynamic shared memory padded to form 3's size
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in "GFX_RANK_FORM=2" "GFX_RANK_FORM=2 GFX_RANK_SMEM_PAD=56000" "GFX_RANK_FORM=2 GFX_RANK_SMEM_PAD=20000"; do
  env $v python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c34_c4.json 2> $O/r2c34_c4.err
  python -c "
import json;d=json.load(open('$O/r2c34_c4.json'));r=d['roofline'];print('c4 [$v]',d['ms_per_step'],'rank call us',r['kernel'].get('us_per_call'),'kernel',r['kernel'].get('us_per_launch'),'correct share',r['correct_kernels'].get('share_of_step'))"
done
