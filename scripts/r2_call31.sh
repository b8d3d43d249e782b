#!/bin/bash
# Round 2, call 31: leaner trio quad (fewer LOP3, merged popcounts, ragged tail split off): tests, config 5 timing, variants
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trio_scan" > $O/r2c31_pytest.log 2>&1
tail -3 $O/r2c31_pytest.log
timeout 600 python scripts/bench_trio_scan.py --steps 5 > $O/r2c31_trio.json 2> $O/r2c31_trio.err; tail -c 600 $O/r2c31_trio.json; echo
for f in genofix_b200/variants/lib_fb*.so; do
  echo "== $f"
  GFX_LIB_PATH=$PWD/$f timeout 600 python scripts/bench_trio_scan.py --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('ms_per_step'), d.get('roofline',{}).get('frac'))"
done
