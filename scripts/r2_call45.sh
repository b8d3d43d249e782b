#!/bin/bash
# Round 2, call 45: uploads of MultiLane's chunks in chunk order (engine.CopyOrder): lane / chunk tests, bench line of config 2
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane or chunks" > $O/r2c45_pytest.log 2>&1
tail -5 $O/r2c45_pytest.log
timeout 600 python bench.py > $O/r2c45_bench_c2.json 2> $O/r2c45_bench_c2.err
tail -3 $O/r2c45_bench_c2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c45_bench_c2.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"], "packed", d.get("e2e_packed_io", {}).get("ms_per_step"))
print("roofline frac", d["roofline"]["frac"], d["roofline"]["us_per_launch"])
PY
