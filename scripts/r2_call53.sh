#!/bin/bash
# Round 2, call 53: full GPU suite, smoke, bench line of config 2 (3 and 4 lanes), launch list of the bench command,
# blocks per SM of the deferred-job kernels after this session's changes
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c53_pytest.log 2>&1
tail -3 $O/r2c53_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > $O/r2c53_bench_c2.json 2> $O/r2c53_bench_c2.err
timeout 600 python bench.py --lanes 4 > $O/r2c53_bench_c2_lanes4.json 2> $O/r2c53_bench_c2_lanes4.err
python - <<'PY'
import json
for f in ("r2c53_bench_c2", "r2c53_bench_c2_lanes4"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "ms_per_step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "rank frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
PY
for v in main kj5 kj6 kg5 kg6; do
  echo "== $v"
  GFX_LIB_PATH=$PWD/genofix_b200/variants/lib_$v.so timeout 300 python scripts/bench_correct.py 10000 2>&1 | tail -1
done | tee $O/r2c53_variants.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 1200 --csv --log-file $O/r2c53_launches_c2.csv \
   python bench.py --steps 1 --warmup 1 > $O/r2c53_ncu.log 2>&1
python scripts/launch_agg.py $O/r2c53_launches_c2.csv > $O/r2c53_launches_c2.txt 2>&1
head -14 $O/r2c53_launches_c2.txt
