#!/bin/bash
# Round-1 evidence on one B200: tests, bench line, ncu launch list of the bench command, full captures of the top kernels,
# config-3 and config-5 shapes.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r1_pytest_gpu.log 2>&1; tail -2 $O/r1_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $O/r1_bench.json 2> $O/r1_bench.err && tail -c 600 $O/r1_bench.json
python bench.py --impl reference --steps 1 --warmup 0 > $O/r1_bench_reference.json 2>> $O/r1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r1_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r1_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_mendel_rank" -s 6 -c 2 -o $O/r1_full_rank -f \
    python scripts/bench_rank.py > $O/r1_ncu_rank.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_correct$|k_correct_jobs" -s 12 -c 3 -o $O/r1_full_correct -f \
    python scripts/bench_correct.py 4096 > $O/r1_ncu_correct.log 2>&1
python scripts/bench_rank.py 100000 10000 10 > $O/r1_rank_c3.log 2>&1; tail -2 $O/r1_rank_c3.log
python bench.py --animals 100000 --generations 10 --snps 20000 --chunk 10000 --steps 2 --warmup 3 --no-cpu-baseline > $O/r1_bench_c3shape.json 2> $O/r1_bench_c3shape.err; tail -c 400 $O/r1_bench_c3shape.json
python scripts/bench_trio_scan.py > $O/r1_trio_scan_c5.log 2>&1; tail -2 $O/r1_trio_scan_c5.log
