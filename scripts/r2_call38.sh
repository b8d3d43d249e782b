#!/bin/bash
# Round 2, call 38: compute-sanitizer (memcheck, then racecheck on the rank kernels) over smoke() and a small rank-forms run
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
cat > /tmp/small_forms.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from genofix_b200 import synth
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
for (na, gens, ns, miss, seed, kw) in [(400, 5, 1003, 0.0, 1, {}), (400, 5, 517, 0.03, 2, {}), (600, 12, 330, 0.05, 3, dict(related_mating_p=0.3, unknown_parents_frac=0.1, sire_frac=0.05, dam_frac=0.3))]:
    rows = synth.make_pedigree(na, gens, seed=seed, **kw)
    obs = synth.inject_errors(synth.gene_drop(rows, ns, seed=seed), 0.01, seed=seed)
    if miss:
        obs = synth.inject_missing(obs, miss, seed=seed)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    out = eng.correct(obs, 0.5, 0.64, init_filter_p=0.9, tiethreshold=0.1)
    tr = eng.trio_scan(obs)
    print(na, ns, miss, int((out != obs).sum()), int(tr[0].sum()))
print("small forms done")
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/small_forms.py > $O/r2c38_memcheck.log 2>&1; echo "memcheck rc $?" | tee -a $O/r2c38_memcheck.log
grep -E "ERROR SUMMARY|small forms done|Invalid|out of bounds" $O/r2c38_memcheck.log | head -8
timeout 1200 compute-sanitizer --tool racecheck --kernel-regex kns=k_mendel_rank --error-exitcode 7 python /tmp/small_forms.py > $O/r2c38_racecheck.log 2>&1; echo "racecheck rc $?" | tee -a $O/r2c38_racecheck.log
grep -E "RACECHECK SUMMARY|small forms done|hazard" $O/r2c38_racecheck.log | head -8
