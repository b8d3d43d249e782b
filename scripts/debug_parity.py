import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
from oracle.genofix_oracle import Oracle
from test_gpu_parity import _random_case
from conftest import load_golden

def report(tag, eng, out, o, obs, ids):
    raw = eng.raw_host().view(np.uint16); ro = o["raw"].view(np.uint16)
    d = np.argwhere(raw != ro)
    print(tag, "raw diffs", len(d), "test_p", eng.stats["test_p"], o["test_p"], "corrected diffs", int((out != o["corrected"]).sum()),
          "tally", int(eng.stats["corrected_per_animal"].sum()), int(o["corrected_per_animal"].sum()),
          "excl", int(eng.stats["excluded_per_animal"].sum()), int(o["excluded_per_animal"].sum()))
    for r, j in d[:8]:
        print("   raw", int(ids[r]), j, "obs", obs[r, j], "gpu", eng.raw_host()[r, j], "oracle", o["raw"][r, j], "flags", eng.schedule.array(4)[r])
    if "E" in o:
        E = eng.E_host(); de = np.argwhere(~((E == o["E"]) | (np.isnan(E) & np.isnan(o["E"]))))
        print("   E diffs", len(de))
        for r, j in de[:5]:
            print("     E", r, j, repr(E[r, j]), repr(o["E"][r, j]), "raw", eng.raw_host()[r, j], o["raw"][r, j], 'obs', obs[r,j])
    dc = np.argwhere(out != o["corrected"])
    for r, j in dc[:8]:
        print("   corr", int(ids[r]), j, "obs", obs[r, j], "gpu", out[r, j], "oracle", o["corrected"][r, j])

for seed, params in [(101, (0.5, 0.64, 2, 0.95, 0.4, 1.0)), (303, (0.6, 0.6, 1, 0.9, 0.2, 1.0)), (404, (0.5, 0.64, 3, 0.92, 0.4, 0.9))]:
    rows, ids, obs = _random_case(seed, s=96 if params[2] == 3 else 160)
    tp, ts, back, q, tie, et = params
    o = Oracle(rows, ids).correct_matrix(obs, tp, ts, back=back, init_filter_p=q, tiethreshold=tie, err_thresh=et)
    eng = Engine(PedigreeDAG.from_table(rows), ids, back)
    out = eng.correct(obs, tp, ts, init_filter_p=q, tiethreshold=tie, err_thresh=et)
    print(eng.schedule.info)
    report("seed %d" % seed, eng, out, o, obs, ids)
g = load_golden("nan_loop")
tp, ts, back, q, tie, et = g["params"]
eng = Engine(PedigreeDAG.from_table(g["ped_rows"]), g["ids"], int(back))
out = eng.correct(g["obs"], tp, ts, init_filter_p=q, tiethreshold=tie, err_thresh=et)
o = dict(raw=g["raw"], test_p=float(g["test_p"]), corrected=g["corrected"], corrected_per_animal=np.array([int(g["errors_all"])]), excluded_per_animal=np.array([int(g["pair_excluded"])]))
report("nan_loop", eng, out, o, g["obs"], g["ids"])
print(eng.raw_host()); print(g["raw"])
