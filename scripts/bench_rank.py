"""Times the fused rank + histogram call on one BASELINE config 2 chunk (10k animals x 10k SNPs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genofix_b200 import synth
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
s = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
gens = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rows = synth.make_pedigree(n, gens, seed=1234)
obs = synth.inject_errors(synth.gene_drop(rows, s, seed=1234), 0.01, seed=1234)
eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
eng._alloc(s)
hin = eng.host_in.numpy(); hin[:, :eng.s] = obs; hin[:, eng.s:] = 9
eng.codes.copy_(eng.host_in); eng.load_codes_device(eng.codes)
for fused in (True, False):
    for _ in range(3):
        eng.mendel_rank_device(fused=fused)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        eng.mendel_rank_device(fused=fused)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("fused" if fused else "plain", "%.1f us per call, %.0f GB/s algorithmic (2.25 B/cell)" % (ms * 1e3, n * s * 2.25 / ms / 1e6))
