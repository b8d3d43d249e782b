"""Times the correction stage (all generations + singles) of one 10k x N-SNP chunk, resident data."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genofix_b200 import synth
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
s = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rows = synth.make_pedigree(10000, 6, seed=1234)
obs = synth.inject_errors(synth.gene_drop(rows, s, seed=1234), 0.01, seed=1234)
eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
eng._alloc(s)
hin = eng.host_in.numpy(); hin[:, :eng.s] = obs; hin[:, eng.s:] = 9
eng.codes.copy_(eng.host_in); eng.load_codes_device(eng.codes)
eng.mendel_rank_device(); eng.thresholds_device(0.95)
for _ in range(2):
    eng.correct_device(0.5, 0.64, 0.4, 1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.correct_device(0.5, 0.64, 0.4, 1.0)
e1.record(); torch.cuda.synchronize()
print("correction stage %.2f ms per chunk of %d SNPs; corrected %d" % (e0.elapsed_time(e1) / 5, s, int(eng.tallies[:eng.n].sum())))
