"""Experiment: correction stage of L independent chunks run concurrently on L streams (one host thread each)."""
import sys, os, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genofix_b200 import synth
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
s = 10000
rows = synth.make_pedigree(10000, 6, seed=1234)
ped = PedigreeDAG.from_table(rows)
for lanes in (1, 2, 3):
    engs, streams = [], []
    for l in range(lanes):
        obs = synth.inject_errors(synth.gene_drop(rows, s, seed=1234 + l), 0.01, seed=1234 + l)
        eng = Engine(ped, rows[:, 0], 2)
        eng._alloc(s)
        hin = eng.host_in.numpy(); hin[:, :eng.s] = obs; hin[:, eng.s:] = 9
        eng.codes.copy_(eng.host_in); eng.load_codes_device(eng.codes)
        engs.append(eng); streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()

    def work(l, reps):
        with torch.cuda.stream(streams[l]):
            for _ in range(reps):
                engs[l].mendel_rank_device(); engs[l].thresholds_device(0.95)
                engs[l].correct_device(0.5, 0.64, 0.4, 1.0)
            streams[l].synchronize()
    for reps in (2, 6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(l, reps)) for l in range(lanes)]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("lanes %d: %.2f ms per chunk (whole chunk: rank + thresholds + correction)" % (lanes, dt * 1e3 / (reps * lanes)))
    del engs
