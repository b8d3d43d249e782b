"""Diagnostic counters of the correction kernels.  They are compiled out of the default build: rebuild with
`NVCC_FLAGS=-DGFX_DEBUG_COUNTERS` (genofix_b200/build.py) to get non-zero numbers."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from genofix_b200 import synth, _lib
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
rows = synth.make_pedigree(10000, 6, seed=1234)
obs = synth.inject_errors(synth.gene_drop(rows, 4096, seed=1234), 0.01, seed=1234)
eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
L = _lib.lib()
out = eng.correct(obs, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
print("test_p", eng.stats["test_p"], "info", eng.schedule.info)
cnt = np.zeros(16, dtype=np.uint64)
_lib.check(L.gfx_debug_counters(eng.schedule.handle, 1, cnt.ctypes.data_as(C.c_void_p)))
names = ["queries", "parent-only", "eliminations", "big reruns", "nodes touched", "suspect cells", "job evals (fast + deferred)", "children visited", "fast: not plainly peelable (+ xpeel ok in deferred)", "fast: guard band (+ x:visited-twice)", "fast: children term known", "x:q-has-kids", "x:ok tree-only", "x:ok obs-kids/obs-mates", "x:ok recursion", "x:failed (3 x 20 bits)"]
# per phase
eng._alloc(obs.shape[1])
eng.mendel_rank_device(); eng.thresholds_device(0.95)
eng.packed_live.copy_(eng.packed_orig); eng.tallies.zero_()
pp = _lib.CorrectParams(eng.test_p, 0.2, 0.4, 0.5, 1.0, eng.cutraw.data_ptr(), eng.tp_raw, 0)
for g in range(eng.schedule.info["n_generations"]):
    torch.cuda.synchronize(); t = time.time()
    _lib.check(L.gfx_correct_pairs(eng.schedule.handle, g, eng.packed_live.data_ptr(), eng.s, eng.wpr,
               eng.raw.data_ptr(), eng.ld_raw, eng.elut.data_ptr(), C.byref(pp), eng.snapshot.data_ptr(), eng.snapshot.numel(), None, eng.tallies.data_ptr(), None))
    torch.cuda.synchronize(); dt = time.time() - t
    _lib.check(L.gfx_debug_counters(eng.schedule.handle, 1, cnt.ctypes.data_as(C.c_void_p)))
    print("gen", g, "%.1f ms (with counters)" % (dt * 1e3), dict(zip(names, [int(x) for x in cnt[:16]])))
ps = _lib.CorrectParams(eng.test_p, 0.2, 0.4, 0.64, 1.0, eng.cutraw.data_ptr(), eng.tp_raw, 0)
torch.cuda.synchronize(); t = time.time()
_lib.check(L.gfx_correct_singles(eng.schedule.handle, eng.packed_live.data_ptr(), eng.s, eng.wpr,
           eng.raw.data_ptr(), eng.ld_raw, eng.elut.data_ptr(), C.byref(ps), eng.snapshot.data_ptr(), eng.snapshot.numel(), None, eng.tallies.data_ptr(), None))
torch.cuda.synchronize(); dt = time.time() - t
_lib.check(L.gfx_debug_counters(eng.schedule.handle, 0, cnt.ctypes.data_as(C.c_void_p)))
print("singles %.1f ms" % (dt * 1e3), dict(zip(names, [int(x) for x in cnt[:16]])))
raw = eng.raw_host().astype(np.float32)
print("raw nonzero frac", float((raw != 0).mean()), "E>=test_p frac", float((eng.E_host() >= eng.test_p).mean()))
