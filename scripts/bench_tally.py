"""Times the per-animal Mendel tally (error_checker.py:193-204) on one BASELINE config 2 chunk (10k animals x 10k SNPs):
fused rank + histogram, then k_rank_rowsum_f16 over the animals with both parents genotyped."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from genofix_b200 import synth, _lib
from genofix_b200.engine import Engine, half_transform_table
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
s = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
rows = synth.make_pedigree(n, 6, seed=1234)
obs = synth.inject_errors(synth.gene_drop(rows, s, seed=1234), 0.01, seed=1234)
eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
sel = np.nonzero((rows[:, 1] > 0) & (rows[:, 2] > 0))[0].astype(np.int32)
E, sums, m16 = eng.mendel_tally(obs, sel, -1.0, want_matrix=False)
dev = eng.raw.device
table, _ = half_transform_table(eng.host_hist.numpy().view(np.uint64), -1.0)
e16 = torch.from_numpy(table.view(np.int16)).to(dev)
rows_dev = torch.from_numpy(sel).to(dev)
out = torch.empty(len(sel), dtype=torch.int16, device=dev)
def call():
    _lib.check(eng.L.gfx_rank_rowsum_f16(eng.raw.data_ptr(), eng.ld_raw, eng.n, eng.s, e16.data_ptr(), rows_dev.data_ptr(),
                                         len(sel), out.data_ptr(), eng._stream()))
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
assert np.array_equal(out.cpu().numpy().view(np.uint16), sums.view(np.uint16))
print(json.dumps({"kernel": "k_rank_rowsum_f16", "animals": n, "snps": s, "rows_tallied": int(len(sel)), "us_per_call": ms * 1e3,
                  "bytes_read": int(len(sel)) * s * 2, "GBps": len(sel) * s * 2 / ms / 1e6,
                  "note": "serial float16 chain per animal (the reference's pandas order): latency bound by design"}))
