"""BASELINE config 5: empirical Mendelian-consistency scan, 1M animals x 50k SNPs, 2-bit packed (12.5 GB in HBM).

Random packed genotypes are generated on the device (the kernel's cost does not depend on the values); a
sample of trios is checked against a torch restatement (torch here is the checker, not the product).
Prints one JSON line with cells/s and the HBM roofline fraction of k_trio_scan (0.25 B/cell algorithmic)."""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from genofix_b200 import _lib, synth
from genofix_b200.engine import Schedule
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG

ap = argparse.ArgumentParser()
ap.add_argument("--animals", type=int, default=1000000)
ap.add_argument("--snps", type=int, default=50000)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
L = _lib.lib()
t0 = time.time()
rows = synth.make_pedigree(a.animals, 10, seed=1234)
ped = PedigreeDAG.from_table(rows)
sched = Schedule(ped, rows[:, 0], 2)
t_sched = time.time() - t0
n, s = a.animals, a.snps
nwords = (s + 15) // 16
wpr = ((nwords + 3) // 4) * 4
packed = torch.empty((n, wpr), dtype=torch.int32, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(1)
for i in range(0, n, 65536):   # 2 % missing (code 3), the rest uniform 0/1/2, built 16 codes per word
    m = min(65536, n - i)
    w = torch.zeros((m, wpr), dtype=torch.int64, device="cuda")
    for k in range(16):
        c = torch.randint(0, 3, (m, wpr), device="cuda", generator=g)
        miss = torch.rand((m, wpr), device="cuda", generator=g) < 0.02
        c = torch.where(miss, torch.full_like(c, 3), c)
        w |= c << (2 * k)
    packed[i:i + m] = (w & 0xFFFFFFFF).to(torch.int64).where(w < 2 ** 31, w - 2 ** 32).to(torch.int32)
inc = torch.empty(n, dtype=torch.int64, device="cuda")
tst = torch.empty(n, dtype=torch.int64, device="cuda")
def run():
    _lib.check(L.gfx_trio_scan(sched.handle, packed.data_ptr(), s, wpr, inc.data_ptr(), tst.data_ptr(), None))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
# check a sample of trios with torch
k = sched._keep
has = np.nonzero((k["sire"] >= 0) & (k["dam"] >= 0))[0]
sel = has[np.random.default_rng(0).choice(len(has), 64, replace=False)]
def unpack(r):
    w = packed[int(r)].to(torch.int64) & 0xFFFFFFFF
    sh = torch.arange(16, device="cuda") * 2
    return ((w[:, None] >> sh[None, :]) & 3).reshape(-1)[:s]
T = torch.tensor([[1, .5, 0, .5, .25, 0, 0, 0, 0], [0, .5, 1, .5, .5, .5, 1, .5, 0], [0, 0, 0, 0, .25, .5, 0, .5, 1]], device="cuda").reshape(3, 3, 3)
ok = True
for a_ in sel:
    kk, ss, dd = unpack(k["row"][a_]), unpack(k["row"][k["sire"][a_]]), unpack(k["row"][k["dam"][a_]])
    pres = (kk < 3) & (ss < 3) & (dd < 3)
    bad = pres & (T[kk.clamp(max=2), ss.clamp(max=2), dd.clamp(max=2)] == 0)
    ok &= int(bad.sum()) == int(inc[k["row"][a_]]) and int(pres.sum()) == int(tst[k["row"][a_]])
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("hbm_gbs", 6650.0))
trios = sched.info["n_trios"]
cells = n * s
ach = cells * 0.25 / (ms / 1e3) / 1e9
print(json.dumps(dict(workload="config 5: trio consistency scan %d animals x %d SNPs, 2-bit packed (%.1f GB)" % (n, s, n * wpr * 4 / 1e9),
                      ms_per_scan=ms, cells_per_s=cells / (ms / 1e3), trios=trios, sample_check_ok=bool(ok),
                      roofline=dict(bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                                    algorithmic_bytes_per_cell=0.25,
                                    note="one warp per (sire, dam) family: kid rows are read once, the parents once per group of 4 kids"),
                      schedule_build_s=t_sched, inconsistent_total=int(inc.sum()), tested_total=int(tst.sum()))))
