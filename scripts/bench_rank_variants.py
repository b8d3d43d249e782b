"""Times the fused rank + histogram call of every library variant in genofix_b200/variants/ on one BASELINE config 2
chunk (10k animals x 10k SNPs) in ONE process, and checks each variant's scores and histogram bit for bit against
the first library (lib_base.so) on that chunk and on a hard matrix (loopy pedigree, missing calls, ragged width)."""
import glob, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genofix_b200 import synth, _lib
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sorted(glob.glob(os.path.join(root, "genofix_b200", "variants", "lib_*.so")), key=lambda p: (not p.endswith("lib_base.so"), p))
rows = synth.make_pedigree(10000, 6, seed=1234)
obs = synth.inject_errors(synth.gene_drop(rows, 10000, seed=1234), 0.01, seed=1234)
rows2 = synth.make_pedigree(1500, 9, sire_frac=0.05, dam_frac=0.4, seed=5, related_mating_p=0.4, unknown_parents_frac=0.1)
obs2 = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows2, 1237, seed=6), 0.03, seed=7), 0.05, seed=8)
rows3 = synth.make_pedigree(100000, 10, seed=1234)            # BASELINE config 3 pedigree, 2 048 SNPs of it
obs3 = synth.inject_errors(synth.gene_drop(rows3, 2048, seed=1234), 0.01, seed=1234)
ped, ped2, ped3 = PedigreeDAG.from_table(rows), PedigreeDAG.from_table(rows2), PedigreeDAG.from_table(rows3)
ref = {}
for path in libs:
    name = os.path.basename(path)[4:-3]
    _lib._lib = None
    _lib.LIB_PATH = path
    out = {"variant": name}
    try:
        for key, (p, r, o) in {"c2": (ped, rows, obs), "hard": (ped2, rows2, obs2), "c3": (ped3, rows3, obs3)}.items():
            eng = Engine(p, r[:, 0], 2)
            eng._alloc(o.shape[1])
            hin = eng.host_in.numpy(); hin[:, :eng.s] = o; hin[:, eng.s:] = 9
            eng.codes.copy_(eng.host_in); eng.load_codes_device(eng.codes)
            eng.mendel_rank_device(fused=True)
            torch.cuda.synchronize()
            raw, hist = eng.raw[:, :eng.s].clone(), eng.hist.clone()
            if name == "base":
                ref[key] = (raw, hist)
            out["same_" + key] = bool(torch.equal(raw, ref[key][0]) and torch.equal(hist, ref[key][1]))
            if key in ("c2", "c3"):
                for fused in (True, False):
                    for _ in range(3):
                        eng.mendel_rank_device(fused=fused)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        eng.mendel_rank_device(fused=fused)
                    e1.record(); torch.cuda.synchronize()
                    out[("us_fused_" if fused else "us_plain_") + key] = e0.elapsed_time(e1) / 20 * 1e3
            del eng
    except Exception as e:  # a variant that fails to launch must not hide the others
        out["error"] = repr(e)[:200]
    print(json.dumps(out), flush=True)
