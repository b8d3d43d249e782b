"""k_mendel_rank3 against k_mendel_rank (GFX_RANK_FORM=2): scores and histogram bit for bit on several data sets, then timings.

usage: bench_rank_forms.py [n_animals] [n_snps]   (timed shape, default BASELINE config 2 chunk 10000 x 10000)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genofix_b200 import synth
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG


def load(eng, obs):
    s = obs.shape[1]
    eng._alloc(s)
    hin = eng.host_in.numpy(); hin[:, :eng.s] = obs; hin[:, eng.s:] = 9
    eng.codes.copy_(eng.host_in); eng.load_codes_device(eng.codes)


def run(eng, form, fused):
    os.environ["GFX_RANK_FORM"] = str(form)
    eng.raw.fill_(-21846)
    eng.mendel_rank_device(fused=fused)
    torch.cuda.synchronize()
    return eng.raw[:, :eng.s].cpu().numpy().view(np.uint16).copy(), (eng.hist.cpu().numpy().copy() if fused else None)


def compare(name, rows, obs, back=2):
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], back)
    load(eng, obs)
    ok = True
    for fused in (True, False):
        r2, h2 = run(eng, 2, fused)
        r3, h3 = run(eng, 3, fused)
        same = np.array_equal(r2, r3) and (h2 is None or np.array_equal(h2, h3))
        ok &= same
        print("%-34s %-5s scores %s  hist %s  cells %d" % (name, "fused" if fused else "plain", np.array_equal(r2, r3),
              "-" if h2 is None else np.array_equal(h2, h3), r2.size), flush=True)
        if not same:
            bad = np.argwhere(r2 != r3)
            print("   first differences (row, snp):", bad[:5].tolist(), [(hex(r2[i, j]), hex(r3[i, j])) for i, j in bad[:5]])
            if h2 is not None:
                d = np.nonzero(h2 != h3)[0]
                print("   hist bins:", [(hex(int(b)), int(h2[b]), int(h3[b])) for b in d[:8]])
    eng.close() if hasattr(eng, "close") else None
    return ok


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    s = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    ok = True
    rng = np.random.default_rng(7)
    if os.environ.get("GFX_FORMS_TIME_ONLY"):
        return timing(n, s, (3,))
    # small and ragged widths, missing calls, unknown parents, deep loopy pedigrees, many children
    for (na, gens, ns, miss, seed) in [(300, 4, 1003, 0.0, 1), (300, 4, 1003, 0.05, 2), (1200, 6, 517, 0.01, 3), (2000, 6, 4096, 0.0, 4),
                                       (2000, 6, 40, 0.2, 5), (500, 3, 16, 0.0, 6), (500, 3, 17, 0.5, 7)]:
        rows = synth.make_pedigree(na, gens, seed=seed)
        obs = synth.inject_errors(synth.gene_drop(rows, ns, seed=seed), 0.01, seed=seed)
        if miss:
            obs = obs.copy(); obs[rng.random(obs.shape) < miss] = 9
        ok &= compare("synthetic %d x %d gens %d miss %.2f" % (na, ns, gens, miss), rows, obs)
    rows = synth.make_pedigree(1500, 15, sire_frac=0.02, dam_frac=0.3, seed=11, related_mating_p=0.25, unknown_parents_frac=0.1)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 640, seed=11), 0.01, seed=11), 0.05, seed=11)
    ok &= compare("deep loopy 1500 x 640, 5% missing", rows, obs)
    obs = synth.inject_errors(synth.gene_drop(rows, 2000, seed=12), 0.01, seed=12)
    ok &= compare("deep loopy 1500 x 2000, none missing", rows, obs)
    # few sires with very many children (m >= 32 and m >= 1024 paths)
    na = 3000
    rows = np.zeros((na, 4), dtype=np.int64); rows[:, 0] = np.arange(1, na + 1); rows[:, 3] = 1 + (np.arange(na) % 2)
    rows[2, 3] = rows[3, 3] = 2; rows[0, 3] = rows[1, 3] = 1
    rows[4:, 1] = 1 + (np.arange(na - 4) % 2); rows[4:, 2] = 3 + (np.arange(na - 4) % 2)
    obs = synth.inject_errors(synth.gene_drop(rows, 333, seed=5), 0.02, seed=5)
    ok &= compare("2 sires x 1498 kids, 3000 x 333", rows, obs)
    rows = synth.make_pedigree(n, 6, seed=1234)
    obs = synth.inject_errors(synth.gene_drop(rows, s, seed=1234), 0.01, seed=1234)
    ok &= compare("timed shape %d x %d" % (n, s), rows, obs)
    print("ALL EQUAL" if ok else "MISMATCH", flush=True)
    timing(n, s, (2, 3))
    return 0 if ok else 1


def timing(n, s, forms):
    rows = synth.make_pedigree(n, 6, seed=1234)
    obs = synth.inject_errors(synth.gene_drop(rows, s, seed=1234), 0.01, seed=1234)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    load(eng, obs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for form in forms:
        os.environ["GFX_RANK_FORM"] = str(form)
        for fused in (True, False):
            for _ in range(3):
                eng.mendel_rank_device(fused=fused)
            torch.cuda.synchronize()
            reps = 20
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            st = torch.cuda.current_stream()
            for e0, e1 in ev:
                if not os.environ.get("GFX_NOFLUSH"):
                    flush.fill_(1)
                e0.record(st)
                eng.mendel_rank_device(fused=fused)
                e1.record(st)
            torch.cuda.synchronize()
            t = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
            ms = float(np.median(t))
            print("form %d %-5s %.1f us per call (min %.1f), %.0f GB/s algorithmic (2.25 B/cell), frac of 6548.2 = %.3f" %
                  (form, "fused" if fused else "plain", ms * 1e3, t.min() * 1e3, n * s * 2.25 / ms / 1e6, n * s * 2.25 / ms / 1e6 / 6548.2), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
