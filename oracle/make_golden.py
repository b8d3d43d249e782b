"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/src) here.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):
    python oracle/make_golden.py [name ...]
The fixtures are committed; the GPU box only reads them.
"""
import contextlib
import io
import os
import re
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genofix_b200 import synth  # noqa: E402  (input generator only)
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (pedigree spec, n_snps, error rate, missing rate, drop-genotype fraction, correctMatrix params)
GENOFIX_DEFAULTS = dict(threshold_pair=0.5, threshold_single=0.64, back=2, init_filter_p=0.95,
                        filter_e=0.99, tiethreshold=0.4, err_thresh=1)       # genofix.py:95-112
SIMULATE_DEFAULTS = dict(threshold_pair=0.5, threshold_single=0.7, back=2, init_filter_p=0.9,
                         filter_e=0.8, tiethreshold=0.05, err_thresh=1)      # simulate.py:89-113
CASES = {
    "toy_a": dict(ped=dict(n_animals=60, n_generations=5, sire_frac=0.3, dam_frac=0.6, seed=7),
                  n_snps=30, err=0.03, miss=0.03, drop=0.0, params=GENOFIX_DEFAULTS),
    "toy_b": dict(ped=dict(n_animals=96, n_generations=6, sire_frac=0.25, dam_frac=0.5, seed=11,
                           related_mating_p=0.6, unknown_parents_frac=0.1),
                  n_snps=32, err=0.02, miss=0.05, drop=0.12, params=SIMULATE_DEFAULTS),
    "deep": dict(ped=dict(n_animals=180, n_generations=15, sire_frac=0.34, dam_frac=0.5, seed=23,
                          related_mating_p=0.5, unknown_parents_frac=0.1),
                 n_snps=24, err=0.02, miss=0.03, drop=0.0,
                 params=dict(GENOFIX_DEFAULTS, init_filter_p=0.9)),
    # D16: a loopy blanket whose shared ancestor is unobserved + contradictory evidence -> NaN rank
    "nan_loop": dict(ped="crafted_nan", n_snps=6, err=0.0, miss=0.0, drop=0.0, params=GENOFIX_DEFAULTS),
    "example_fam": dict(ped="example.fam", n_snps=12, err=0.01, miss=0.0, drop=0.0,
                        params=GENOFIX_DEFAULTS),
}


def build_inputs(case):
    if case["ped"] == "example.fam":
        rows = np.loadtxt("/root/reference/examples/simulate/example.fam", dtype=np.int64, usecols=(0, 1, 2, 3))
        # founders that only appear as parents need rows for the gene drop; order parents first
        known = set(rows[:, 0])
        extra = []
        for col, sex in ((1, 1), (2, 2)):
            for p in np.unique(rows[:, col]):
                if p > 0 and p not in known:
                    extra.append((p, 0, 0, sex))
                    known.add(p)
        drop_rows = np.array(extra + [tuple(r) for r in rows], dtype=np.int64)
        drop_rows = topo_sort(drop_rows)
        ped_rows = rows
    elif case["ped"] == "crafted_nan":
        # A(1) x X(2) -> S(4);  A(1) x Y(3) -> D(5);  S x D -> 6,7,8 ; extra founders 9,10 -> 11
        drop_rows = np.array([[1, 0, 0, 1], [2, 0, 0, 2], [3, 0, 0, 2], [9, 0, 0, 1], [10, 0, 0, 2],
                              [4, 1, 2, 1], [5, 1, 3, 2], [6, 4, 5, 1], [7, 4, 5, 2], [8, 4, 5, 1],
                              [11, 9, 10, 2]], dtype=np.int64)
        ped_rows = drop_rows
    else:
        drop_rows = synth.make_pedigree(**case["ped"])
        ped_rows = drop_rows
    truth = synth.gene_drop(drop_rows, case["n_snps"], seed=1234)
    obs = synth.inject_errors(truth, case["err"], seed=1234)
    if case["miss"] > 0:
        obs = synth.inject_missing(obs, case["miss"], seed=4321)
    ids = drop_rows[:, 0].copy()
    if case["ped"] == "crafted_nan":
        r = {int(k): i for i, k in enumerate(ids)}
        obs[r[4], 0] = 9   # S missing
        obs[r[1], 0] = 9   # shared grandsire A missing
        obs[r[5], 0] = 2   # D = 2 ...
        obs[r[3], 0] = 0   # ... but her dam Y = 0  -> P(evidence) = 0 inside K's component
        obs[r[7], 3] = 9
    if case["drop"] > 0:
        rs = np.random.Generator(np.random.PCG64(99))
        keep = rs.random(len(ids)) >= case["drop"]
        ids, truth, obs = ids[keep], truth[keep], obs[keep]
    return ped_rows, ids, truth, obs


def topo_sort(rows):
    by = {int(r[0]): r for r in rows}
    out, done = [], set()

    def visit(k):
        stack = [k]
        while stack:
            x = stack[-1]
            if x in done:
                stack.pop()
                continue
            pend = [int(p) for p in by[x][1:3] if p > 0 and int(p) in by and int(p) not in done]
            if pend:
                stack.extend(pend)
            else:
                done.add(x)
                out.append(by[x])
                stack.pop()
    for k in by:
        visit(k)
    return np.array(out, dtype=np.int64)



# pedigree/error_checker.py main(): per-animal Mendel tallies in the reference's float16 arithmetic
ERRCHK_CASES = {
    "errchk_a": dict(ped=dict(n_animals=60, n_generations=5, sire_frac=0.3, dam_frac=0.6, seed=7),
                     n_snps=30, err=0.03, miss=0.03, drop=0.0, kids="all"),
    "errchk_b": dict(ped=dict(n_animals=96, n_generations=6, sire_frac=0.25, dam_frac=0.5, seed=11,
                              related_mating_p=0.6, unknown_parents_frac=0.1),
                     n_snps=32, err=0.02, miss=0.05, drop=0.12, kids="odd"),
    # long rows: the float16 accumulator passes 256 and loses low bits, so the summation order is visible
    "errchk_long": dict(ped=dict(n_animals=48, n_generations=4, sire_frac=0.3, dam_frac=0.6, seed=5),
                        n_snps=1500, err=0.3, miss=0.02, drop=0.0, kids="all"),
    # the reference's own example pedigree: 4 560 animals, 4 235 kids with both parents genotyped (104 s of reference time)
    "errchk_example": dict(ped="example.fam", n_snps=12, err=0.01, miss=0.0, drop=0.0, kids="all"),
}


def make_errchk(name):
    case = ERRCHK_CASES[name]
    ped_rows, ids, truth, obs = build_inputs(case)
    kids = [int(x) for x in ids] if case["kids"] == "all" else [int(x) for x in ids if int(x) % 2 == 1]
    t0 = time.time()
    r = rh.run_error_checker(ped_rows, ids, obs, kids)
    dt = time.time() - t0
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ped_rows=ped_rows, ids=ids, obs=obs.astype(np.uint8),
                        kids=np.array(kids, dtype=np.int64), with_parents=r["with_parents"],
                        probs_errors=r["probs_errors"].view(np.uint16), sums=r["sums"].view(np.uint16),
                        reference_seconds=dt)
    print("%s: %d animals x %d SNPs, %d trio kids, reference %.1f s" % (name, len(ids), obs.shape[1],
                                                                       len(r["with_parents"]), dt))


# empirical/preindex.py main(): Mendel-filtered LD index (quantile cut, mask, window frequencies)
PREIDX_CASES = {
    "preidx_a": dict(ped=dict(n_animals=60, n_generations=5, sire_frac=0.3, dam_frac=0.6, seed=7),
                     n_snps=30, err=0.03, miss=0.03, drop=0.0, surround=1, q=0.9, seed_chunk=10000, split=15),
    "preidx_b": dict(ped=dict(n_animals=96, n_generations=6, sire_frac=0.25, dam_frac=0.5, seed=11,
                              related_mating_p=0.6, unknown_parents_frac=0.1),
                     n_snps=40, err=0.05, miss=0.05, drop=0.12, surround=2, q=0.8, seed_chunk=10000, split=22),
    # four seed chunks: the printed cut / kept counts are per chunk, the pickled index is the LAST chunk's (:339-346)
    "preidx_c": dict(ped=dict(n_animals=60, n_generations=5, sire_frac=0.3, dam_frac=0.6, seed=7),
                     n_snps=30, err=0.03, miss=0.03, drop=0.0, surround=1, q=0.9, seed_chunk=15, split=15),
}


def make_preidx(name):
    case = PREIDX_CASES[name]
    ped_rows, ids, truth, obs = build_inputs(case)
    chrom = ["chrA" if i < case["split"] else "chrB" for i in range(case["n_snps"])]
    t0 = time.time()
    res, stdout = rh.run_preindex(ped_rows, ids, obs, chrom, surround=case["surround"], init_filter_p=case["q"],
                                  seedskidschunk=case["seed_chunk"])
    dt = time.time() - t0
    cut = [float(x) for x in re.findall(r"cuttoff [0-9.]+-quantile = (\S+)", stdout)]
    kept = [(int(a), int(b)) for a, b in re.findall(r"calc empirical ld on genotype with (\d+) of (\d+)", stdout)]
    arrays = {}
    for ci, c in enumerate(sorted(res)):
        w = max(len(f) for f in res[c]["freq"])
        tab = np.full((len(res[c]["freq"]), w), np.nan)
        for i, f in enumerate(res[c]["freq"]):
            tab[i, :len(f)] = f
        arrays["freq_" + c] = tab
        arrays["wlen_" + c] = np.array([len(x) for x in res[c]["windows"]])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ped_rows=ped_rows, ids=ids, obs=obs.astype(np.uint8),
                        chrom=np.array(chrom), surround=case["surround"], q=case["q"], seed_chunk=case["seed_chunk"],
                        quantQ=np.array(cut), kept=np.array(kept), chroms=np.array(sorted(res)),
                        reference_seconds=dt, **arrays)
    print("%s: %d animals x %d SNPs, cut %s, kept %s, reference %.1f s" % (name, len(ids), obs.shape[1], cut, kept, dt))

# correctMatrix WITH the empirical (LD-window) blend of the ranks (:619-656): the reference builds the index itself
EMP_CASES = {
    "emp_a": dict(ped=dict(n_animals=90, n_generations=5, sire_frac=0.3, dam_frac=0.6, seed=7),
                  n_snps=40, err=0.07, miss=0.0, drop=0.0, params=dict(GENOFIX_DEFAULTS, init_filter_p=0.8, tiethreshold=0.0),
                  emp=dict(surround=1, pseudocount=0.0001, weight_empirical=3)),
    "emp_b": dict(ped=dict(n_animals=96, n_generations=6, sire_frac=0.25, dam_frac=0.5, seed=11,
                           related_mating_p=0.6, unknown_parents_frac=0.1),
                  n_snps=26, err=0.04, miss=0.0, drop=0.12,
                  # no tie band: a tie writes a 9 and the next window lookup over that cell raises in the reference (D4)
                  params=dict(SIMULATE_DEFAULTS, tiethreshold=0.0, threshold_pair=0.3, threshold_single=0.4, init_filter_p=0.7),
                  emp=dict(surround=2, pseudocount=0.0001, weight_empirical=2)),
}


def make_case(name):
    case = CASES[name] if name in CASES else EMP_CASES[name]
    ped_rows, ids, truth, obs = build_inputs(case)
    cols = ["SNP%d" % i for i in range(case["n_snps"])]
    df = pd.DataFrame(obs, index=[int(x) for x in ids], columns=cols)
    t0 = time.time()
    r = rh.run_correct_matrix(df, ped_rows, emp=case.get("emp"),
                              truth_df=pd.DataFrame(truth, index=df.index, columns=cols), **case["params"])
    dt = time.time() - t0
    extra = {}
    if case.get("emp"):
        import itertools
        emp = case["emp"]
        idx = r["emp_index"]
        wmax = 2 * emp["surround"] + 1
        freq = np.full((len(cols), 3 ** wmax), np.nan)
        wstart = np.zeros(len(cols), dtype=np.int64)
        wlen = np.zeros(len(cols), dtype=np.int64)
        for i, c in enumerate(cols):
            win = idx.getWindow(c)
            wlen[i] = len(win)
            wstart[i] = cols.index(win[0]) if win else 0
            for k, v in enumerate(itertools.product([0, 1, 2], repeat=len(win))):
                freq[i, k] = idx.snp2frequency[c][tuple(zip(win, v))]
        rowi = {int(k): i for i, k in enumerate(ids)}
        rec = []
        for snp, res in r["emp_results"]:
            for (kid, o), v in res.items():
                rec.append([rowi[int(kid)], cols.index(snp), int(o)] + [float(x) for x in v])
        extra = dict(emp_surround=emp["surround"], emp_pseudocount=emp["pseudocount"], emp_weight=emp["weight_empirical"],
                     emp_freq=freq, emp_wstart=wstart, emp_wlen=wlen, emp_probs=np.array(rec, dtype=np.float64).reshape(-1, 6))
    m = re.search(r"setting initial filter to (\S+) quantile threshold", r["stdout"])
    test_p = float(m.group(1))
    m = re.search(r"n errors pairs\+single: not rank1 excluded (\d+) passed (\d+)", r["stdout"])
    errors_all = int(m.group(2))
    m = re.search(r"n errors pairs: not rank1 excluded (\d+) passed (\d+)", r["stdout"])
    pair_excluded, pair_passed = int(m.group(1)), int(m.group(2))
    snpi = {c: i for i, c in enumerate(cols)}
    seen = {}
    prec = []
    for s, d, res in r["pair_results"]:
        occ = seen.get((s, d), 0)
        seen[(s, d)] = occ + 1
        if res is None:
            continue
        for snp in res.prob_sire:
            prec.append([occ, s, d, snpi[snp]] + list(np.asarray(res.prob_sire[snp], dtype=float)) +
                        list(np.asarray(res.prob_dam[snp], dtype=float)))
    srec = []
    for k, res in r["single_results"]:
        if res is None:
            continue
        for snp in res.prob:
            srec.append([k, snpi[snp]] + list(np.asarray(res.prob[snp], dtype=float)))
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        ped_rows=ped_rows, ids=ids, truth=truth, obs=obs, raw=r["raw"], corrected=r["corrected"],
        test_p=np.float64(test_p), errors_all=errors_all, pair_excluded=pair_excluded, pair_passed=pair_passed,
        pair_post=np.array(prec, dtype=np.float64).reshape(-1, 10),
        single_post=np.array(srec, dtype=np.float64).reshape(-1, 5),
        params=np.array([case["params"][k] for k in ("threshold_pair", "threshold_single", "back",
                                                     "init_filter_p", "tiethreshold", "err_thresh")], dtype=np.float64), **extra)
    print("%s: %d x %d, reference correctMatrix %.1fs, %d cells changed, errors_all=%d, test_p=%r, raw NaN=%d" % (
        name, obs.shape[0], obs.shape[1], dt, int((r["corrected"] != obs).sum()), errors_all, test_p,
        int(np.isnan(r["raw"].astype(float)).sum())))


def make_kats():
    ref = rh.load()
    model, jad3, pdag = ref["model"], ref["jad3"], ref["pdag"]
    rs = np.random.Generator(np.random.PCG64(2024))
    n_cases, width = 400, 14
    kid = np.full((n_cases, width), -1, dtype=np.int8)
    par = np.full((n_cases, width), -1, dtype=np.int8)  # -1 = None
    pk = np.zeros((n_cases, 3))
    dk = np.full((n_cases, 3, width), np.nan, dtype=np.float16)
    for c in range(n_cases):
        n = int(rs.integers(1, width + 1))
        ks = rs.choice([0, 1, 2, 9], size=n, p=[.3, .3, .3, .1])
        ps = rs.choice([-1, 0, 1, 2, 9], size=n, p=[.3, .2, .2, .2, .1])
        kid[c, :n] = ks
        par[c, :n] = ps
        plist = [None if p < 0 else int(p) for p in ps]
        pk[c] = np.asarray(model.generate_probs_kids(list(ks), plist), dtype=float)
        inc = [k in (0, 1, 2) for k in ks]
        ks_i = [int(k) for k, i in zip(ks, inc) if i]
        ps_i = [p for p, i in zip(plist, inc) if i]
        for o in range(3):
            d = model.generate_probs_differences_kids(ks_i, ps_i, o) if ks_i else []
            dk[c, o, :len(d)] = d
    dp = np.zeros((4, 4, 3))
    for i, p1 in enumerate([0, 1, 2, 9]):
        for j, p2 in enumerate([0, 1, 2, 9]):
            for o in range(3):
                dp[i, j, o] = model.generate_probs_differences_parent(p1, p2, o)
    # countJointFrq + windows
    cj = {}
    for w in (3, 5, 7):
        tab = rs.integers(0, 3, size=(300, w)).astype(np.uint8)
        mask = rs.random((300, w)) < 0.97
        import itertools
        sv = list(itertools.product([0, 1, 2], repeat=w))
        names = ["S%d" % i for i in range(w)]
        res = jad3.JointAllellicDistribution.countJointFrq(tab, mask, names, sv, 1e-4)
        cj["cj_tab_%d" % w] = tab
        cj["cj_mask_%d" % w] = mask
        cj["cj_val_%d" % w] = np.array([res[tuple(zip(names, v))] for v in sv])
    wins = []
    import contextlib
    import io
    for n in (5, 12):
        names = ["SNP%d" % i for i in range(n)]
        for sur in (1, 2, 3):
            with contextlib.redirect_stdout(io.StringIO()):
                j = jad3.JointAllellicDistribution("1", names, surround_size=sur)
            for i, nm in enumerate(names):
                win = j.getWindow(nm)
                a = names.index(win[0]) if win else -1
                wins.append([n, sur, i, a, a + len(win)])
    # pedigree structure on example.fam
    ped = pdag.PedigreeDAG.from_file("/root/reference/examples/simulate/example.fam")
    gens = sorted((g, len(k)) for g, k in ped.generationIndex.items())
    animals = sorted(ped.males | ped.females)
    pick = [animals[i] for i in rs.choice(len(animals), size=120, replace=False)]
    gen_of = {a: g for g, ks in ped.generationIndex.items() for a in ks}
    struct = []
    for a in pick:
        for depth in (2, 3):
            anc = [x for x in ped.get_parents_depth([a], depth) if x is not None]
            if anc:
                nodes = sorted(ped.balance_nodes(list(set([a] + list(anc)))))
            else:
                nodes = []
            struct.append([a, depth, gen_of[a], len(anc), len(nodes)] + nodes + [0] * (40 - len(nodes)))
    np.savez_compressed(os.path.join(OUT, "kats.npz"), kid=kid, par=par, probs_kids=pk, diff_kids=dk,
                        diff_parent=dp, windows=np.array(wins), gen_sizes=np.array(gens),
                        blanket_struct=np.array(struct, dtype=np.int64),
                        n_animals=len(animals), n_males=len(ped.males), n_females=len(ped.females), **cj)
    print("kats written")


# --------------------------------------------------------------------------------------------
# round 2: surfaces that had no pinned answers yet (split_pedigree, getCountTable, getEmpProbs)
# --------------------------------------------------------------------------------------------
def _no_lone_pedigree(n, gens, seed, **kw):
    """synthetic pedigree without animals that are neither parents nor have parents (split_pedigree raises on those)"""
    from genofix_b200 import synth
    rows = synth.make_pedigree(n, gens, seed=seed, **kw)
    parents = set(rows[:, 1]) | set(rows[:, 2])
    keep = [(r[1] > 0 or r[2] > 0 or r[0] in parents) for r in rows]
    return rows[np.array(keep)]


def make_kats2():
    """tests/golden/kats2.npz: PedigreeDAG.split_pedigree (pedigree/pedigree_dag.py:41-122) on example.fam and two synthetic
    pedigrees, JointAllellicDistribution.getCountTable (empirical/jalleledist3.py:101-117) on indexes counted by the
    reference's countJointFrqAll, and CorrectGenotypes.getEmpProbs (correct_genotypes3.py:69-82)."""
    import itertools
    import multiprocessing
    import pandas as pd
    from genofix_b200 import synth
    ref = rh.load()
    pdag, jad3, cg3 = ref["pdag"], ref["jad3"], ref["cg3"]
    out = {}
    # ---- split_pedigree: cluster number of every member, in the reference's list order ----
    peds = {"example": (np.loadtxt(os.path.join(rh.REF_SRC, "..", "examples", "simulate", "example.fam"), dtype=np.int64,
                                   usecols=(0, 1, 2, 3)), [100, 10]),
            "syn_a": (_no_lone_pedigree(1500, 8, 5, sire_frac=0.1, dam_frac=0.5), [100, 30]),
            "syn_b": (_no_lone_pedigree(800, 6, 9, sire_frac=0.3, dam_frac=0.6, related_mating_p=0.3), [50])}
    for name, (rows, sizes) in peds.items():
        out["split_rows_" + name] = rows
        for mcs in sizes:
            cl = pdag.PedigreeDAG.from_table(rows).split_pedigree(mcs)
            flat = np.array([(i, int(a)) for i, c in enumerate(cl) for a in sorted(c)], dtype=np.int64)
            out["split_%s_%d" % (name, mcs)] = flat
            print("split", name, mcs, len(cl), "clusters")
    lone = synth.make_pedigree(300, 4, seed=3)
    out["split_rows_lone"] = lone
    try:
        pdag.PedigreeDAG.from_table(lone).split_pedigree(100)
        out["split_lone_error"] = np.array("")
    except Exception as e:  # noqa: B902 -- the message is the pinned behaviour
        out["split_lone_error"] = np.array(str(e))
    # ---- LD index counted by the reference, getCountTable / getEmpProbs on it ----
    rs = np.random.Generator(np.random.PCG64(77))
    n, s = 250, 14
    names = ["SNP%d" % i for i in range(s)]
    base = rs.integers(0, 3, size=(n, 1))
    G = np.where(rs.random((n, s)) < 0.6, base, rs.integers(0, 3, size=(n, s))).astype(np.uint8)   # some LD
    df = pd.DataFrame(G, index=np.arange(1, n + 1), columns=names)
    out["emp_G"] = G
    for sur in (1, 2, 3):
        with contextlib.redirect_stdout(io.StringIO()), rh.deterministic_pools():
            emp = jad3.JointAllellicDistribution("1", names, surround_size=sur)
            emp.countJointFrqAll(df, threads=1)
        freq = np.zeros((s, 3 ** (2 * sur + 1)))
        wlen = np.zeros(s, dtype=np.int64)
        for i, snp in enumerate(names):
            win = emp.getWindow(snp)
            wlen[i] = len(win)
            for k, v in enumerate(itertools.product([0, 1, 2], repeat=len(win))):
                freq[i, k] = emp.snp2frequency[snp][tuple(zip(win, v))]
        out["emp_freq_%d" % sur] = freq
        out["emp_wlen_%d" % sur] = wlen
        # getCountTable for every (animal of the first 40, SNP): complete evidence (no 9: the reference raises there, D4)
        ct = np.zeros((40, s, 3))
        for a in range(40):
            ev = {snp: int(G[a, j]) for j, snp in enumerate(names)}
            for j, snp in enumerate(names):
                ct[a, j] = emp.getCountTable(ev, snp)
        out["emp_counttable_%d" % sur] = ct
        # getEmpProbs on the window chunk of every SNP (D3: the worker state is installed directly, initializerEmp
        # copies a `frequency` attribute jalleledist3 objects do not have)
        multiprocessing.current_process().indexemp = emp
        ep = np.full((s, n, 3), np.nan)
        for j, snp in enumerate(names):
            win = emp.getWindow(snp)
            if snp not in win:
                continue
            res = cg3.CorrectGenotypes.getEmpProbs(df.loc[:, win].copy(deep=True), snp)
            for (kid, _obs), v in res.items():
                ep[j, kid - 1] = v
        out["emp_probs_%d" % sur] = ep
    np.savez_compressed(os.path.join(OUT, "kats2.npz"), **out)
    print("kats2 written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or ["kats"] + list(CASES) + list(ERRCHK_CASES) + list(PREIDX_CASES)
    for n in names:
        if n == "kats":
            make_kats()
        elif n == "kats2":
            make_kats2()
        elif n in ERRCHK_CASES:
            make_errchk(n)
        elif n in PREIDX_CASES:
            make_preidx(n)
        else:
            make_case(n)
