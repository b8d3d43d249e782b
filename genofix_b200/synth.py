"""Synthetic pedigrees and genotype matrices of the shapes named in BASELINE.json.

Input generators only ("Synthetic inputs").  They follow the shape of what
the reference's simulator produces (src/simulate/simulate.py:187-230 gene drop with free
recombination, :317-331 error injection) but are independent numpy code: the reference's
gene-drop RNG is clock-seeded (src/quickgsim/__init__.py:15-30) so its outputs cannot be
reproduced anyway.  All RNG is np.random.Generator(PCG64(seed)).
"""
import math

import numpy as np


def make_pedigree(n_animals, n_generations, sire_frac=0.05, dam_frac=0.5, seed=1234,
                  related_mating_p=0.0, unknown_parents_frac=0.0):
    """Return int64 rows [kid, sire, dam, sex] with ids 1..n_animals, generation-ordered.

    Generation 0 are founders (sire = dam = 0).  Each later generation picks `sire_frac` of the
    previous generation's males and `dam_frac` of its females; every selected dam is mated to
    one selected sire and litter sizes are multinomial.  `related_mating_p` is the probability
    that a dam is mated to a sire that shares a grand-parent with her (forces blanket loops);
    `unknown_parents_frac` of the non-founders get BOTH parents set unknown (the only
    missing-parent form inside the reference's domain).
    """
    rs = np.random.Generator(np.random.PCG64(seed))
    per_gen = int(math.ceil(n_animals / n_generations))
    rows = np.zeros((n_animals, 4), dtype=np.int64)
    rows[:, 0] = np.arange(1, n_animals + 1)
    sex = rs.integers(1, 3, size=n_animals)
    rows[:, 3] = sex
    gen_of = np.minimum(np.arange(n_animals) // per_gen, n_generations - 1)
    sire = np.zeros(n_animals, dtype=np.int64)
    dam = np.zeros(n_animals, dtype=np.int64)
    for g in range(1, n_generations):
        prev = np.nonzero(gen_of == g - 1)[0]
        cur = np.nonzero(gen_of == g)[0]
        if len(cur) == 0:
            continue
        males = prev[sex[prev] == 1]
        females = prev[sex[prev] == 2]
        if len(males) == 0 or len(females) == 0:
            # degenerate tiny generation: flip one animal's sex so breeding can continue
            if len(males) == 0:
                sex[prev[0]] = 1
            else:
                sex[prev[0]] = 2
            rows[:, 3] = sex
            males = prev[sex[prev] == 1]
            females = prev[sex[prev] == 2]
        n_s = max(1, int(round(len(males) * sire_frac)))
        n_d = max(1, int(round(len(females) * dam_frac)))
        sires = rs.choice(males, size=n_s, replace=False)
        dams = rs.choice(females, size=n_d, replace=False)
        mate = rs.choice(sires, size=n_d, replace=True)
        if related_mating_p > 0 and g >= 3:
            # grand-parent sets (by index); 0-index animal ids are idx+1
            def gps(a):
                out = set()
                for p in (sire[a], dam[a]):
                    if p > 0:
                        for q in (sire[p - 1], dam[p - 1]):
                            if q > 0:
                                out.add(q)
                return out
            sire_gps = [gps(s) for s in sires]
            for i, d in enumerate(dams):
                if rs.random() < related_mating_p:
                    dg = gps(d)
                    cand = [sires[k] for k in range(n_s) if sire_gps[k] & dg]
                    if cand:
                        mate[i] = cand[int(rs.integers(0, len(cand)))]
        litter = rs.multinomial(len(cur), np.full(n_d, 1.0 / n_d))
        d_of_kid = np.repeat(np.arange(n_d), litter)
        sire[cur] = mate[d_of_kid] + 1
        dam[cur] = dams[d_of_kid] + 1
    if unknown_parents_frac > 0:
        nonf = np.nonzero(sire > 0)[0]
        k = int(round(len(nonf) * unknown_parents_frac))
        drop = rs.choice(nonf, size=k, replace=False)
        sire[drop] = 0
        dam[drop] = 0
    rows[:, 1] = sire
    rows[:, 2] = dam
    return rows


def gene_drop(rows, n_snps, seed=1234, snp_block=4096):
    """Per-SNP independent Mendelian gene drop.  rows must list parents before offspring.

    Returns uint8 [n_animals, n_snps] genotypes in row order (0/1/2)."""
    rs = np.random.Generator(np.random.PCG64(seed))
    n = rows.shape[0]
    idx = {int(k): i for i, k in enumerate(rows[:, 0])}
    sire_i = np.array([idx.get(int(s), -1) if s > 0 else -1 for s in rows[:, 1]], dtype=np.int64)
    dam_i = np.array([idx.get(int(d), -1) if d > 0 else -1 for d in rows[:, 2]], dtype=np.int64)
    # depth levels so a whole level is vectorised
    level = np.zeros(n, dtype=np.int64)
    for i in range(n):
        l = 0
        if sire_i[i] >= 0:
            l = max(l, level[sire_i[i]] + 1)
        if dam_i[i] >= 0:
            l = max(l, level[dam_i[i]] + 1)
        level[i] = l
    out = np.empty((n, n_snps), dtype=np.uint8)
    for s0 in range(0, n_snps, snp_block):
        s1 = min(n_snps, s0 + snp_block)
        w = s1 - s0
        pat = np.zeros((n, w), dtype=np.uint8)
        mat = np.zeros((n, w), dtype=np.uint8)
        for l in range(int(level.max()) + 1):
            a = np.nonzero(level == l)[0]
            for hap, par in ((pat, sire_i), (mat, dam_i)):
                p = par[a]
                known = p >= 0
                draw = rs.integers(0, 2, size=(len(a), w), dtype=np.uint8)
                if known.any():
                    ak = a[known]
                    pk = p[known]
                    pick = draw[known].astype(bool)
                    hap[ak] = np.where(pick, pat[pk], mat[pk])
                if (~known).any():
                    hap[a[~known]] = draw[~known]
        out[:, s0:s1] = pat + mat
    return out


def inject_errors(G, rate=0.01, seed=1234):
    """ceil(N*S*rate) cells drawn with replacement; new value uniform over the two other
    genotypes (shape of src/simulate/simulate.py:318-331).  Returns a new array."""
    rs = np.random.Generator(np.random.PCG64(seed))
    n, s = G.shape
    k = int(math.ceil(n * s * rate))
    r = rs.integers(0, n, size=k)
    c = rs.integers(0, s, size=k)
    shift = rs.integers(1, 3, size=k).astype(np.uint8)
    out = G.copy()
    v = out[r, c]
    ok = v <= 2
    # a cell drawn twice keeps the last draw (the reference mutates it twice; either way it
    # ends up different from the truth with the same distribution)
    out[r[ok], c[ok]] = (v[ok] + shift[ok]) % 3
    return out


def inject_missing(G, rate=0.05, seed=4321):
    rs = np.random.Generator(np.random.PCG64(seed))
    out = G.copy()
    m = rs.random(G.shape) < rate
    out[m] = 9
    return out


def complete_rows(rows):
    """Pedigree rows with a founder row (sire = dam = 0) added for every animal that only appears as a parent,
    ordered parents before offspring -- what gene_drop needs (the reference's example.fam lists kids only)."""
    rows = np.asarray(rows, dtype=np.int64)
    known = set(int(x) for x in rows[:, 0])
    extra = []
    for col, sex in ((1, 1), (2, 2)):
        for p in np.unique(rows[:, col]):
            if p > 0 and int(p) not in known:
                extra.append((int(p), 0, 0, sex))
                known.add(int(p))
    allrows = np.array(extra + [tuple(int(v) for v in r[:4]) for r in rows], dtype=np.int64)
    by = {int(r[0]): r for r in allrows}
    out, done = [], set()
    for k in by:
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
    return np.array(out, dtype=np.int64)
