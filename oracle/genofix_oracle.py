"""CPU oracle: a plain numpy / Python restatement of GenoFix's Mendelian-rank + correction path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import this module; the product (genofix_b200/) never does.

Parity status: PINNED by execution.  The reference's own tests pin nothing on this path
(SURVEY.md section 4), so tests/golden/*.npz hold outputs of the UNMODIFIED reference run in the
build container through oracle/ref_harness.py (pgmpy replaced by the exact-inference shim in
oracle/shim -- that inference routine is the only restated arithmetic behind the goldens), and
tests/test_oracle_golden.py checks this restatement against them bit for bit.

Every function cites the reference lines it follows (paths relative to /root/reference/src).
Deliberate, documented deviations (both are *fixings* of under-specified behaviour):
  * pairs of a generation are applied in ascending (sire id, dam id) order, singles in ascending id;
  * correctSingle's tie set is the flat 1-D set.
"""
import itertools
import warnings

import numpy as np

# bayesnet/model.py:24-30
ALLELE_PRIOR = np.array([0.25, 0.5, 0.25])
MENDEL = np.array([[1, .5, 0, .5, .25, 0, 0, 0, 0],
                   [0, .5, 1, .5, .5, .5, 1, .5, 0],
                   [0, 0, 0, 0, .25, .5, 0, .5, 1]], dtype=float)
_SIRE_AXIS = np.array([0, 0, 0, 1, 1, 1, 2, 2, 2])
_DAM_AXIS = np.array([0, 1, 2, 0, 1, 2, 0, 1, 2])
_T3 = MENDEL.reshape(3, 3, 3)  # [kid, sire, dam]


# --------------------------------------------------------------------------------------------
# bayesnet/model.py:65-133  (children heuristic and score pieces; the reference's own numpy code)
# --------------------------------------------------------------------------------------------
def _kid_rows(kidstates, oneparentstates):
    """model.py:73-84 -- per usable child, the Mendel row masked by the other parent, row-normalised."""
    kidstates = np.asarray(kidstates)
    oneparentstates = np.asarray(oneparentstates)
    include = [x in [0, 1, 2] for x in kidstates]
    if np.count_nonzero(include) == 0:
        return None
    table = np.array([MENDEL[x] for x in kidstates[include]])
    for i, parent in enumerate(oneparentstates[include]):
        if parent is not None and parent in [0, 1, 2]:
            table[i, np.not_equal(_SIRE_AXIS, parent)] = 0
        rowsum = np.sum(table[i, ])
        if rowsum > 0:
            table[i, ] = np.divide(table[i, ], rowsum)
    return table


def probs_kids(kidstates, oneparentstates):
    """model.py:65-89 generate_probs_kids."""
    table = _kid_rows(kidstates, oneparentstates)
    if table is None:
        return [1 / 3, 1 / 3, 1 / 3]
    res = [np.nanmean(table[:, np.equal(_DAM_AXIS, x)]) for x in [0, 1, 2]]
    if sum(res) > 0:
        res = res / sum(res)
    return res


def probs_differences_kids(kidstates, oneparentstates, observedstate):
    """model.py:93-116 generate_probs_differences_kids (float16 table)."""
    stat = np.zeros((len(kidstates), 3), np.float16)
    table = _kid_rows(kidstates, oneparentstates)
    if table is None:
        return []
    for i in range(table.shape[0]):
        stat[i, :] = [np.nansum(table[i, np.equal(_DAM_AXIS, x)]) for x in [0, 1, 2]]
    return [np.nanmax(row) - row[observedstate] for row in stat]


def probs_differences_parent(parent1, parent2, observedstate):
    """model.py:118-133 generate_probs_differences_parent."""
    if parent1 in [0, 1, 2] and parent2 in [0, 1, 2]:
        column = _T3[:, parent1, parent2]
    elif parent1 in [0, 1, 2]:
        column = np.nanmean(MENDEL[:, np.equal(_SIRE_AXIS, parent1)], 1, dtype=float)
    elif parent2 in [0, 1, 2]:
        column = np.nanmean(MENDEL[:, np.equal(_DAM_AXIS, parent2)], 1, dtype=float)
    else:
        return 0
    return np.nanmax(column) - column[observedstate]


# --------------------------------------------------------------------------------------------
# pedigree/pedigree_dag.py  (flat restatement)
# --------------------------------------------------------------------------------------------
class OraclePedigree(object):
    """pedigree_dag.py:188-214 from_table, :171-175 recursiveDepth, :221-231 get_parents_depth,
    :124-168 balance_nodes/get_subset, :245-274 get_parents/get_kids."""

    def __init__(self, rows):
        rows = np.asarray(rows, dtype=np.int64)
        self.sire = {}
        self.dam = {}
        self.kids = {}  # parent -> children in edge-insertion order (networkx successors order)
        self.males = set()
        self.females = set()
        for kid, s, d, sex in rows:
            kid, s, d, sex = int(kid), int(s), int(d), int(sex)
            if s > 0:
                self.males.add(s)
            if d > 0:
                self.females.add(d)
            if kid > 0:
                (self.males if sex == 1 else self.females if sex == 2 else set()).add(kid)
                if s > 0:
                    self.sire[kid] = s
                    lst = self.kids.setdefault(s, [])
                    if kid not in lst:
                        lst.append(kid)
                if d > 0:
                    self.dam[kid] = d
                    lst = self.kids.setdefault(d, [])
                    if kid not in lst:
                        lst.append(kid)
        self.animals = self.males | self.females
        self.generation = {}
        for a in self.animals:
            self.generation[a] = min(self._line(a, self.sire), self._line(a, self.dam))
        self.generation_index = {}
        for a, g in self.generation.items():
            self.generation_index.setdefault(g, set()).add(a)

    @staticmethod
    def _line(a, d):
        n = 0
        while a in d:
            a = d[a]
            n += 1
        return n

    def parents(self, kid):
        return (self.sire.get(kid), self.dam.get(kid))

    def children(self, parent):
        return self.kids.get(parent, [])

    def ancestors(self, focal, depth):
        """Level-by-level ancestors, duplicates kept (pedigree_dag.py:221-231)."""
        out = []
        level = list(focal)
        for _ in range(depth):
            nxt = []
            for k in level:
                for p in self.parents(k):
                    if p is not None:
                        nxt.append(p)
            if not nxt:
                break
            out.extend(nxt)
            level = nxt
        return out

    def blanket(self, focal, depth):
        """Node set of the sub-network: focal + ancestors, closed under 'partner of an included
        parent' (pedigree_dag.py:124-139).  Raises KeyError for single-known-parent nodes like the
        reference (D9).  Returns (nodes, parent map restricted to the set)."""
        nodes = list(dict.fromkeys(list(focal) + self.ancestors(focal, depth)))
        changed = True
        while changed:
            changed = False
            for node in [n for n in nodes if n in self.sire or n in self.dam]:
                s = self.sire[node]
                d = self.dam[node]
                if s in nodes and d not in nodes:
                    nodes.append(d)
                    changed = True
                if d in nodes and s not in nodes:
                    nodes.append(s)
                    changed = True
        nset = set(nodes)
        par = {}
        for k in nodes:
            s = self.sire.get(k)
            d = self.dam.get(k)
            if s in nset and d in nset:
                par[k] = (s, d)
            elif (s in nset) != (d in nset):
                raise KeyError(k)
        return nodes, par


# --------------------------------------------------------------------------------------------
# exact inference on a blanket (stands where the reference calls pgmpy, correct_genotypes3.py:119-121,
# 242-245, 388-390; model construction bayesnet/model.py:173-230)
# --------------------------------------------------------------------------------------------
def blanket_marginals(nodes, par, query, evidence):
    """P(q | evidence) for each q in query on the blanket network: founders-of-the-subset get the
    allele prior, others the 3x3x3 Mendel CPD.  Evidence-reduced factors with empty scope are
    dropped; the result is divided by its sum (0/0 -> NaN).  Sum-product variable elimination."""
    in_model = set(par.keys())
    for k, (s, d) in par.items():
        in_model.add(s)
        in_model.add(d)
    factors = []  # (vars tuple, ndarray)
    for v in nodes:
        if v not in in_model:
            continue
        if v in par:
            vs = [v, par[v][0], par[v][1]]
            t = _T3
        else:
            vs = [v]
            t = ALLELE_PRIOR
        idx = []
        keep = []
        for x in vs:
            if x in evidence:
                idx.append(int(evidence[x]))
            else:
                idx.append(slice(None))
                keep.append(x)
        if not keep:
            continue
        t = t[tuple(idx)]
        if len(set(keep)) != len(keep):  # sire == dam cannot happen in a well-formed pedigree
            raise ValueError("malformed blanket")
        factors.append((tuple(keep), t))
    # restrict to the component connected to the query (everything else would reduce to dropped scalars)
    comp = set(query)
    grew = True
    while grew:
        grew = False
        for vs, _ in factors:
            if comp & set(vs) and not set(vs) <= comp:
                comp |= set(vs)
                grew = True
    factors = [f for f in factors if comp & set(f[0])]
    elim = [v for v in comp if v not in query]

    def contract(fs, out_vars):
        local = {}
        args = []
        for vs, t in fs:
            args.append(t)
            args.append([local.setdefault(x, len(local)) for x in vs])
        return np.einsum(*args, [local.setdefault(x, len(local)) for x in out_vars])
    while elim:
        # fewest neighbours first
        def deg(x):
            s = set()
            for vs, _ in factors:
                if x in vs:
                    s |= set(vs)
            return len(s)
        z = min(sorted(elim), key=deg)
        elim.remove(z)
        inv = [f for f in factors if z in f[0]]
        factors = [f for f in factors if z not in f[0]]
        out_vars = tuple(dict.fromkeys(x for vs, _ in inv for x in vs if x != z))
        t = contract(inv, out_vars)
        if out_vars:
            factors.append((out_vars, t))
    res = {}
    qv = list(query)
    joint = contract(factors, qv) if factors else np.ones([3] * len(qv))
    for i, q in enumerate(qv):
        m = joint.sum(axis=tuple(a for a in range(len(qv)) if a != i)) if len(qv) > 1 else joint
        with np.errstate(invalid="ignore", divide="ignore"):
            res[q] = np.asarray(m / m.sum(), dtype=float)
    return res


def blanket_marginals_bruteforce(nodes, par, query, evidence):
    """Independent check of blanket_marginals by explicit enumeration (small blankets only)."""
    in_model = set(par.keys())
    for k, (s, d) in par.items():
        in_model.update((s, d))
    free = [v for v in nodes if v in in_model and v not in evidence]
    # component of the query
    comp = set(query)
    grew = True
    fam = [([v, par[v][0], par[v][1]] if v in par else [v]) for v in nodes if v in in_model]
    while grew:
        grew = False
        for vs in fam:
            u = set(x for x in vs if x not in evidence)
            if u & comp and not u <= comp:
                comp |= u
                grew = True
    free = [v for v in free if v in comp]
    fam = [vs for vs in fam if set(x for x in vs if x not in evidence) & comp]
    joint = np.zeros([3] * len(query))
    for states in itertools.product([0, 1, 2], repeat=len(free)):
        a = dict(evidence)
        a.update(zip(free, states))
        p = 1.0
        for vs in fam:
            if len(vs) == 3:
                p *= _T3[a[vs[0]], a[vs[1]], a[vs[2]]]
            else:
                p *= ALLELE_PRIOR[a[vs[0]]]
            if p == 0.0:
                break
        joint[tuple(a[q] for q in query)] += p
    res = {}
    for i, q in enumerate(query):
        m = joint.sum(axis=tuple(x for x in range(len(query)) if x != i)) if len(query) > 1 else joint
        with np.errstate(invalid="ignore", divide="ignore"):
            res[q] = m / m.sum()
    return res


# --------------------------------------------------------------------------------------------
# correct_genotypes3.py
# --------------------------------------------------------------------------------------------
class Oracle(object):
    def __init__(self, ped_rows, ids):
        self.ped = ped_rows if isinstance(ped_rows, OraclePedigree) else OraclePedigree(ped_rows)
        self.ids = [int(x) for x in ids]
        self.row = {k: i for i, k in enumerate(self.ids)}

    # ---- correct_genotypes3.py:85-166 ------------------------------------------------------
    def mendel_rank_animal(self, G, kid, back, want_probs=False):
        ped = self.ped
        S = G.shape[1]
        raw = np.ones(S, dtype=np.float16)
        probs_out = np.zeros((S, 3), dtype=np.float16) if want_probs else None
        anc_list = ped.ancestors([kid], back)
        nodes = par = None
        model_nodes = set()
        if len(anc_list) > 0:
            nodes, par = ped.blanket([kid], back)
            model_nodes = set(par.keys())
            for s, d in par.values():
                model_nodes.update((s, d))
        ev_nodes = [x for x in model_nodes if x != kid and x in self.row]
        kids_g = [c for c in ped.children(kid) if c in self.row]
        cache = {}
        krow = self.row[kid]
        for j in range(S):
            obs = G[krow, j]
            anc = anc_err = None
            if len(anc_list) > 0:
                ev = {x: int(G[self.row[x], j]) for x in ev_nodes if G[self.row[x], j] != 9}
                key = tuple(sorted(ev.items()))
                if key not in cache:
                    cache[key] = np.array(blanket_marginals(nodes, par, [kid], ev)[kid], dtype=np.float16)
                anc = cache[key]
                if obs in (0, 1, 2):
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        anc_err = np.nanmax(anc) - anc[obs]
            states = np.array([G[self.row[c], j] for c in kids_g])
            kid_err = None
            kprobs = None
            usable = [x in (0, 1, 2) for x in states]
            if len(kids_g) > 0 and any(usable):
                sk = states[usable]
                sp = [None] * len(sk)  # D7: other parent never filled (:137-141)
                if want_probs:
                    kprobs = probs_kids(sk, sp)
                if obs in (0, 1, 2):
                    kid_err = probs_differences_kids(sk, sp, obs)
            if obs in (0, 1, 2):
                if anc_err is not None and kid_err is not None:
                    raw[j] = np.sum(kid_err, dtype=np.float16) + np.multiply(anc_err, 2, dtype=np.float16)
                elif anc_err is not None:
                    raw[j] = np.multiply(anc_err, 2, dtype=np.float16)
                elif kid_err is not None:
                    raw[j] = np.sum(kid_err, dtype=np.float16)
            if want_probs:
                if anc is not None and kprobs is not None:
                    probs_out[j] = np.multiply(anc, kprobs, dtype=np.float16)
                elif anc is not None:
                    probs_out[j] = anc
                elif kprobs is not None:
                    probs_out[j] = np.array(kprobs, dtype=np.float16)
        return (raw, probs_out) if want_probs else raw

    def mendel_rank(self, G, back=2):
        raw = np.ones(G.shape, dtype=np.float16)
        for kid in self.ids:
            raw[self.row[kid]] = self.mendel_rank_animal(G, kid, back)
        return raw

    # ---- correct_genotypes3.py:593-606,666-667 -----------------------------------------------
    @staticmethod
    def transform(raw, G, init_filter_p):
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            E = np.log(np.log(raw.astype(np.float64) + 1) + 1)
            E = E / np.nanmax(E)
            E[G == 9] = 1
            q = np.quantile(E.flatten(), [0.95, 0.99, init_filter_p])
        return E, q[2]

    # ---- helpers shared by correctPair / correctSingle ---------------------------------------
    def _kids_term(self, G, focal, j, require_genotyped_other):
        """:256-278 (pairs) and :398-408 (singles): children set order, other parent's observed state."""
        ped = self.ped
        kids = [c for c in set(list(ped.children(focal))) if c in self.row]
        if len(kids) == 0:
            return None
        sk = [G[self.row[c], j] for c in kids]
        sp = [None] * len(kids)
        for i, c in enumerate(kids):
            other = [x for x in ped.parents(c) if x is not None and x != focal]
            if len(other) == 1 and (other[0] in self.row or not require_genotyped_other):
                sp[i] = G[self.row[other[0]], j]
        return probs_kids(sk, sp)

    @staticmethod
    def _combine(anc, kids):
        """:280-304 / :410-422."""
        if anc is not None and kids is not None:
            p = np.multiply(anc, kids)
        elif anc is not None:
            p = anc
        elif kids is not None:
            p = kids
        else:
            return None
        with np.errstate(invalid="ignore"):
            if np.sum(p) > 0:
                p = np.divide(p, np.sum(p))
        return np.asarray(p, dtype=float)

    # ---- correct_genotypes3.py:169-327 -------------------------------------------------------
    def correct_pair(self, G, E, sire, dam, back, test_p, suspect_t=0.2, cache=None):
        ped = self.ped
        rs, rd = self.row[sire], self.row[dam]
        with np.errstate(invalid="ignore"):
            suspect = np.nonzero((E[rs] >= test_p) | (E[rd] >= test_p))[0]
        if len(suspect) == 0:
            return None
        model = None
        if any(p is not None for p in ped.parents(sire)) and any(p is not None for p in ped.parents(dam)):
            nodes, par = ped.blanket([sire, dam], back)
            mn = set(par.keys())
            for s, d in par.values():
                mn.update((s, d))
            model = (nodes, par, mn)
        allindiv = {sire, dam} | set(ped.children(sire)) | set(ped.children(dam))
        if model is not None:
            allindiv |= model[2]
        genotyped = [x for x in allindiv if x in self.row]
        cache = {} if cache is None else cache
        out = {}
        for j in suspect:
            vals = {x: int(G[self.row[x], j]) for x in genotyped}
            os_, od = vals[sire], vals[dam]
            ep = {x: E[self.row[x], j] for x, v in vals.items() if v in (0, 1, 2)}
            if os_ in (0, 1, 2) and od in (0, 1, 2):
                pmax = max(E[rs, j], E[rd, j])
            elif os_ not in (0, 1, 2) and od not in (0, 1, 2):
                pmax = 1
            elif os_ in (0, 1, 2):
                pmax = E[rs, j]
            else:
                pmax = E[rd, j]
            sus = set(x for x, e in ep.items() if e > (pmax - suspect_t))
            anc_s = anc_d = None
            if model is not None:
                ev = {x: v for x, v in vals.items()
                      if x != sire and x != dam and v != 9 and x not in sus and x in model[2]}
                key = tuple(sorted(ev.items()))
                if key not in cache:
                    cache[key] = blanket_marginals(model[0], model[1], [sire, dam], ev)
                anc_s, anc_d = cache[key][sire], cache[key][dam]
            ks = self._kids_term(G, sire, j, True)
            kd = self._kids_term(G, dam, j, True)
            ps = self._combine(anc_s, ks)
            pd_ = self._combine(anc_d, kd)
            if ps is None or pd_ is None:
                raise Exception("illegal lone node in the pedigree graph")
            out[int(j)] = (ps, pd_, ps[os_] if os_ != 9 else np.nan, pd_[od] if od != 9 else np.nan)
        return out

    # ---- correct_genotypes3.py:330-430 -------------------------------------------------------
    def correct_single(self, G, E, kid, back, test_p, suspect_t=0.2):
        ped = self.ped
        rk = self.row[kid]
        with np.errstate(invalid="ignore"):
            suspect = np.nonzero(E[rk] >= test_p)[0]
        if len(suspect) == 0:
            return None
        model = None
        if any(p is not None for p in ped.parents(kid)):
            nodes, par = ped.blanket([kid], back)
            mn = set(par.keys())
            for s, d in par.values():
                mn.update((s, d))
            model = (nodes, par, mn)
        allindiv = {kid} | set(ped.children(kid))
        if model is not None:
            allindiv |= model[2]
        genotyped = [x for x in allindiv if x in self.row]
        cache = {}
        out = {}
        for j in suspect:
            vals = {x: int(G[self.row[x], j]) for x in genotyped}
            obs = vals[kid]
            ep = {x: E[self.row[x], j] for x, v in vals.items() if v in (0, 1, 2)}
            pmax = E[rk, j] if obs in (0, 1, 2) else 1
            sus = set(x for x, e in ep.items() if e > (pmax - suspect_t))
            anc = None
            if model is not None:
                ev = {x: v for x, v in vals.items() if x != kid and v != 9 and x not in sus and x in model[2]}
                key = tuple(sorted(ev.items()))
                if key not in cache:
                    cache[key] = blanket_marginals(model[0], model[1], [kid], ev)[kid]
                anc = cache[key]
            kt = self._kids_term(G, kid, j, False)
            p = self._combine(anc, kt)
            if p is None:
                return None  # ":418-419 illegal lone node": the whole animal is skipped
            out[int(j)] = (p, p[obs] if obs != 9 else np.nan)
        return out

    # ---- decision rule :741-744, :809-824 (pairs), :932-938 (singles) ------------------------
    @staticmethod
    def _decide(G, E, focal_row, j, p, obsprob, parent_rows, tie, threshold, err_thresh, emp=None, single=False):
        """Returns (new_code or None, counted_error, counted_excluded).  emp: the LD index (emp_index) -- the posterior is then
        replaced by its mean with the weighted window frequencies of the focal animal's LIVE row (:746-779 pairs: only the
        best states and their probability change; :883-908 singles: the observed state's probability too)."""
        obs = G[focal_row, j]  # LIVE matrix
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if emp is not None:
                norm = emp_state_probs(G[focal_row], j, emp)
                comb = np.nanmean([np.asarray(p, dtype=float), norm * emp["weight"]], 0, dtype=float)
                if single or np.nansum(comb) > 0:
                    comb = np.divide(comb, np.sum(comb))
                p = comb
                if single:
                    obsprob = 0. if obs == 9 else comb[obs]
            mx = np.nanmax(p)
        with np.errstate(invalid="ignore"):
            states = np.where(np.array(p >= (mx - tie), dtype=bool))[0]
            if obs != 9:
                cond = obs not in (0, 1, 2) or (obs not in states and ((mx - obsprob) >= threshold))
            else:
                cond = True
        if not cond:
            return None, 0, 0
        ranks = [E[r, j] for r in parent_rows]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mr = np.nanmax([0] + ranks)
        if E[focal_row, j] * err_thresh >= mr:
            if len(set(states)) == 1:
                return int(states[0]), 1, 0
            return 9, 1, 0
        return None, 0, 1

    # ---- correct_genotypes3.py:531-950 -------------------------------------------------------
    def correct_matrix(self, G, threshold_pair, threshold_single, back=2, init_filter_p=0.9,
                       tiethreshold=0.05, err_thresh=1, raw=None, keep_posteriors=False, transformed=None, emp=None):
        """transformed = (E, test_p): ranks already transformed and the threshold already taken over the WHOLE matrix
        of which G is a block of columns (correct_matrix_parallel); the correction of a column needs nothing else."""
        ped = self.ped
        G0 = np.array(G, dtype=np.uint8)
        G = G0.copy()
        if transformed is not None:
            E, test_p = transformed
        else:
            if raw is None:
                raw = self.mendel_rank(G0, back)
            E, test_p = self.transform(raw, G0, init_filter_p)
            if emp is not None:
                E, test_p = emp_blend_ranks(E, G0, emp, init_filter_p)     # :619-667
        corrected_n = np.zeros(len(self.ids), dtype=np.int32)
        excluded_n = np.zeros(len(self.ids), dtype=np.int32)
        pairs = set()
        for kid in self.ids:
            s, d = ped.parents(kid)
            if s is not None and d is not None:
                pairs.add((s, d))
        prow = {}
        for a in self.ids:
            prow[a] = [self.row[p] for p in ped.parents(a) if p is not None and p in self.row]
        posts = {}
        for gen in sorted(ped.generation_index.keys()):
            members = ped.generation_index[gen]
            gp = sorted((s, d) for s, d in pairs if (s in members or d in members)
                        and s in self.row and d in self.row)
            snap = G.copy()
            results = [(s, d, self.correct_pair(snap, E, s, d, back + 1, test_p)) for s, d in gp]
            for s, d, res in results:
                if res is None:
                    continue
                if keep_posteriors:
                    posts[(gen, s, d)] = res
                for j in sorted(res.keys()):
                    ps, pd_, ops, opd = res[j]
                    for focal, p, op in ((s, ps, ops), (d, pd_, opd)):
                        fr = self.row[focal]
                        new, ne, nx = self._decide(G, E, fr, j, p, op, prow[focal], tiethreshold,
                                                   threshold_pair, err_thresh, emp=emp)
                        corrected_n[fr] += ne
                        excluded_n[fr] += nx
                        if new is not None:
                            G[fr, j] = new
        paired = set(x for p in pairs for x in p)
        singles = [k for k in self.ids if k not in paired]
        snap = G.copy()
        sposts = {}
        for kid in sorted(singles):
            res = self.correct_single(snap, E, kid, back, test_p)
            if res is None:
                continue
            if keep_posteriors:
                sposts[kid] = res
            fr = self.row[kid]
            for j in sorted(res.keys()):
                p, op = res[j]
                new, ne, nx = self._decide(G, E, fr, j, p, op if G[fr, j] != 9 else 0., prow[kid],
                                           tiethreshold, threshold_single, err_thresh, emp=emp, single=True)
                corrected_n[fr] += ne
                if new is not None:
                    G[fr, j] = new
        return dict(corrected=G, raw=raw, E=E, test_p=test_p, corrected_per_animal=corrected_n,
                    excluded_per_animal=excluded_n, pair_posteriors=posts, single_posteriors=sposts)


def _par_rank(args):
    rows, ids, block, back = args
    return Oracle(rows, ids).mendel_rank(block, back)


def _par_correct(args):
    rows, ids, block, E, test_p, kw = args
    r = Oracle(rows, ids).correct_matrix(block, transformed=(E, test_p), **kw)
    return r["corrected"], r["corrected_per_animal"], r["excluded_per_animal"]


def correct_matrix_parallel(rows, ids, G, threshold_pair, threshold_single, back=2, init_filter_p=0.9, tiethreshold=0.05,
                            err_thresh=1, procs=None):
    """Oracle.correct_matrix of ONE matrix on several processes (same result, checked in tests/test_oracle_golden.py):
    the Mendel rank of a column and its correction given the threshold are independent of the other columns; only the
    normalisation by the maximum and the quantile (:600-606,666-667) look at the whole matrix, so those run once in
    between.  For the parity tests at BASELINE sizes."""
    import multiprocessing as mp
    import os
    procs = procs or os.cpu_count() or 1
    G = np.asarray(G, dtype=np.uint8)
    s = G.shape[1]
    procs = max(1, min(procs, s))
    cuts = [int(round(i * s / procs)) for i in range(procs + 1)]
    blocks = [(cuts[i], cuts[i + 1]) for i in range(procs) if cuts[i + 1] > cuts[i]]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(blocks)) as pool:
        raws = pool.map(_par_rank, [(rows, ids, np.ascontiguousarray(G[:, a:b]), back) for a, b in blocks])
        raw = np.concatenate(raws, axis=1)
        E, test_p = Oracle.transform(raw, G, init_filter_p)
        kw = dict(threshold_pair=threshold_pair, threshold_single=threshold_single, back=back, init_filter_p=init_filter_p,
                  tiethreshold=tiethreshold, err_thresh=err_thresh)
        res = pool.map(_par_correct, [(rows, ids, np.ascontiguousarray(G[:, a:b]), np.ascontiguousarray(E[:, a:b]), test_p, kw)
                                      for a, b in blocks])
    corrected = np.concatenate([r[0] for r in res], axis=1)
    return dict(corrected=corrected, raw=raw, E=E, test_p=test_p, corrected_per_animal=sum(r[1] for r in res),
                excluded_per_animal=sum(r[2] for r in res))


# --------------------------------------------------------------------------------------------
# empirical/jalleledist3.py:56-75 (windows), :120-141 (countJointFrq)
# --------------------------------------------------------------------------------------------
def ld_window(n_snps, pos, surround):
    """jalleledist3.py:62-75 incl. the end-of-chromosome quirk (D11): returns (start, stop)."""
    start = max(0, pos - surround)
    stop = pos + surround + 1
    if stop >= n_snps:
        stop = n_snps - 1
    return start, stop


def count_joint_frq(table, mask, w, pseudocount, sum_inner_count=True):
    """jalleledist3.py:120-141 on a uint8 [N,w] window: dict state tuple -> float."""
    subset = table[np.all(mask, axis=1), :]
    out = {}
    for values in itertools.product([0, 1, 2], repeat=w):
        hit = np.ones(subset.shape[0], dtype=bool)
        for c, v in enumerate(values):
            hit &= subset[:, c] == v
        val = np.count_nonzero(hit) + pseudocount
        if sum_inner_count:
            for i in range(int((w - 2) / 2)):
                cols = list(range(w))[i:-i] if i > 0 else []  # conditions[0:-0] is empty (D12)
                if cols:
                    hit = np.ones(subset.shape[0], dtype=bool)
                    for c in cols:
                        hit &= subset[:, c] == values[c]
                    obs = np.count_nonzero(hit)
                else:
                    obs = 1  # np.count_nonzero(np.logical_and.reduce([])) == 1
                obs = obs - val
                val = (obs / (i + 1)) + val
        out[values] = val
    return out


# --------------------------------------------------------------------------------------------
# the empirical (LD-window) blend: jalleledist3.py:101-117 (getCountTable), correct_genotypes3.py:69-82 (getEmpProbs),
# :619-667 (ranks), :746-779 / :883-908 (decisions)
# --------------------------------------------------------------------------------------------
def emp_index(G, surround, pseudocount=1e-4, weight=3, mask=None):
    """The index JointAllellicDistribution(...).countJointFrqAll(table) holds, as arrays: per SNP of the matrix its window
    [wstart, wstart + wlen) in column order (D11 at the chromosome end) and the 3^wlen stored frequencies in
    itertools.product order (first window SNP = most significant digit)."""
    G = np.asarray(G)
    S = G.shape[1]
    mask = np.ones(G.shape, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
    wstart, wlen, freq = np.zeros(S, dtype=np.int64), np.zeros(S, dtype=np.int64), []
    for i in range(S):
        a, b = ld_window(S, i, surround)
        w = max(b - a, 0)
        wstart[i], wlen[i] = a, w
        if w == 0:
            freq.append(np.zeros(0))
            continue
        res = count_joint_frq(G[:, a:b], mask[:, a:b], w, pseudocount)
        freq.append(np.array([res[v] for v in itertools.product([0, 1, 2], repeat=w)]))
    return dict(wstart=wstart, wlen=wlen, freq=freq, weight=weight)


def emp_count_table(row, j, emp):
    """getCountTable: stored frequency of each state of SNP j given the animal's calls in j's window.  A missing call (9) in
    the window is summed over its completions: what jalleledist3.py:84-96 sets out to do before it raises, so
    this case has no reference behaviour; without missing calls it is the reference's lookup.  When the window does not
    hold SNP j itself (D11: the last SNP of the chromosome) the three keys are the same one."""
    a, w = int(emp["wstart"][j]), int(emp["wlen"][j])
    tab = emp["freq"][j].reshape((3,) * w) if w > 0 else emp["freq"][j]
    out = []
    for state in (0, 1, 2):
        idx = []
        for c in range(a, a + w):
            if c == j:
                idx.append(state)
            else:
                v = int(row[c])
                idx.append(v if v in (0, 1, 2) else slice(None))
        sub = tab[tuple(idx)]
        if np.ndim(sub) == 0:
            out.append(sub)                      # the reference's single key (np.sum of one value)
        else:
            acc = None                           # completions in C order (first window SNP slowest), left to right
            for v in np.asarray(sub).reshape(-1):
                acc = v if acc is None else acc + v
            out.append(acc)
    return out


def emp_state_probs(row, j, emp):
    """:752-755 / :888-889: the count table normalised to sum 1 (the pseudocount keeps the sum positive)."""
    c = emp_count_table(row, j, emp)
    return np.divide(c, np.nansum(c))


def emp_blend_ranks(E, G, emp, init_filter_p):
    """correct_genotypes3.py:619-667: every observed cell's rank becomes the mean of itself and weight * (max - own) of its
    normalised window frequencies; then the matrix is divided by its maximum, missing cells are set to 1 and the threshold is
    the init_filter_p quantile (linear interpolation this time, :667)."""
    E = np.array(E, dtype=np.float64)
    G = np.asarray(G)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for j in range(G.shape[1]):
            a, w = int(emp["wstart"][j]), int(emp["wlen"][j])
            if not (a <= j < a + w):
                continue      # :629-631 "missing SNP_id": a window that does not hold its own SNP (D11) is skipped here
            for r in range(G.shape[0]):
                obs = G[r, j]
                if obs in (0, 1, 2):
                    c = emp_count_table(G[r], j, emp)
                    if np.nansum(c) > 0:
                        norm = np.divide(c, np.nansum(c))
                        empdiff = np.nanmax(norm) - norm[obs]
                        E[r, j] = np.mean([E[r, j], empdiff * emp["weight"]])
        E = np.divide(E, np.nanmax(E))
        E[G == 9] = 1
        q = np.quantile(E.flatten(), [0.95, 0.99, init_filter_p])
    return E, q[2]


# --------------------------------------------------------------------------------------------
# pedigree/error_checker.py:138-153,193-204 ; gfutils/extract_founders.py:106-110
# --------------------------------------------------------------------------------------------
def trio_scan(ped, ids, G):
    """Integer Mendelian-inconsistency counts per genotyped animal with both parents genotyped:
    number of SNPs where P(kid | sire, dam) == 0 under the 3x3x3 table (model.py:28-30), and the
    number of SNPs tested (all three calls present)."""
    row = {int(k): i for i, k in enumerate(ids)}
    inc = np.zeros(len(ids), dtype=np.int64)
    tested = np.zeros(len(ids), dtype=np.int64)
    for k in ids:
        s, d = ped.parents(int(k))
        if s in row and d in row:
            a, b, c = G[row[int(k)]], G[row[s]], G[row[d]]
            ok = (a <= 2) & (b <= 2) & (c <= 2)
            t = _T3[np.minimum(a, 2), np.minimum(b, 2), np.minimum(c, 2)]
            tested[row[int(k)]] = ok.sum()
            inc[row[int(k)]] = (ok & (t == 0)).sum()
    return inc, tested


def founders(ped, restrict_ids=None):
    """extract_founders.py:106-110: animals lacking a sire or a dam, after optional restriction
    of the pedigree to `restrict_ids` (get_subset(ids, balance_parents=False))."""
    if restrict_ids is None:
        animals = ped.animals
        has_s = set(ped.sire)
        has_d = set(ped.dam)
    else:
        rs = set(int(x) for x in restrict_ids)
        animals = set(x for x in ped.animals if x in rs)
        has_s = set(k for k, s in ped.sire.items() if k in rs and s in rs)
        has_d = set(k for k, d in ped.dam.items() if k in rs and d in rs)
    return sorted(x for x in animals if x not in has_s or x not in has_d)


# --------------------------------------------------------------------------------------------
# pedigree/error_checker.py:138-204 and empirical/preindex.py:166-170,259-266: per-animal Mendel tally
# --------------------------------------------------------------------------------------------
def trio_candidates(ped, all_ids, kids):
    """error_checker.py:138-153: `animalswithparents` = kids of the list (in genotype-file order) whose sire AND dam
    are genotyped; `candidates` = those kids and their parents (the reference builds a set; here: sorted ids)."""
    genotyped = set(int(x) for x in all_ids)
    wanted = set(int(x) for x in kids)
    with_parents = []
    for k in (int(x) for x in all_ids):
        if k in wanted:
            s, d = ped.parents(k)
            if s in genotyped and d in genotyped:
                with_parents.append(k)
    cand = set()
    for k in with_parents:
        cand.add(k)
        cand.update(ped.parents(k))
    return with_parents, sorted(cand)


def half_transform(raw16, G, missing_value):
    """error_checker.py:193-197 / preindex.py:259-263 on a FLOAT16 matrix: every step of
    log(log(x + 1) + 1) / nanmax(.) is a float16 ufunc (computed in float32, rounded to float16)."""
    raw16 = np.asarray(raw16, dtype=np.float16)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        E = np.log(np.log(raw16 + 1) + 1)
        assert E.dtype == np.float16
        E = E / np.nanmax(E)
        E[np.asarray(G) == 9] = missing_value
    return E


def half_rowsum(E16):
    """DataFrame.sum(axis=1) of a float16 frame (error_checker.py:201, preindex.py:265): pandas reduces the
    [SNP, animal] block along its first axis with dtype float16, i.e. one float16 accumulator per animal that
    takes the SNP columns in order (checked against the reference's own output in tests/golden/errchk_*.npz)."""
    E16 = np.asarray(E16, dtype=np.float16)
    acc = np.zeros(E16.shape[0], dtype=np.float16)
    for j in range(E16.shape[1]):
        acc = acc + E16[:, j]
    return acc


def error_checker(ped_rows, all_ids, G, kids, missing_value=-1.0, back=2):
    """error_checker.py main (:138-204) without the file handling.  Returns (animalswithparents, candidates,
    probs_errors float16 [len(animalswithparents), S], individualSumProbs float16).  The Mendel ranks see ONLY the
    candidate animals as genotyped (the reference restricts the matrix to them, :155-157)."""
    ped = ped_rows if isinstance(ped_rows, OraclePedigree) else OraclePedigree(ped_rows)
    with_parents, cand = trio_candidates(ped, all_ids, kids)
    row_all = {int(k): i for i, k in enumerate(all_ids)}
    Gc = np.asarray(G)[[row_all[k] for k in cand]]
    raw = Oracle(ped, cand).mendel_rank(Gc, back)            # :164-187 (correct_genotypes2's twin of mendelProbsSingle)
    E = half_transform(raw, Gc, missing_value)
    pos = {k: i for i, k in enumerate(cand)}
    Ew = E[[pos[k] for k in with_parents]]
    sums = half_rowsum(Ew)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sums = sums / np.nanmax(sums) if len(sums) else sums  # :202
    return with_parents, cand, Ew, sums


# --------------------------------------------------------------------------------------------
# empirical/preindex.py:166-346: Mendel-filtered LD index per kid-seed chunk and chromosome
# --------------------------------------------------------------------------------------------
PREINDEX_SKIP = ("", "0", "-999", "-9")   # preindex.py:198


def preindex_chunks(ped, all_ids, seedskidschunk):
    """preindex.py:166-189: kids with both parents genotyped, in genotype-file order, cut into seed chunks; per chunk
    the kids plus their genotyped parents (`candidate_kids`)."""
    genotyped = set(int(x) for x in all_ids)
    with_parents = [k for k in (int(x) for x in all_ids) if all(p in genotyped for p in ped.parents(k))]
    n = max(1, seedskidschunk)
    chunks = [with_parents[i:i + n] for i in range(0, len(with_parents), n)]   # chunk(l, n, 0, 0), :160-163
    out = []
    for kid_ids in (c for c in chunks if len(c) > 0):
        cand = set(kid_ids)
        for k in kid_ids:
            cand.update(ped.parents(k))
        out.append((kid_ids, sorted(cand)))
    return out


def preindex_chunk(ped_rows, all_ids, G, cand, chrom_cols, surround, init_filter_p, pseudocount=1e-4, back=2):
    """One seed chunk of preindex.py (:191-346).  chrom_cols: {chromosome name: column indices in map order}.
    The animals evaluated are the candidates whose sire and dam are candidates too (:216-223) and the Mendel ranks see
    ONLY those animals as genotyped (:224).  The rank quantiles are taken on the first chromosome of the chunk and
    reused for the others (:301-302).  Returns {chromosome: dict(eval_ids, quantQ, n_unmasked, sums, freq)} with
    freq[i] = float64 [3^w_i] of SNP i's window in state-product order (empty windows: zero-length)."""
    ped = ped_rows if isinstance(ped_rows, OraclePedigree) else OraclePedigree(ped_rows)
    row_all = {int(k): i for i, k in enumerate(all_ids)}
    cset = set(cand)
    eval_ids = sorted(k for k in cand if all(p in cset for p in ped.parents(k)))
    G = np.asarray(G)
    quants = None
    out = {}
    for chrom in sorted(chrom_cols):
        cols = list(chrom_cols[chrom])
        if chrom in PREINDEX_SKIP or len(cols) < 2:
            continue
        Ge = G[[row_all[k] for k in eval_ids]][:, cols]
        raw = Oracle(ped, eval_ids).mendel_rank(Ge, back)                        # :230-253
        E = half_transform(raw, Ge, 1.0)                                         # :259-263
        sums = half_rowsum(E)                                                    # :265-266
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sums = sums / np.nanmax(sums)
            if quants is None:
                quants = np.nanquantile(E.flatten(), [0.95, 0.99, init_filter_p], method="interpolated_inverted_cdf")
        quantQ = quants[2]
        mask = np.array(E <= quantQ, dtype=bool)                                 # :331
        S = len(cols)
        freq = []
        for i in range(S):                                                       # jalleledist3.py:143-174
            a, b = ld_window(S, i, surround)
            w = max(b - a, 0)
            if w == 0:
                freq.append(np.zeros(0))
                continue
            res = count_joint_frq(Ge[:, a:b], mask[:, a:b], w, pseudocount)
            freq.append(np.array([res[v] for v in itertools.product([0, 1, 2], repeat=w)]))
        out[chrom] = dict(eval_ids=eval_ids, quantQ=quantQ, quants=quants, n_unmasked=int(np.count_nonzero(mask)),
                          sums=sums, freq=freq)
    return out
