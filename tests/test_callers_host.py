"""Host logic of the two callers around the path -- pedigree/error_checker.py (mendel_error_tally) and
empirical/preindex.py (index_chunk) -- on CPU against the outputs of the UNMODIFIED reference (tests/golden/errchk_*.npz,
preidx_*.npz).  The two device calls (Engine.mendel_tally, window_counts_device) are replaced by oracle-backed stand-ins,
so what is checked is the product's Python: trio selection, candidate restriction, the float16 transform table, the
histogram quantile, the mask and the window bookkeeping.  The same fixtures run against the real kernels in
tests/test_gpu_parity.py."""
import numpy as np
import pandas as pd
import pytest

from conftest import load_golden
from genofix_b200.engine import half_transform_table
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
from oracle import genofix_oracle as go


def _stand_in_engine(ped_rows):
    class StandIn(object):
        def __init__(self, ped, ids, back):
            self.ids, self.back = [int(x) for x in ids], back

        def mendel_tally(self, G, rows, missing_value, want_matrix=True):
            raw = go.Oracle(ped_rows, self.ids).mendel_rank(G, self.back)          # stands in for k_mendel_rank<HIST>
            pat = raw.view(np.uint16).astype(np.int64)
            self.tally_hist = np.bincount(np.where(G == 9, 65536, pat).ravel(), minlength=65537).astype(np.uint64)
            self.tally_table, m = half_transform_table(self.tally_hist, missing_value)
            E = self.tally_table[np.where(G == 9, 0xFFFF, pat)]
            E = E if rows is None else E[rows]
            return E, go.half_rowsum(E), m                                       # stands in for k_rank_rowsum_f16
    return StandIn


def _stand_in_window_counts(G, mask, starts, lens):
    """stands in for k_window_counts: per window the full histogram, then the histograms of the inner windows
    (cells l .. w-1-l) over the rows whose whole window passes the mask (include/genofix_b200.h)"""
    from genofix_b200.empirical.jalleledist3 import counters_per_window, inner_levels
    out = np.zeros((len(starts), counters_per_window(max(lens))), np.int64)
    for i, (a, w) in enumerate(zip(starts, lens)):
        passes = mask[:, a:a + w].all(axis=1) if mask is not None else np.ones(len(G), bool)
        off = 0
        for l in range(inner_levels(w) if w > 0 else 0):
            sub = G[:, a + l:a + w - l]
            ok = passes & (sub <= 2).all(axis=1)
            idx = (sub[ok].astype(np.int64) * (3 ** np.arange(w - 2 * l - 1, -1, -1))).sum(axis=1)
            np.add.at(out[i], off + idx, 1)
            off += 3 ** (w - 2 * l)
    return out


def test_window_frequencies_with_missing_outer_cells_match_reference_semantics():
    """a row with a missing call in an OUTER window cell still counts in the inner windows
    (jalleledist3.py:133-138); checked against the oracle restatement of countJointFrq on every state."""
    import itertools
    from genofix_b200.empirical.jalleledist3 import frequencies_from_counts
    rs = np.random.default_rng(11)
    for w in (5, 6, 7, 8, 9):
        n = 300
        G = rs.choice([0, 1, 2, 9], size=(n, w), p=[.3, .3, .3, .1]).astype(np.uint8)
        for mask in (np.ones((n, w), bool), rs.random((n, w)) < 0.97):
            counts = _stand_in_window_counts(G, mask, [0], [w])[0]
            vals = frequencies_from_counts(counts, w, 1e-4, True)
            ref = go.count_joint_frq(G, mask, w, 1e-4, sum_inner_count=True)
            want = np.array([ref[v] for v in itertools.product([0, 1, 2], repeat=w)])
            assert np.array_equal(vals, want), w


@pytest.mark.parametrize("name", ["errchk_a", "errchk_b", "errchk_long", "errchk_example"])
def test_error_checker_host_logic_reproduces_reference_files(name, monkeypatch):
    import genofix_b200.pedigree.error_checker as ec
    g = load_golden(name)
    monkeypatch.setattr(ec, "Engine", _stand_in_engine(g["ped_rows"]))
    df = pd.DataFrame(g["obs"], index=[int(x) for x in g["ids"]], columns=["SNP%d" % i for i in range(g["obs"].shape[1])])
    pe, sums = ec.mendel_error_tally(df, PedigreeDAG.from_table(g["ped_rows"]), g["kids"])
    assert [int(x) for x in pe.index] == [int(x) for x in g["with_parents"]]
    assert np.array_equal(pe.to_numpy().view(np.uint16), g["probs_errors"])
    assert np.array_equal(sums.view(np.uint16), g["sums"])


@pytest.mark.parametrize("name", ["preidx_a", "preidx_b", "preidx_c"])
def test_preindex_host_logic_reproduces_reference_indexes(name, monkeypatch):
    import genofix_b200.empirical.jalleledist3 as jd
    import genofix_b200.empirical.preindex as pp
    g = load_golden(name)
    monkeypatch.setattr(pp, "Engine", _stand_in_engine(g["ped_rows"]))
    monkeypatch.setattr(jd, "window_counts_device", _stand_in_window_counts)
    ped = PedigreeDAG.from_table(g["ped_rows"])
    chunks = pp.seed_chunks(ped, g["ids"], int(g["seed_chunk"]))
    assert len(chunks) == (4 if name == "preidx_c" else 1)
    row = {int(k): i for i, k in enumerate(g["ids"])}
    names = ["SNP%d" % i for i in range(g["obs"].shape[1])]
    c2s = {"0": ["pad"]}
    for n, c in zip(names, g["chrom"]):
        c2s.setdefault(str(c), []).append(n)
    k = 0
    for ci, (_kid_ids, cand) in enumerate(chunks):
        df = pd.DataFrame(g["obs"][[row[x] for x in cand]], index=cand, columns=names)
        res = pp.index_chunk(df, ped, c2s, int(g["surround"]), float(g["q"]))
        assert sorted(res) == [str(c) for c in g["chroms"]]
        for c in sorted(res):
            empC, info = res[c]
            assert info["quantQ"] == g["quantQ"][k]              # printed per (seed chunk, chromosome)
            assert (info["n_unmasked"], len(info["eval_ids"]) * len(c2s[c])) == tuple(g["kept"][k])
            k += 1
            if ci == len(chunks) - 1:                              # the pickled index is the last chunk's (:339-346)
                for i, s_ in enumerate(empC.snp_ordered):
                    w = int(g["wlen_" + c][i])
                    assert len(empC.windows[s_]) == w
                    if w:
                        assert np.array_equal(empC._freq[s_], g["freq_" + c][i][:3 ** w])
    assert k == len(g["quantQ"])


# ---- the LD index handed to the blend kernels (correct_genotypes3.ld_index_arrays) and the lookup surface -----------------
@pytest.mark.parametrize("name", ["emp_a", "emp_b"])
def test_ld_index_arrays_and_lookups_match_the_reference(name, monkeypatch):
    """countJointFrqAll with the oracle-backed stand-in for k_window_counts fills the product's index; its stored frequencies,
    the arrays gfx_emp_rank_blend / gfx_emp_apply receive, getCountTable and getEmpProbs are then the reference's
    (tests/golden/emp_*.npz: the frequencies of the index the reference built and every getEmpProbs result it returned)."""
    import multiprocessing
    from genofix_b200 import correct_genotypes3 as cg3
    from genofix_b200.empirical import jalleledist3 as jad
    g = load_golden(name)
    obs = g["obs"]
    S = obs.shape[1]
    names = ["SNP%d" % i for i in range(S)]
    monkeypatch.setattr(jad, "window_counts_device", _stand_in_window_counts)
    emp = jad.JointAllellicDistribution("1", names, surround_size=int(g["emp_surround"]), pseudocount=float(g["emp_pseudocount"]))
    df = pd.DataFrame(obs, index=[int(x) for x in g["ids"]], columns=names)
    emp.countJointFrqAll(df)
    wstart, wlen, freq, weight = cg3.ld_index_arrays(emp, names, float(g["emp_weight"]))
    assert np.array_equal(wstart, g["emp_wstart"]) and np.array_equal(wlen, g["emp_wlen"]) and weight == float(g["emp_weight"])
    for j in range(S):
        assert np.array_equal(freq[j], g["emp_freq"][j, :3 ** int(wlen[j])]), j
    # getEmpProbs on the window chunk of a SNP, like correctMatrix calls it (:624-633)
    cg3.initializerEmp(emp)
    want = {}
    for r, j, o, a, b, c in g["emp_probs"]:
        want.setdefault(int(j), {})[(int(g["ids"][int(r)]), int(o))] = np.array([a, b, c])
    for j in sorted(want)[::3]:
        win = emp.getWindow(names[j])
        got = cg3.CorrectGenotypes.getEmpProbs(df.loc[:, win].copy(), names[j])
        assert set(got) == set(want[j])
        for k, v in got.items():
            assert np.array_equal(np.asarray(v, dtype=float), want[j][k]), (j, k)
    # a window that reaches outside the columns is refused (the reference would treat the SNP as unobserved and raise, D4)
    with pytest.raises(ValueError):
        cg3.ld_index_arrays(emp, names[1:], 3.0)
    with pytest.raises(KeyError):
        cg3.ld_index_arrays(emp, names + ["nope"], 3.0)


def test_oracle_blend_with_missing_calls_sums_over_completions():
    """The case the reference cannot run (a 9 inside a window raises): the oracle's count table is the sum of the
    stored frequencies over every completion of the missing calls, and without missing calls it is the single stored entry."""
    import itertools
    rs = np.random.default_rng(3)
    G = rs.integers(0, 3, size=(200, 9)).astype(np.uint8)
    emp = go.emp_index(G, 2, 1e-4, weight=3)
    row = G[5].copy()
    j = 4
    a, w = int(emp["wstart"][j]), int(emp["wlen"][j])
    tab = emp["freq"][j].reshape((3,) * w)
    full = go.emp_count_table(row, j, emp)
    for s in range(3):
        idx = [int(row[c]) if c != j else s for c in range(a, a + w)]
        assert full[s] == tab[tuple(idx)]
    row[3] = 9
    row[6] = 9
    part = go.emp_count_table(row, j, emp)
    for s in range(3):
        tot = 0.0
        first = True
        for x3, x6 in itertools.product(range(3), repeat=2):          # column 3 slower than column 6: C order
            idx = [int(row[c]) for c in range(a, a + w)]
            idx[3 - a], idx[6 - a], idx[j - a] = x3, x6, s
            tot = tab[tuple(idx)] if first else tot + tab[tuple(idx)]
            first = False
        assert part[s] == tot
    p = go.emp_state_probs(row, j, emp)
    assert abs(float(np.sum(p)) - 1.0) < 1e-12
