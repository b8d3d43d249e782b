"""CPU-side checks of the product's host logic: the C ABI loads and exports what include/*.h declares,
PedigreeDAG flattening, the host copies of the device math (inference engine, children term), the
histogram->threshold transform, and the schedule (pairs / singles / generations) against the oracle.
No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from genofix_b200 import _lib
from genofix_b200.engine import Schedule, rank_transform_table, weighted_quantile_linear
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
from oracle import genofix_oracle as go


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "genofix_b200.h")).read()
    declared = set(re.findall(r"\b(gfx_[a-z0-9_]+)\s*\(", hdr))
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.EXPORTS)
    assert L.gfx_version() >= 100


def test_pedigree_structure_example_fam():
    g = load_golden("kats")
    ped = PedigreeDAG.from_table(load_golden("example_fam")["ped_rows"])
    assert len(ped) == int(g["n_animals"])
    assert len(ped.males) == int(g["n_males"]) and len(ped.females) == int(g["n_females"])
    assert sorted((k, len(v)) for k, v in ped.generationIndex.items()) == [tuple(x) for x in g["gen_sizes"]]
    for rec in g["blanket_struct"]:
        a, depth, gen, nanc, nn = (int(x) for x in rec[:5])
        anc = list(ped.get_parents_depth([a], depth))
        assert len(anc) == nanc
        assert ped.generation[ped._idx(a)] == gen
        nodes = sorted(ped.balance_nodes(list(set([a] + anc)))) if anc else []
        assert nodes == sorted(int(x) for x in rec[5:5 + nn])


def test_pedigree_surface_matches_oracle_pedigree():
    rows = load_golden("toy_b")["ped_rows"]
    ped = PedigreeDAG.from_table(rows)
    op = go.OraclePedigree(rows)
    assert ped.males == op.males and ped.females == op.females
    assert ped.kid2sire == op.sire and ped.kid2dam == op.dam
    for a in op.animals:
        assert ped.get_parents(a) == op.parents(a)
        assert list(ped.get_kids(a)) == op.children(a)
        assert ped.generation[ped._idx(a)] == op.generation[a]
    off, kid = ped.children_csr_reference_order()
    for a in op.animals:
        i = ped._idx(a)
        mine = [int(ped.ids[k]) for k in kid[off[i]:off[i + 1]]]
        assert mine == list(set(op.children(a)))
    sub = ped.get_subset([int(rows[-1, 0])] + list(ped.get_parents_depth([int(rows[-1, 0])], 2)))
    assert set(sub.nodes) <= set(ped.nodes)
    assert len(ped.as_ped()) == int(((rows[:, 1] > 0) & (rows[:, 2] > 0)).sum())


def test_single_known_parent_raises_keyerror():
    rows = np.array([[1, 0, 0, 1], [2, 0, 0, 2], [3, 1, 0, 1], [4, 1, 2, 2]])
    ped = PedigreeDAG.from_table(rows)
    with pytest.raises(KeyError):
        Schedule(ped, [1, 2, 3, 4], 2, host_only=True)


def _topo(nodes, par, focal):
    order, done = [], set()
    for root in list(focal) + list(nodes):
        st = [root]
        while st:
            x = st[-1]
            if x in done:
                st.pop()
                continue
            pend = [p for p in par.get(x, ()) if p not in done]
            if pend:
                st.extend(pend)
            else:
                done.add(x)
                order.append(x)
                st.pop()
    return order


@pytest.mark.parametrize("fixture", ["deep", "toy_b"])
def test_inference_engine_matches_oracle_ve(fixture):
    """gfx_infer_host / gfx_infer_pair_host (the routine the kernels run) vs the oracle's variable
    elimination on random evidence, loopy blankets included: bit-identical posteriors."""
    L = _lib.lib()
    ped = go.OraclePedigree(load_golden(fixture)["ped_rows"])
    rs = np.random.default_rng(5)
    animals = sorted(ped.animals)
    n_checked = n_nan = 0
    for trial in range(1500):
        a = int(rs.choice(animals))
        if trial % 2 == 0:
            focal, depth = [a], int(rs.integers(1, 4))
        else:
            s, d = ped.parents(a)
            if s is None:
                continue
            focal, depth = [s, d], int(rs.integers(2, 4))
        if not ped.ancestors(focal, depth):
            continue
        nodes, par = ped.blanket(focal, depth)
        if any(f not in par for f in focal):
            continue
        order = _topo(nodes, par, focal)
        pos = {v: i for i, v in enumerate(order)}
        n = len(order)
        p1 = np.full(n, -1, np.int8)
        p2 = np.full(n, -1, np.int8)
        for k, (s_, d_) in par.items():
            p1[pos[k]], p2[pos[k]] = pos[s_], pos[d_]
        pobs = rs.choice([0.3, 0.7, 0.95])
        ev = {x: int(rs.integers(0, 3)) for x in order if x not in focal and rs.random() < pobs}
        code = np.full(n, 3, np.uint8)
        for x, v in ev.items():
            code[pos[x]] = v
        ref = go.blanket_marginals(nodes, par, focal, ev)
        out = np.zeros(6)
        vp = lambda arr: arr.ctypes.data_as(C.c_void_p)  # noqa: E731
        if len(focal) == 1:
            rc = L.gfx_infer_host(n, vp(p1), vp(p2), vp(code), pos[focal[0]], vp(out))
        else:
            rc = L.gfx_infer_pair_host(n, vp(p1), vp(p2), vp(code), pos[focal[0]], pos[focal[1]], vp(out))
        if rc == _lib.GFX_E_TOO_WIDE:
            continue
        assert rc == 0
        for i, q in enumerate(focal):
            assert np.array_equal(out[3 * i:3 * i + 3], ref[q], equal_nan=True), (focal, ev)
            n_nan += int(np.isnan(ref[q]).any())
        n_checked += 1
    assert n_checked > 500 and n_nan > 0


def test_oracle_ve_matches_bruteforce():
    ped = go.OraclePedigree(load_golden("deep")["ped_rows"])
    rs = np.random.default_rng(9)
    animals = sorted(ped.animals)
    done = 0
    for _ in range(400):
        a = int(rs.choice(animals))
        if not ped.ancestors([a], 2):
            continue
        nodes, par = ped.blanket([a], 2)
        ev = {x: int(rs.integers(0, 3)) for x in nodes if x != a and rs.random() < 0.6}
        r1 = go.blanket_marginals(nodes, par, [a], ev)[a]
        r2 = go.blanket_marginals_bruteforce(nodes, par, [a], ev)[a]
        assert np.allclose(r1, r2, rtol=1e-12, atol=0, equal_nan=True)
        done += 1
    assert done > 100


def test_children_term_matches_numpy_for_long_litters():
    """gfx_kids_term_host vs the oracle's numpy code for up to 400 children (exercises numpy's
    pairwise-summation blocks of 128 and the recursive halving)."""
    L = _lib.lib()
    rs = np.random.default_rng(3)
    for n in [1, 2, 5, 7, 8, 9, 15, 16, 17, 40, 42, 43, 44, 64, 85, 86, 100, 127, 128, 129, 200, 257, 400]:
        for _ in range(6):
            ks = rs.choice([0, 1, 2, 9], size=n, p=[.3, .3, .3, .1]).astype(np.uint8)
            ps = rs.choice([0, 1, 2, 9], size=n).astype(np.uint8)
            ref = np.asarray(go.probs_kids(list(ks), [None if p == 9 else int(p) for p in ps]), dtype=float)
            out = np.zeros(3)
            k3 = np.where(ks > 2, 3, ks).astype(np.uint8)
            p3 = np.where(ps > 2, 3, ps).astype(np.uint8)
            rc = L.gfx_kids_term_host(n, k3.ctypes.data_as(C.c_void_p), p3.ctypes.data_as(C.c_void_p),
                                      out.ctypes.data_as(C.c_void_p))
            assert rc == 0
            assert np.array_equal(out, ref), (n, out, ref)


def test_weighted_quantile_is_numpy_quantile():
    rs = np.random.default_rng(0)
    for _ in range(300):
        k = int(rs.integers(1, 12))
        vals = rs.choice(np.round(rs.random(40), 3), size=k, replace=False)
        cnt = rs.integers(1, 50, size=k)
        x = np.repeat(vals, cnt)
        for q in (0.0, 0.5, 0.9, 0.95, 0.99, 1.0, float(rs.random())):
            assert weighted_quantile_linear(vals, cnt, q) == np.quantile(x, q)


@pytest.mark.parametrize("name", ["toy_a", "toy_b", "deep", "example_fam", "nan_loop"])
def test_transform_table_matches_reference_threshold(name):
    """Histogram of the golden rank matrix -> E table and quantile == what the reference printed."""
    g = load_golden(name)
    raw = g["raw"].view(np.uint16)
    obs = g["obs"]
    hist = np.zeros(65537, dtype=np.uint64)
    np.add.at(hist, raw[obs != 9].astype(np.int64), 1)
    hist[65536] = int((obs == 9).sum())
    q = g["params"][3]
    elut, test_p, n_nan = rank_transform_table(hist, q)
    assert test_p == g["test_p"] or (np.isnan(test_p) and np.isnan(g["test_p"]))
    E, tp = go.Oracle.transform(g["raw"], obs, q)
    mine = elut[raw]
    mine[obs == 9] = 1.0
    assert np.array_equal(mine, E, equal_nan=True)
    assert n_nan == int(np.isnan(g["raw"].astype(float))[obs != 9].sum())


@pytest.mark.parametrize("name", ["toy_a", "toy_b", "deep", "example_fam"])
def test_schedule_matches_oracle_bookkeeping(name):
    """pairs / per-generation jobs / singles of the native schedule vs correct_genotypes3.py:556,684-686,851-853."""
    g = load_golden(name)
    ids = [int(x) for x in g["ids"]]
    ped = PedigreeDAG.from_table(g["ped_rows"])
    s = Schedule(ped, ids, 2, host_only=True)
    op = go.OraclePedigree(g["ped_rows"])
    rowset = set(ids)
    pairs = set()
    for k in ids:
        sd = op.parents(k)
        if sd[0] is not None and sd[1] is not None:
            pairs.add(sd)
    sched_pairs = sorted(p for p in pairs if p[0] in rowset and p[1] in rowset)
    mine = [(ids[a], ids[b]) for a, b in s.array(0).reshape(-1, 2)]
    assert mine == sched_pairs
    gen_off, jobs = s.array(1), s.array(3)
    for gen in range(s.info["n_generations"]):
        members = op.generation_index.get(gen, set())
        want = [p for p in sched_pairs if p[0] in members or p[1] in members]
        got = [mine[j] for j in jobs[gen_off[gen]:gen_off[gen + 1]]]
        assert got == want
    paired = set(x for p in pairs for x in p)
    # singles: unpaired animals, minus the "lone" ones (no ancestors, no genotyped children) that
    # correct_genotypes3.py:417-419 skips wholesale
    def lone(k):
        return op.parents(k) == (None, None) and not any(c in rowset for c in op.children(k))
    assert sorted(ids[r] for r in s.array(2)) == sorted(k for k in ids if k not in paired and not lone(k))
    assert s.info["n_lone"] == sum(1 for k in ids if k not in paired and lone(k))
    assert s.info["n_trios"] == sum(1 for k in ids if op.parents(k)[0] in rowset and op.parents(k)[1] in rowset)
    assert sorted(s.founders()) == go.founders(op)
    some = ids[::3]
    assert sorted(s.founders(some)) == go.founders(op, some)


def test_chunking_known_answers():
    """genofix.py:207-210 chunk(): SURVEY appendix B.2."""
    from genofix_b200.chunking import chunk
    def spans(L, n, s):
        b = 2 * s + 1
        return [(c[0], c[-1], len(c)) for c in chunk(list(range(L)), n, b, b)]
    assert spans(100, 1000, 4) == [(0, 99, 100)]
    assert spans(2500, 1000, 4) == [(0, 1008, 1009), (1000, 2008, 1009), (2000, 2499, 500)]
    assert spans(25, 10, 1) == [(0, 12, 13), (10, 24, 15), (20, 24, 5)]
    assert spans(23, 10, 3) == [(0, 22, 23), (10, 22, 13)]


def test_packed_container_roundtrip(tmp_path):
    """GFX2BIT1 container: pack/unpack, GensCache surface, SNP-block reads at word boundaries and ragged ends, write-back."""
    from genofix_b200.gfutils.packedgens import PackedGens, pack_codes, unpack_codes, words_per_row
    rs = np.random.default_rng(0)
    G = rs.choice([0, 1, 2, 9], size=(37, 1003), p=[.3, .4, .28, .02]).astype(np.uint8)
    P = pack_codes(G)
    assert P.shape == (37, words_per_row(1003)) and P.dtype == np.uint32
    assert np.array_equal(unpack_codes(P, 1003), G)
    assert np.all(unpack_codes(P, P.shape[1] * 16)[:, 1003:] == 9)          # padding cells are "missing"
    path = str(tmp_path / "x.g2b")
    PackedGens.write(path, np.arange(100, 137), ["s%d" % i for i in range(1003)], G)
    g = PackedGens(path)
    assert g.all_ids == list(range(100, 137)) and g.snps[5] == "s5" and g.n_snps == 1003
    assert np.array_equal(g.getMatrix([105, 100, 999]), G[[5, 0]])           # unknown ids are skipped like the reference
    for a, b in [(0, 1003), (16, 516), (992, 1003), (160, 176)]:
        blk = g.packed_columns(a, b)
        assert np.array_equal(blk, pack_codes(G[:, a:b])), (a, b)
    for a, b in [(5, 100), (33, 1003), (1, 2), (15, 17), (999, 1003)]:      # starts inside a word: shifted across word boundaries
        blk = g.packed_columns(a, b)
        assert np.array_equal(blk, pack_codes(G[:, a:b])), (a, b)
    with pytest.raises(ValueError):
        g.packed_columns(5, 2000)
    g2 = PackedGens(path, mode="r+")
    g2.store_columns(16, 516, pack_codes((G[:, 16:516] + 1) % 3))
    g2.packed.flush()
    M = PackedGens(path).getMatrix(range(100, 137))
    assert np.array_equal(M[:, 16:516], (G[:, 16:516] + 1) % 3) and np.array_equal(M[:, :16], G[:, :16]) and np.array_equal(M[:, 516:], G[:, 516:])
    g2.store_columns(603, 777, pack_codes((G[:, 603:777] + 2) % 3))             # unaligned write-back
    g2.packed.flush()
    M2 = PackedGens(path).getMatrix(range(100, 137))
    assert np.array_equal(M2[:, 603:777], (G[:, 603:777] + 2) % 3) and np.array_equal(M2[:, 516:603], G[:, 516:603]) and np.array_equal(M2[:, 777:], G[:, 777:])
    # text -> container
    txt = tmp_path / "g.ssv"
    with open(txt, "w") as f:
        f.write("id " + " ".join("s%d" % i for i in range(20)) + "\n")
        for i in range(5):
            f.write("%d %s\n" % (i + 1, " ".join(str(x) for x in G[i, :20])))
    g3 = PackedGens.from_text(str(txt), str(tmp_path / "g.g2b"))
    assert np.array_equal(g3.getMatrix([1, 2, 3, 4, 5]), G[:5, :20]) and g3.snps == ["s%d" % i for i in range(20)]


def test_levels_from_parents_orders_parents_first():
    from genofix_b200.simulate_device import levels_from_parents
    #        0   1   2   3   4   5
    sire = [-1, -1, 0, 0, 2, -1]
    dam = [-1, -1, 1, 1, 3, 4]
    order, off = levels_from_parents(sire, dam)
    pos = {int(r): i for i, r in enumerate(order)}
    for k in range(6):
        for p in (sire[k], dam[k]):
            if p >= 0:
                assert pos[p] < pos[k]
    assert list(off) == [0, 2, 4, 5, 6]                 # founders {0,1}, then {2,3}, {4}, {5}
    with pytest.raises(ValueError):
        levels_from_parents([1, 0], [-1, -1])           # a cycle


@pytest.mark.parametrize("name", ["errchk_a", "errchk_b", "errchk_long"])
def test_error_checker_host_side_matches_oracle(name):
    """Trio selection (error_checker.py:138-153) and the float16 transform table (:193-197) of the product's host
    side against the oracle and the reference's own output; no GPU involved."""
    from genofix_b200.engine import half_transform_table
    from genofix_b200.pedigree.error_checker import trio_candidates
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    g = load_golden(name)
    ped = PedigreeDAG.from_table(g["ped_rows"])
    wp, cand = trio_candidates(ped, g["ids"], g["kids"])
    owp, ocand = go.trio_candidates(go.OraclePedigree(g["ped_rows"]), g["ids"], g["kids"])
    assert wp == owp and cand == ocand
    assert np.array_equal(np.array(wp), g["with_parents"])
    # table from the histogram of the oracle's scores == transform of the matrix itself
    row = {int(k): i for i, k in enumerate(g["ids"])}
    Gc = g["obs"][[row[k] for k in cand]]
    raw = go.Oracle(g["ped_rows"], cand).mendel_rank(Gc, 2)
    pat = raw.view(np.uint16).astype(np.int64)
    hist = np.zeros(65537, dtype=np.uint64)
    np.add.at(hist, np.where(Gc == 9, 65536, pat).ravel(), 1)
    table, _m = half_transform_table(hist, -1.0)
    E = table[np.where(Gc == 9, 0xFFFF, pat)]
    assert np.array_equal(E.view(np.uint16), go.half_transform(raw, Gc, -1.0).view(np.uint16))
    pos = {k: i for i, k in enumerate(cand)}
    assert np.array_equal(E[[pos[k] for k in wp]].view(np.uint16), g["probs_errors"])


def test_weighted_nanquantile_is_numpy_nanquantile_on_float16():
    """preindex.py:270,301: np.nanquantile(float16 data, q, method='interpolated_inverted_cdf') from a histogram of the
    values, bit for bit (numpy subtracts the neighbours in float16 and interpolates in float64)."""
    import warnings
    from genofix_b200.engine import weighted_nanquantile
    rs = np.random.default_rng(0)
    for trial in range(200):
        n = int(rs.integers(1, 400))
        x = rs.choice(rs.random(int(rs.integers(1, 12))).astype(np.float16), size=n)
        if trial % 5 == 0:
            x[rs.integers(0, n)] = np.nan
        qs = [0.95, 0.99, float(rs.random())]
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = np.nanquantile(x, qs, method="interpolated_inverted_cdf")
        vals, cnt = np.unique(x.view(np.uint16), return_counts=True)
        got = weighted_nanquantile(vals.view(np.float16), cnt, qs, 0.0, 1.0)
        assert np.array_equal(ref, got, equal_nan=True)


@pytest.mark.parametrize("name", ["preidx_a", "preidx_b"])
def test_preindex_seed_chunks_match_oracle(name):
    from genofix_b200.empirical.preindex import seed_chunks
    g = load_golden(name)
    ped = PedigreeDAG.from_table(g["ped_rows"])
    for size in (10000, 7, 1):
        assert seed_chunks(ped, g["ids"], size) == go.preindex_chunks(go.OraclePedigree(g["ped_rows"]), g["ids"], size)


# ---- round 2: surfaces that had no pinned answers ----------------------------------------
def test_product_model_functions_match_reference_kats():
    """genofix_b200.bayesnet.model (the product functions, not the oracle's copies) against the answers of the
    reference's bayesnet/model.py:65-133 stored in kats.npz."""
    from genofix_b200.bayesnet import model
    g = load_golden("kats")
    for c in range(len(g["kid"])):
        n = int((g["kid"][c] >= 0).sum())
        ks = [int(k) for k in g["kid"][c, :n]]
        ps = [None if p < 0 else int(p) for p in g["par"][c, :n]]
        assert np.array_equal(np.asarray(model.generate_probs_kids(ks, ps), dtype=float), g["probs_kids"][c])
        inc = [k in (0, 1, 2) for k in ks]
        ks_i = [k for k, i in zip(ks, inc) if i]
        ps_i = [p for p, i in zip(ps, inc) if i]
        for o in range(3):
            d = model.generate_probs_differences_kids(ks_i, ps_i, o) if ks_i else []
            want = g["diff_kids"][c, o, :len(ks_i)]
            assert len(d) == len(ks_i)
            assert np.array_equal(np.array(d, dtype=np.float16).view(np.uint16), want.view(np.uint16))
    # unusable children: one (zero) entry each, after the usable ones (model.py:100,113-115)
    d = model.generate_probs_differences_kids([9, 0, 9, 2], [None, None, 1, None], 0)
    assert len(d) == 4 and float(d[2]) == 0.0 and float(d[3]) == 0.0 and float(d[1]) > 0
    assert model.generate_probs_differences_kids([9, 9], [None, 1], 1) == []
    assert np.allclose(model.generate_probs_kids([1, 0, 1, 0, 1], [2, 2, 2, 2, 2]), [2 / 3, 1 / 3, 0])      # model.py:91
    for args, want in (((2, 2, 1), 1.0), ((1, 1, 0), 0.25), ((2, 9, 0), 0.5), ((9, 2, 0), 0.5), ((9, 9, 1), 0)):   # :136-140
        assert model.generate_probs_differences_parent(*args) == want
    for i, p1 in enumerate([0, 1, 2, 9]):
        for j, p2 in enumerate([0, 1, 2, 9]):
            for o in range(3):
                assert model.generate_probs_differences_parent(p1, p2, o) == g["diff_parent"][i, j, o]


@pytest.mark.parametrize("name,mcs", [("example", 100), ("example", 10), ("syn_a", 100), ("syn_a", 30), ("syn_b", 50)])
def test_split_pedigree_matches_reference(name, mcs):
    """PedigreeDAG.split_pedigree against the clusters the reference's own method returned (pedigree_dag.py:41-122),
    in the reference's list order."""
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    g = load_golden("kats2")
    want = g["split_%s_%d" % (name, mcs)]
    cl = PedigreeDAG.from_table(g["split_rows_" + name]).split_pedigree(mcs)
    mine = np.array([(i, int(a)) for i, c in enumerate(cl) for a in sorted(c)], dtype=np.int64)
    assert np.array_equal(mine, want)


def test_split_pedigree_lone_node_raises_like_the_reference():
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    g = load_golden("kats2")
    msg = str(g["split_lone_error"])
    assert msg.endswith("illegal lone node in pedigree")
    with pytest.raises(Exception) as e:
        PedigreeDAG.from_table(g["split_rows_lone"]).split_pedigree(100)
    assert str(e.value) == msg


@pytest.mark.parametrize("sur", [1, 2, 3])
def test_count_table_and_emp_probs_match_reference(sur):
    """JointAllellicDistribution.getCountTable (jalleledist3.py:101-117) and CorrectGenotypes.getEmpProbs
    (correct_genotypes3.py:69-82) on an index holding the frequencies the reference's countJointFrqAll stored
    (complete evidence: with a missing call the reference raises)."""
    import pandas as pd
    from genofix_b200.correct_genotypes3 import CorrectGenotypes, initializerEmp
    from genofix_b200.empirical.jalleledist3 import JointAllellicDistribution
    g = load_golden("kats2")
    G = g["emp_G"]
    n, s = G.shape
    names = ["SNP%d" % i for i in range(s)]
    emp = JointAllellicDistribution("1", names, surround_size=sur)
    for i, snp in enumerate(names):
        w = int(g["emp_wlen_%d" % sur][i])
        assert len(emp.getWindow(snp)) == w
        emp._freq[snp] = g["emp_freq_%d" % sur][i][:3 ** w].copy()
    ct = g["emp_counttable_%d" % sur]
    for a in range(ct.shape[0]):
        ev = {snp: int(G[a, j]) for j, snp in enumerate(names)}
        for j, snp in enumerate(names):
            assert np.array_equal(np.asarray(emp.getCountTable(ev, snp), dtype=float), ct[a, j]), (a, snp)
    initializerEmp(emp)
    df = pd.DataFrame(G, index=np.arange(1, n + 1), columns=names)
    ep = g["emp_probs_%d" % sur]
    for j, snp in enumerate(names):
        win = emp.getWindow(snp)
        if snp not in win:
            assert np.isnan(ep[j]).all()          # D11: the last SNP's window does not contain it
            continue
        res = CorrectGenotypes.getEmpProbs(df.loc[:, win], snp)
        got = np.full((n, 3), np.nan)
        for (kid, obs), v in res.items():
            assert obs == G[kid - 1, j]
            got[kid - 1] = v
        assert np.array_equal(got, ep[j], equal_nan=True), snp


def test_copy_order_hands_out_upload_events_in_chunk_order():
    """MultiLane's upload order (engine.CopyOrder): chunk k's lane blocks until chunk k - 1's upload has been enqueued and
    receives its event; out-of-range chunks and a failed lane release every waiter with None."""
    import threading
    import time
    from genofix_b200.engine import CopyOrder
    order = CopyOrder(5)
    assert order.wait_for(-1) is None and order.wait_for(5) is None
    got = {}

    def lane(l):
        for k in range(l, 5, 3):
            got[k] = order.wait_for(k - 1)
            time.sleep(0.01 * ((k * 7) % 3))
            order.publish(k, "upload-%d" % k)
    th = [threading.Thread(target=lane, args=(l,)) for l in (2, 1, 0)]
    for x in th:
        x.start()
    for x in th:
        x.join(timeout=10)
        assert not x.is_alive()
    assert got == {0: None, 1: "upload-0", 2: "upload-1", 3: "upload-2", 4: "upload-3"}
    # a lane that dies before publishing must not leave the others waiting
    order = CopyOrder(3)
    res = []
    w = threading.Thread(target=lambda: res.append(order.wait_for(0)))
    w.start()
    time.sleep(0.05)
    assert w.is_alive()
    order.fail()
    w.join(timeout=10)
    assert not w.is_alive() and res == [None]
