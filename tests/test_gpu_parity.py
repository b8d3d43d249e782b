"""Parity of the CUDA path (through the C ABI) with the golden fixtures of the unmodified reference
and with the oracle on seeded inputs.  Integer / byte results are bit-exact; posteriors are compared
to 1e-5 relative in log space (BASELINE.json north_star) and additionally counted for bit equality."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

LOG_RTOL = 1e-5   # posteriors: |log p_gpu - log p_ref| <= 1e-5 * max(1, |log p_ref|)


def _engine(g, back=2):
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    ped = PedigreeDAG.from_table(g["ped_rows"])
    return Engine(ped, g["ids"], back)


def _same_f16(a, b):
    """bit-exact float16 comparison; NaN payloads are not significant"""
    a, b = np.asarray(a), np.asarray(b)
    an, bn = np.isnan(a.astype(np.float32)), np.isnan(b.astype(np.float32))
    return bool(np.array_equal(an, bn) and np.array_equal(a.view(np.uint16)[~an], b.view(np.uint16)[~bn]))


def _log_close(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    if not np.array_equal(a == 0, b == 0):
        return False
    m = (a > 0) & (b > 0)
    la, lb = np.log(a[m]), np.log(b[m])
    return bool(np.all(np.abs(la - lb) <= LOG_RTOL * np.maximum(1.0, np.abs(lb))))


@pytest.mark.parametrize("name", ["toy_a", "toy_b", "deep", "nan_loop", "example_fam"])
def test_correct_matrix_matches_reference_golden(name):
    g = load_golden(name)
    tp, ts, back, q, tie, et = g["params"]
    eng = _engine(g, int(back))
    out = eng.correct(g["obs"], tp, ts, init_filter_p=q, tiethreshold=tie, err_thresh=et, keep_posteriors=True)
    raw = eng.raw_host()
    assert _same_f16(raw, g["raw"]), "rank matrix not bit-exact"
    tpv = eng.stats["test_p"]
    assert tpv == g["test_p"] or (np.isnan(tpv) and np.isnan(g["test_p"]))
    assert np.array_equal(out, g["corrected"]), "%d corrected calls differ" % int((out != g["corrected"]).sum())
    assert int(eng.stats["corrected_per_animal"].sum()) == int(g["errors_all"])
    assert int(eng.stats["excluded_per_animal"].sum()) == int(g["pair_excluded"])
    if np.isnan(tpv):
        return
    # posteriors of every suspect (pair, SNP) and (single, SNP)
    ids = [int(x) for x in g["ids"]]
    pairs = eng.schedule.array(0).reshape(-1, 2)
    gen_off, jobs = eng.schedule.array(1), eng.schedule.array(3)
    mine, occ = {}, {}
    for gen, (post, res) in enumerate(eng.stats["posteriors"]["pairs"]):
        for k, j in enumerate(jobs[gen_off[gen]:gen_off[gen + 1]]):
            s, d = ids[pairs[j][0]], ids[pairs[j][1]]
            o = occ.get((s, d), 0)
            occ[(s, d)] = o + 1
            for snp in np.nonzero(res[2 * k] | res[2 * k + 1])[0]:
                mine[(o, s, d, int(snp))] = np.concatenate([post[2 * k, snp], post[2 * k + 1, snp]])
    ref = {(int(x[0]), int(x[1]), int(x[2]), int(x[3])): x[4:] for x in g["pair_post"]}
    assert set(mine) == set(ref)
    nbit = 0
    for k, v in ref.items():
        assert _log_close(mine[k], v), (k, mine[k], v)
        nbit += int(np.array_equal(mine[k], v, equal_nan=True))
    post, res = eng.stats["posteriors"]["singles"]
    smine = {}
    for k, r in enumerate(eng.schedule.array(2)):
        for snp in np.nonzero(res[k])[0]:
            smine[(ids[r], int(snp))] = post[k, snp]
    sref = {(int(x[0]), int(x[1])): x[2:] for x in g["single_post"]}
    assert set(smine) == set(sref)
    for k, v in sref.items():
        assert _log_close(smine[k], v), (k, smine[k], v)
    print("%s: %d/%d pair posteriors bit-identical" % (name, nbit, len(ref)))


def _random_case(seed, n=400, gens=7, s=160, related=0.3, unknown=0.05, miss=0.03, drop=0.1):
    from genofix_b200 import synth
    rows = synth.make_pedigree(n, gens, sire_frac=0.2, dam_frac=0.5, seed=seed, related_mating_p=related,
                               unknown_parents_frac=unknown)
    truth = synth.gene_drop(rows, s, seed=seed)
    obs = synth.inject_missing(synth.inject_errors(truth, 0.02, seed=seed), miss, seed=seed + 1)
    ids = rows[:, 0].copy()
    if drop > 0:
        keep = np.random.default_rng(seed).random(len(ids)) >= drop
        ids, obs = ids[keep], obs[keep]
    return rows, ids, obs


@pytest.mark.parametrize("seed,params", [(101, (0.5, 0.64, 2, 0.95, 0.4, 1.0)), (202, (0.5, 0.7, 2, 0.9, 0.05, 1.0)),
                                         (303, (0.6, 0.6, 1, 0.9, 0.2, 1.0)), (404, (0.5, 0.64, 3, 0.92, 0.4, 0.9)),
                                         (505, (0.5, 0.64, 3, 0.9, 0.4, 1.0))])
def test_correct_matrix_matches_oracle_on_seeded_inputs(seed, params):
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle.genofix_oracle import Oracle
    if seed == 505:   # look-back 3 (pair blankets of depth 4) on a larger, less inbred pedigree
        rows, ids, obs = _random_case(seed, n=1500, gens=6, s=48, related=0.0, unknown=0.0, miss=0.0, drop=0.0)
    else:
        rows, ids, obs = _random_case(seed, s=96 if params[2] == 3 else 160)
    tp, ts, back, q, tie, et = params
    o = Oracle(rows, ids).correct_matrix(obs, tp, ts, back=back, init_filter_p=q, tiethreshold=tie, err_thresh=et)
    eng = Engine(PedigreeDAG.from_table(rows), ids, back)
    out = eng.correct(obs, tp, ts, init_filter_p=q, tiethreshold=tie, err_thresh=et)
    assert _same_f16(eng.raw_host(), o["raw"])
    assert eng.stats["test_p"] == o["test_p"] or (np.isnan(eng.stats["test_p"]) and np.isnan(o["test_p"]))
    assert np.array_equal(eng.E_host(), o["E"], equal_nan=True)
    assert np.array_equal(out, o["corrected"]), int((out != o["corrected"]).sum())
    assert np.array_equal(eng.stats["corrected_per_animal"], o["corrected_per_animal"])
    assert np.array_equal(eng.stats["excluded_per_animal"], o["excluded_per_animal"])


def test_edge_shapes_and_ragged_inputs():
    """1 SNP, 15/16/17 SNPs (word boundaries), all-missing rows and columns, a pedigree of founders only."""
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle.genofix_oracle import Oracle
    rows, ids, obs = _random_case(7, n=120, gens=5, s=40, drop=0.0)
    for s in (1, 15, 16, 17, 33):
        sub = obs[:, :s].copy()
        if s >= 15:
            sub[:, 3] = 9
            sub[5, :] = 9
        o = Oracle(rows, ids).correct_matrix(sub, 0.5, 0.64, back=2, init_filter_p=0.9, tiethreshold=0.4)
        eng = Engine(PedigreeDAG.from_table(rows), ids, 2)
        out = eng.correct(sub, 0.5, 0.64, init_filter_p=0.9, tiethreshold=0.4)
        assert _same_f16(eng.raw_host(), o["raw"]), s
        assert np.array_equal(out, o["corrected"]), s
    frows = np.array([[i, 0, 0, 1 + i % 2] for i in range(1, 9)])
    fobs = np.random.default_rng(1).integers(0, 3, size=(8, 20)).astype(np.uint8)
    eng = Engine(PedigreeDAG.from_table(frows), frows[:, 0], 2)
    out = eng.correct(fobs, 0.5, 0.64)
    assert np.array_equal(out, fobs) and np.all(eng.raw_host() == np.float16(1.0))


def test_pack_unpack_roundtrip():
    import torch
    from genofix_b200 import _lib
    L = _lib.lib()
    rs = np.random.default_rng(0)
    for n, s in [(3, 1), (5, 16), (7, 17), (64, 1000), (33, 4097)]:
        G = rs.choice([0, 1, 2, 9], size=(n, s)).astype(np.uint8)
        ld = ((s + 15) // 16) * 16 if s % 2 == 0 else s
        host = np.full((n, ld), 9, np.uint8)
        host[:, :s] = G
        d = torch.from_numpy(host).cuda()
        wpr = (((s + 15) // 16 + 3) // 4) * 4
        packed = torch.empty((n, wpr), dtype=torch.int32, device="cuda")
        out = torch.zeros_like(d)
        _lib.check(L.gfx_pack2(d.data_ptr(), n, s, ld, packed.data_ptr(), wpr, None))
        _lib.check(L.gfx_unpack2(packed.data_ptr(), n, s, wpr, out.data_ptr(), ld, None))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy()[:, :s], G)
        p = packed.cpu().numpy().view(np.uint32)
        j = int(rs.integers(0, s))
        code = (p[:, j // 16] >> (2 * (j % 16))) & 3
        assert np.array_equal(np.where(code == 3, 9, code), G[:, j])


def test_snp_permutation_property_at_scale():
    """Every SNP column is independent given the pedigree and the chunk-level threshold is permutation
    invariant, so correcting a column-permuted matrix must give the column-permuted result.  Run at
    10k animals x 4096 SNPs (BASELINE config 2 pedigree), where the oracle is too slow to be the check."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    rows = synth.make_pedigree(10000, 6, seed=1234)
    obs = synth.inject_errors(synth.gene_drop(rows, 4096, seed=1234), 0.01, seed=1234)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    a = eng.correct(obs, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    tally_a = eng.stats["corrected_per_animal"].copy()
    perm = np.random.default_rng(5).permutation(obs.shape[1])
    b = eng.correct(np.ascontiguousarray(obs[:, perm]), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    assert np.array_equal(a[:, perm], b)
    assert np.array_equal(tally_a, eng.stats["corrected_per_animal"])
    changed = int((a != obs).sum())
    assert 0 < changed < obs.size // 10
    # a sample of the big run against the oracle: the first 6 SNP columns as their own chunk
    from oracle.genofix_oracle import Oracle
    sub = np.ascontiguousarray(obs[:, :6])
    o = Oracle(rows, rows[:, 0]).correct_matrix(sub, 0.5, 0.64, back=2, init_filter_p=0.95, tiethreshold=0.4)
    c = eng.correct(sub, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    assert _same_f16(eng.raw_host(), o["raw"])
    assert np.array_equal(c, o["corrected"])


@pytest.mark.parametrize("params", [(0.5, 0.64, 0.95, 0.4), (0.25, 0.5, 0.9, 0.25), (0.75, 0.75, 0.9, 0.0),
                                    (0.5, 0.5, 0.8, 0.5), (1.0 / 3.0, 2.0 / 3.0, 0.9, 1.0 / 3.0)])
def test_exact_integer_decisions_equal_the_literal_replay(params):
    """The product decides from class counts of the children term in exact arithmetic and replays numpy's float64
    sums only inside a 1e-9 guard band; with posteriors requested every cell takes the literal whole-cell kernel.
    Both must give the same calls and tallies -- on thresholds chosen to hit exact rational ties (1/2, 1/4, 3/4, 1/3,
    tie threshold 0) and on a pedigree with loops, missing calls and ungenotyped animals."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    tp, ts, q, tie = params
    for case in range(2):
        if case == 0:
            rows = synth.make_pedigree(3000, 6, seed=77)
            ids = rows[:, 0]
            obs = synth.inject_errors(synth.gene_drop(rows, 256, seed=77), 0.01, seed=77)
        else:
            rows, ids, obs = _random_case(78, n=900, gens=8, s=192, related=0.3, unknown=0.05, miss=0.04, drop=0.1)
        eng = Engine(PedigreeDAG.from_table(rows), ids, 2)
        fast = eng.correct(obs, tp, ts, init_filter_p=q, tiethreshold=tie)
        st_fast = (eng.stats["corrected_per_animal"].copy(), eng.stats["excluded_per_animal"].copy())
        lit = eng.correct(obs, tp, ts, init_filter_p=q, tiethreshold=tie, keep_posteriors=True)
        assert np.array_equal(fast, lit), int((fast != lit).sum())
        assert np.array_equal(st_fast[0], eng.stats["corrected_per_animal"])
        assert np.array_equal(st_fast[1], eng.stats["excluded_per_animal"])
        assert case == 1 or (fast != obs).any()


def test_trio_scan_matches_oracle():
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    rows, ids, obs = _random_case(11, n=600, gens=6, s=1003, miss=0.02, drop=0.15)
    eng = Engine(PedigreeDAG.from_table(rows), ids, 2)
    inc, tested = eng.trio_scan(obs)
    ri, rt = go.trio_scan(go.OraclePedigree(rows), [int(x) for x in ids], obs)
    assert np.array_equal(inc, ri) and np.array_equal(tested, rt)
    assert inc.sum() > 0


def test_window_counts_match_oracle():
    import itertools
    import torch
    from genofix_b200.empirical.jalleledist3 import window_counts_device
    from oracle import genofix_oracle as go
    rs = np.random.default_rng(4)
    n, s = 500, 40
    G = rs.choice([0, 1, 2, 9], size=(n, s), p=[.32, .33, .32, .03]).astype(np.uint8)
    mask = rs.random((n, s)) < 0.97
    for sur in (1, 2, 3):
        wins = [go.ld_window(s, i, sur) for i in range(s)]
        counts = window_counts_device(G, mask, [a for a, b in wins], [b - a for a, b in wins])
        for i, (a, b) in enumerate(wins):
            w = b - a
            if w == 0:
                continue
            ref = go.count_joint_frq(G[:, a:b], mask[:, a:b] & (G[:, a:b] != 9), w, 0.0, sum_inner_count=False)
            vals = np.array([ref[v] for v in itertools.product([0, 1, 2], repeat=w)])
            assert np.array_equal(counts[i, :3 ** w], vals.astype(np.int64)), (sur, i)


@pytest.mark.parametrize("all_pass", [True, False])
def test_window_frequencies_with_missing_calls_match_oracle(all_pass):
    """The stored frequencies (pseudocount, inner-window recurrence, D12) on data WITH missing calls: rows with a 9
    in an outer cell of the window count in the inner windows (jalleledist3.py:133-138)."""
    import itertools
    from genofix_b200.empirical.jalleledist3 import frequencies_from_counts, window_counts_device
    from oracle import genofix_oracle as go
    rs = np.random.default_rng(5)
    n, s = 400, 24
    G = rs.choice([0, 1, 2, 9], size=(n, s), p=[.3, .3, .3, .1]).astype(np.uint8)
    mask = np.ones((n, s), bool) if all_pass else rs.random((n, s)) < 0.97
    for sur in (2, 3, 4):
        wins = [go.ld_window(s, i, sur) for i in range(s)]
        counts = window_counts_device(G, None if all_pass else mask, [a for a, b in wins], [b - a for a, b in wins])
        for i, (a, b) in enumerate(wins):
            w = b - a
            if w == 0 or (w == 9 and i % 4):   # 3^9 states in a Python loop: a quarter of the widest windows
                continue
            ref = go.count_joint_frq(G[:, a:b], mask[:, a:b], w, 1e-4, sum_inner_count=True)
            want = np.array([ref[v] for v in itertools.product([0, 1, 2], repeat=w)])
            assert np.array_equal(frequencies_from_counts(counts[i], w, 1e-4, True), want), (sur, i)


def test_config4_deep_loopy_pedigree_matches_oracle():
    """BASELINE config 4 shape: 15 generations, few sires (inbreeding loops at depth 2/3), 10 % of the
    non-founders with both parents unknown, 5 % missing calls."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle.genofix_oracle import Oracle
    rows = synth.make_pedigree(1500, 15, sire_frac=0.06, dam_frac=0.3, seed=44, related_mating_p=0.25,
                               unknown_parents_frac=0.1)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 48, seed=44), 0.01, seed=44), 0.05, seed=45)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    assert eng.schedule.info["n_loopy"] > 50          # the loopy-blanket paths are really exercised
    out = eng.correct(obs, 0.5, 0.64, init_filter_p=0.9, tiethreshold=0.4)
    o = Oracle(rows, rows[:, 0]).correct_matrix(obs, 0.5, 0.64, back=2, init_filter_p=0.9, tiethreshold=0.4)
    assert _same_f16(eng.raw_host(), o["raw"])
    assert eng.stats["test_p"] == o["test_p"] or (np.isnan(eng.stats["test_p"]) and np.isnan(o["test_p"]))
    assert np.array_equal(out, o["corrected"])
    assert np.array_equal(eng.stats["corrected_per_animal"], o["corrected_per_animal"])


def test_config3_scale_permutation_property():
    """BASELINE config 3 pedigree (100k animals, 10 generations) on a 1024-SNP slice: column permutation
    commutes with correction, tallies are permutation invariant, and the rank kernel agrees with the
    schedule-independent trio scan on which trios are Mendel-consistent."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    rows = synth.make_pedigree(100000, 10, seed=1234)
    obs = synth.inject_errors(synth.gene_drop(rows, 1024, seed=7), 0.01, seed=7)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    a = eng.correct(obs, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    ta = eng.stats["corrected_per_animal"].copy()
    raw_a = eng.raw_host().copy()
    perm = np.random.default_rng(3).permutation(obs.shape[1])
    b = eng.correct(np.ascontiguousarray(obs[:, perm]), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    assert np.array_equal(a[:, perm], b)
    assert np.array_equal(ta, eng.stats["corrected_per_animal"])
    assert np.array_equal(raw_a[:, perm].view(np.uint16), eng.raw_host().view(np.uint16))
    inc, tested = eng.trio_scan(obs)
    # a Mendel-inconsistent trio cell always has a positive score
    ped = eng.schedule._keep
    row_of = ped["row"]
    has = (ped["sire"] >= 0) & (ped["dam"] >= 0)
    kid_rows = row_of[has]
    s_rows, d_rows = row_of[ped["sire"][has]], row_of[ped["dam"][has]]
    sel = np.random.default_rng(1).choice(len(kid_rows), 2000, replace=False)
    k, s, d = obs[kid_rows[sel]], obs[s_rows[sel]], obs[d_rows[sel]]
    T = np.array([[1, .5, 0, .5, .25, 0, 0, 0, 0], [0, .5, 1, .5, .5, .5, 1, .5, 0], [0, 0, 0, 0, .25, .5, 0, .5, 1]]).reshape(3, 3, 3)
    bad = T[k, s, d] == 0
    assert np.array_equal(inc[kid_rows[sel]], bad.sum(axis=1))
    assert np.all(raw_a[kid_rows[sel]][bad].astype(np.float32) > 0)
    assert 0 < int((a != obs).sum()) < obs.size // 10


def _two_pass_and_fused(eng, obs):
    """(raw, hist) of gfx_mendel_rank + gfx_rank_hist and of the fused gfx_mendel_rank_hist on the same matrix"""
    import torch
    eng._alloc(obs.shape[1])
    hin = eng.host_in.numpy()
    hin[:, :eng.s] = obs
    hin[:, eng.s:] = 9
    eng.codes.copy_(eng.host_in)
    eng.load_codes_device(eng.codes)
    res = []
    for fused in (False, True):
        eng.raw.fill_(0x1234)
        eng.mendel_rank_device(fused=fused)
        if not fused:
            plain = eng.raw[:, :eng.s].cpu().numpy().view(np.uint16).copy()
            eng.thresholds_device(0.9)
        torch.cuda.synchronize()
        eng.check_status()
        res.append((eng.raw[:, :eng.s].cpu().numpy().view(np.uint16).copy(), eng.hist.cpu().numpy().view(np.uint64).copy()))
    return plain, res[0], res[1]


@pytest.mark.parametrize("case", ["c2_clean", "missing", "loopy", "big_litters", "ragged"])
def test_fused_rank_histogram_equals_two_passes(case):
    """The fused kernel (lane-private table fast path, packed-half children term, (class, m) counters) against
    the per-cell definition: every score pattern, every 0xFFFF marker and every one of the 65 537 bins."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    if case == "c2_clean":
        rows = synth.make_pedigree(3000, 6, seed=1234)
        obs = synth.inject_errors(synth.gene_drop(rows, 2048, seed=3), 0.01, seed=3)
        ids = rows[:, 0]
    elif case == "missing":
        rows, ids, obs = _random_case(21, n=900, gens=6, s=700, miss=0.04, drop=0.1)
    elif case == "loopy":
        rows = synth.make_pedigree(1500, 15, sire_frac=0.06, dam_frac=0.3, seed=44, related_mating_p=0.25,
                                   unknown_parents_frac=0.1)
        obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 530, seed=44), 0.01, seed=44), 0.02, seed=45)
        ids = rows[:, 0]
    elif case == "big_litters":   # 2 sires per generation: m far above 32 on the sires' rows
        rows = synth.make_pedigree(1200, 4, sire_frac=0.004, dam_frac=0.2, seed=9)
        obs = synth.inject_errors(synth.gene_drop(rows, 512, seed=9), 0.02, seed=9)
        ids = rows[:, 0]
    else:
        rows, ids, obs = _random_case(33, n=300, gens=5, s=2049, miss=0.01, drop=0.0)
    eng = Engine(PedigreeDAG.from_table(rows), ids, 2)
    plain, (raw2, hist2), (raw1, hist1) = _two_pass_and_fused(eng, obs)
    assert np.array_equal(raw1, raw2), "%d score patterns differ" % int((raw1 != raw2).sum())
    assert np.array_equal(hist1, hist2), np.nonzero(hist1 != hist2)[0][:10]
    assert int(hist1.sum()) == obs.size
    assert np.array_equal(raw1 == 0xFFFF, obs > 2)
    # plain (unfused, unmarked) scores: 1.0 where the marker goes
    assert np.array_equal(np.where(raw1 == 0xFFFF, 0x3C00, raw1), plain)


@pytest.mark.parametrize("case", ["clean_wide", "missing", "ragged", "deep_loopy_missing", "deep_loopy_clean", "huge_litters"])
def test_rank_kernel_third_form_equals_second_form(case):
    """k_mendel_rank3 (boolean-network classes, PRMT tables, TMA-staged tasks, work list for loopy rows) against
    k_mendel_rank (GFX_RANK_FORM=2): every score pattern and every one of the 65 537 histogram bins, fused and plain."""
    import os
    import torch
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    if case == "clean_wide":
        rows = synth.make_pedigree(2000, 6, seed=4)
        obs = synth.inject_errors(synth.gene_drop(rows, 4096, seed=4), 0.01, seed=4)
    elif case == "missing":
        rows = synth.make_pedigree(1200, 6, seed=3)
        obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 517, seed=3), 0.01, seed=3), 0.02, seed=3)
    elif case == "ragged":
        rows = synth.make_pedigree(500, 3, seed=7)
        obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 17, seed=7), 0.01, seed=7), 0.3, seed=7)
    elif case.startswith("deep_loopy"):
        rows = synth.make_pedigree(1500, 15, sire_frac=0.02, dam_frac=0.3, seed=11, related_mating_p=0.25, unknown_parents_frac=0.1)
        obs = synth.inject_errors(synth.gene_drop(rows, 640, seed=11), 0.01, seed=11)
        if case.endswith("missing"):
            obs = synth.inject_missing(obs, 0.05, seed=11)
    else:
        n = 3000    # two sires with 1 498 children each: m >= 64 and m >= 1024 paths
        rows = np.zeros((n, 4), dtype=np.int64)
        rows[:, 0] = np.arange(1, n + 1)
        rows[:, 3] = 1 + (np.arange(n) % 2)
        rows[0, 3] = rows[1, 3] = 1
        rows[2, 3] = rows[3, 3] = 2
        rows[4:, 1] = 1 + (np.arange(n - 4) % 2)
        rows[4:, 2] = 3 + (np.arange(n - 4) % 2)
        obs = synth.inject_errors(synth.gene_drop(rows, 333, seed=5), 0.02, seed=5)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    eng._alloc(obs.shape[1])
    hin = eng.host_in.numpy()
    hin[:, :eng.s] = obs
    hin[:, eng.s:] = 9
    eng.codes.copy_(eng.host_in)
    eng.load_codes_device(eng.codes)
    old = os.environ.get("GFX_RANK_FORM")
    try:
        res = {}
        for form in ("2", "3"):
            os.environ["GFX_RANK_FORM"] = form
            for fused in (True, False):
                eng.raw.fill_(-21846)
                eng.mendel_rank_device(fused=fused)
                torch.cuda.synchronize()
                res[(form, fused)] = (eng.raw[:, :eng.s].cpu().numpy().view(np.uint16).copy(),
                                      eng.hist.cpu().numpy().copy() if fused else None)
    finally:
        if old is None:
            os.environ.pop("GFX_RANK_FORM", None)
        else:
            os.environ["GFX_RANK_FORM"] = old
    for fused in (True, False):
        a, b = res[("2", fused)], res[("3", fused)]
        assert np.array_equal(a[0], b[0]), "%d scores differ" % int((a[0] != b[0]).sum())
        if fused:
            assert np.array_equal(a[1], b[1])
            assert int(a[1].sum()) == obs.size


def test_pipelined_chunks_equal_chunk_by_chunk_calls():
    """Engine.correct_chunks (three streams, two buffer sets) against Engine.correct on every chunk, including a
    ragged last chunk that forces a re-allocation in the middle of the pipeline."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    rows = synth.make_pedigree(2000, 6, seed=5)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 1700, seed=5), 0.01, seed=5), 0.01, seed=6)
    bounds = [(0, 512), (512, 1024), (1024, 1536), (1536, 1700)]
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    ref = []
    for a, b in bounds:
        out = eng.correct(np.ascontiguousarray(obs[:, a:b]), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
        ref.append((out, eng.stats["test_p"], eng.stats["corrected_per_animal"].copy(), eng.stats["excluded_per_animal"].copy()))
    eng2 = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    got = list(eng2.correct_chunks((np.ascontiguousarray(obs[:, a:b]) for a, b in bounds), 0.5, 0.64,
                                   init_filter_p=0.95, tiethreshold=0.4))
    assert len(got) == len(ref)
    for (out, st), (rout, rtp, rc, rx) in zip(got, ref):
        assert np.array_equal(out, rout)
        assert st["test_p"] == rtp
        assert np.array_equal(st["corrected_per_animal"], rc) and np.array_equal(st["excluded_per_animal"], rx)
    assert sum(int((o != obs[:, a:b]).sum()) for (o, _), (a, b) in zip(got, bounds)) > 0


@pytest.mark.parametrize("lanes", [2, 3])
def test_multilane_results_do_not_depend_on_lane_count(lanes):
    """Two / three lanes (one host thread and stream set each, one shared schedule; bench.py runs three) give the
    chunk-by-chunk results."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine, MultiLane
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    rows = synth.make_pedigree(2000, 6, seed=15)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 1500, seed=15), 0.01, seed=15), 0.01, seed=16)
    bounds = [(0, 300), (300, 600), (600, 900), (900, 1200), (1200, 1500)]
    ped = PedigreeDAG.from_table(rows)
    eng = Engine(ped, rows[:, 0], 2)
    ref = []
    for a, b in bounds:
        out = eng.correct(np.ascontiguousarray(obs[:, a:b]), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
        ref.append((out, eng.stats["test_p"], eng.stats["corrected_per_animal"].copy()))
    ml = MultiLane(ped, rows[:, 0], 2, lanes=lanes)
    from genofix_b200.engine import pinned_array
    # a mix of pinned (moved by the copy engines directly) and pageable (staged) inputs and outputs
    outs = [pinned_array((len(rows), b - a)) if i % 2 else np.empty((len(rows), b - a), np.uint8) for i, (a, b) in enumerate(bounds)]
    ins = []
    for i, (a, b) in enumerate(bounds):
        x = pinned_array((len(rows), b - a)) if i % 3 == 0 else np.empty((len(rows), b - a), np.uint8)
        x[...] = obs[:, a:b]
        ins.append(x)
    got = list(ml.correct_chunks(ins, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4, outs=outs))
    for k, ((out, st), (rout, rtp, rc)) in enumerate(zip(got, ref)):
        assert out is outs[k]
        assert np.array_equal(out, rout) and st["test_p"] == rtp and np.array_equal(st["corrected_per_animal"], rc)


def test_huge_litter_paths_match_oracle():
    """One sire with 2 300 genotyped children: the children units exceed the packed-half range of the rank kernel
    (m >= 1024 -> per-cell path) and the litter exceeds the in-register children term of the correction kernel
    (fallback kernel); both must still be bit-exact."""
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from genofix_b200 import synth
    from oracle.genofix_oracle import Oracle
    rs = np.random.default_rng(77)
    rows = [[1, 0, 0, 1]] + [[i, 0, 0, 2] for i in range(2, 42)]          # 1 sire, 40 dams (founders)
    kid = 42
    for k in range(2300):
        rows.append([kid, 1, int(rs.integers(2, 42)), 1 + k % 2])
        kid += 1
    rows = np.array(rows, dtype=np.int64)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 8, seed=5), 0.02, seed=5), 0.01, seed=6)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    out = eng.correct(obs, 0.5, 0.64, init_filter_p=0.9, tiethreshold=0.4)
    o = Oracle(rows, rows[:, 0]).correct_matrix(obs, 0.5, 0.64, back=2, init_filter_p=0.9, tiethreshold=0.4)
    assert _same_f16(eng.raw_host(), o["raw"])
    assert float(eng.raw_host()[0].astype(np.float32).max()) > 341.0      # m >= 1024 really happened on the sire's row
    assert eng.stats["test_p"] == o["test_p"]
    assert np.array_equal(out, o["corrected"])
    assert np.array_equal(eng.stats["corrected_per_animal"], o["corrected_per_animal"])


def test_packed_io_equals_uint8_io(tmp_path):
    """correct_chunks(packed_io=True) fed from a GFX2BIT1 container (word-aligned SNP blocks, ragged last block, pinned and
    pageable buffers) returns the packed form of what the uint8 path returns."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine, MultiLane, pinned_array
    from genofix_b200.gfutils.packedgens import PackedGens, pack_codes, unpack_codes
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    rows = synth.make_pedigree(2000, 6, seed=15)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 1100, seed=15), 0.01, seed=15), 0.01, seed=16)
    path = str(tmp_path / "m.g2b")
    PackedGens.write(path, rows[:, 0], ["snp%d" % i for i in range(obs.shape[1])], obs)
    g = PackedGens(path)
    bounds = [(0, 400), (400, 800), (800, 1100)]
    ped = PedigreeDAG.from_table(rows)
    eng = Engine(ped, rows[:, 0], 2)
    ref = [eng.correct(np.ascontiguousarray(obs[:, a:b]), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4) for a, b in bounds]
    chunks = []
    for i, (a, b) in enumerate(bounds):
        blk = g.packed_columns(a, b, out=pinned_array((len(rows), ((b - a + 15) // 16 + 3) // 4 * 4), np.uint32) if i % 2 == 0 else None)
        chunks.append((blk, b - a))
    ml = MultiLane(ped, rows[:, 0], 2, lanes=2)
    got = list(ml.correct_chunks(chunks, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4, packed_io=True))
    for (P, st), r, (a, b) in zip(got, ref, bounds):
        assert P.dtype == np.uint32 and np.array_equal(unpack_codes(P, b - a), r)
        assert np.array_equal(P, pack_codes(r))


def test_very_wide_eliminations_get_a_result():
    """A small, inbred population (6 sires per generation) with 1 % missing calls: suspect filtering opens more than nine
    variables of a pair blanket at once.  That used to end in GFX_E_TOO_WIDE; the third-level table pool answers it."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle.genofix_oracle import Oracle
    rows = synth.make_pedigree(1500, 6, seed=25)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 1100, seed=25), 0.01, seed=25), 0.01, seed=26)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    for a, b in [(0, 400), (400, 800), (800, 1100)]:
        out = eng.correct(np.ascontiguousarray(obs[:, a:b]), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)   # must not raise
        assert out.shape == (len(rows), b - a)
    sub = np.ascontiguousarray(obs[:, :32])
    out = eng.correct(sub, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    o = Oracle(rows, rows[:, 0]).correct_matrix(sub, 0.5, 0.64, back=2, init_filter_p=0.95, tiethreshold=0.4)
    assert _same_f16(eng.raw_host(), o["raw"]) and np.array_equal(out, o["corrected"])


def test_device_simulator_properties_and_evaluation():
    """Device gene drop / error injection / evaluation: Mendelian consistency of the drop (trio scan finds nothing),
    allele frequency 1/2, error rate and 'never the same genotype', evaluation counts equal to numpy on the unpacked
    matrices, and the whole simulate loop (drop -> errors -> correct -> score) on packed device data."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.gfutils.packedgens import unpack_codes
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from genofix_b200.simulate_device import DeviceSimulator
    rows = synth.make_pedigree(3000, 6, seed=8)
    s = 1000
    sim = DeviceSimulator(rows, s, seed=99)
    truth = sim.gene_drop()
    T = unpack_codes(truth.cpu().numpy().view(np.uint32), s)
    assert set(np.unique(T)) <= {0, 1, 2} and abs(T.mean() - 1.0) < 0.02
    assert np.all(truth.cpu().numpy().view(np.uint32)[:, (s + 15) // 16:] == 0xFFFFFFFF)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    inc, tested = eng.trio_scan(T)
    assert inc.sum() == 0 and tested.sum() > 0                       # a clean drop is Mendel consistent
    assert np.array_equal(unpack_codes(sim.gene_drop().cpu().numpy().view(np.uint32), s), T)   # deterministic in the seed
    obs = sim.inject_errors(truth, 0.01)
    O = unpack_codes(obs.cpu().numpy().view(np.uint32), s)
    rate = (O != T).mean()
    assert 0.008 < rate < 0.012 and set(np.unique(O)) <= {0, 1, 2}
    inc2, _ = eng.trio_scan(O)
    assert inc2.sum() > 0
    # correct the packed matrix in place on the device and score it there
    eng.use_packed(obs, s)
    eng.correct_resident(0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    K = unpack_codes(eng.packed_live.cpu().numpy().view(np.uint32), s)
    ev = sim.confusion(truth, obs, eng.packed_live)
    was_err, changed = T != O, K != O
    assert ev["errors"] == int(was_err.sum()) and ev["cells"] == T.size
    assert ev["fixed"] == int((was_err & (K == T)).sum())
    assert ev["blanked"] == int((was_err & changed & (K == 9)).sum())
    assert ev["wrong_fix"] == int((was_err & changed & (K != T) & (K != 9)).sum())
    assert ev["new_errors"] == int((~was_err & changed).sum())
    assert ev["missed"] == int((was_err & ~changed).sum())
    assert ev["fixed"] > 0.1 * ev["errors"] and ev["new_errors"] < ev["fixed"]      # the correction does net good
    print(ev)


# ---- pedigree/error_checker.py: per-animal Mendel tallies in the reference's float16 arithmetic -------------------
def _errchk_frame(g):
    import pandas as pd
    return pd.DataFrame(g["obs"], index=[int(x) for x in g["ids"]], columns=["SNP%d" % i for i in range(g["obs"].shape[1])])


@pytest.mark.parametrize("name", ["errchk_a", "errchk_b", "errchk_long"])
def test_error_checker_matches_reference_golden(name):
    """probs_errors.csv and individual_sum_probs_errors.csv of the UNMODIFIED reference (error_checker.py:199-204),
    bit for bit: k_mendel_rank<HIST> -> float16 table -> k_rank_rowsum_f16."""
    from genofix_b200.pedigree.error_checker import mendel_error_tally
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    g = load_golden(name)
    pe, sums = mendel_error_tally(_errchk_frame(g), PedigreeDAG.from_table(g["ped_rows"]), g["kids"])
    assert [int(x) for x in pe.index] == [int(x) for x in g["with_parents"]]
    assert pe.to_numpy().dtype == np.float16
    assert np.array_equal(pe.to_numpy().view(np.uint16), g["probs_errors"])
    assert sums.dtype == np.float16 and np.array_equal(sums.view(np.uint16), g["sums"])


@pytest.mark.parametrize("missing_value", [-1.0, 1.0])
def test_error_checker_matches_oracle_seeded(missing_value):
    """Seeded pedigree with ungenotyped animals, missing calls and a kids list that is a subset; missing_value = 1 is
    preindex.py:262-263."""
    import pandas as pd
    from genofix_b200 import synth
    from genofix_b200.pedigree.error_checker import mendel_error_tally
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    rows = synth.make_pedigree(260, 6, sire_frac=0.2, dam_frac=0.5, seed=41, related_mating_p=0.3)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 150, seed=5), 0.05, seed=6), 0.04, seed=7)
    rs = np.random.Generator(np.random.PCG64(8))
    keep = rs.random(len(rows)) >= 0.15
    ids, obs = rows[keep, 0], obs[keep]
    kids = [int(x) for x in ids if rs.random() < 0.7]
    df = pd.DataFrame(obs, index=[int(x) for x in ids], columns=["S%d" % i for i in range(obs.shape[1])])
    pe, sums = mendel_error_tally(df, PedigreeDAG.from_table(rows), kids, missing_value=missing_value)
    wp, _cand, E, osums = go.error_checker(rows, ids, obs, kids, missing_value=missing_value)
    assert len(wp) > 50 and [int(x) for x in pe.index] == wp
    assert np.array_equal(pe.to_numpy().view(np.uint16), E.view(np.uint16))
    assert np.array_equal(sums.view(np.uint16), osums.view(np.uint16))


def test_mendel_tally_is_the_float16_chain_at_scale():
    """3 000 animals x 4 099 SNPs (ragged width, accumulators far beyond 2048 where float16 steps are 2): the kernel's
    sums equal the column-by-column float16 chain over the very matrix it summed, for all rows and for a row subset."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    rows = synth.make_pedigree(3000, 6, sire_frac=0.1, dam_frac=0.5, seed=17)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 4099, seed=2), 0.2, seed=3), 0.03, seed=4)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    E, sums, m16 = eng.mendel_tally(obs, None, 1.0)
    assert E.shape == obs.shape and E.dtype == np.float16
    assert np.all(E[obs == 9] == np.float16(1.0)) and np.nanmax(E[obs != 9]) == np.float16(1.0)
    assert np.array_equal(sums.view(np.uint16), go.half_rowsum(E).view(np.uint16))
    assert float(np.nanmax(sums)) > 600
    sel = np.random.Generator(np.random.PCG64(1)).permutation(3000)[:777].astype(np.int32)
    E2, sums2, _ = eng.mendel_tally(None, sel, -1.0)
    assert np.array_equal(E2 == np.float16(-1.0), obs[sel] == 9)
    assert np.array_equal(sums2.view(np.uint16), go.half_rowsum(E2).view(np.uint16))


def test_error_checker_cli_writes_the_reference_files(tmp_path):
    """src/pedigree/error_checker.py with the reference's command line on files laid out as the reference needs them
    (oracle/ref_harness.run_error_checker); the two csv files parse back to the golden float16 values."""
    import subprocess
    import sys
    import pandas as pd
    from conftest import ROOT
    g = load_golden("errchk_a")
    d = str(tmp_path)
    S = g["obs"].shape[1]
    snps = ["SNP%d" % i for i in range(S)]
    with open(d + "/gens.txt", "w") as f:
        f.write("id " + " ".join(snps) + " pad\n")
        for k, row in zip(g["ids"], g["obs"]):
            f.write(str(int(k)) + " " + " ".join(str(int(x)) for x in row) + " 0\n")
    with open(d + "/ped.txt", "w") as f:
        for r in g["ped_rows"]:
            f.write(" ".join(str(int(x)) for x in r[:4]) + "\n")
    with open(d + "/kids.txt", "w") as f:
        f.write("kid\n" + "".join("%d\n" % int(k) for k in g["kids"]))
    with open(d + "/map.csv", "w") as f:
        f.write("snpid,chrom,pos,topAllele,B\n")
        for i, s in enumerate(snps):
            f.write("%s,chrA,%d,A,B\n" % (s, 100 * i + 1))
        f.write("pad,chrB,1,A,B\n")
    subprocess.run([sys.executable, ROOT + "/src/pedigree/error_checker.py", "-g", d + "/gens.txt", "-p", d + "/ped.txt",
                    "-k", d + "/kids.txt", "-s", d + "/map.csv", "-c", "chrA", "-o", d + "/out"], check=True)
    pe = pd.read_csv(d + "/out/probs_errors.csv", index_col=0)
    assert [int(x) for x in pe.index] == [int(x) for x in g["with_parents"]]
    assert np.array_equal(pe.to_numpy().astype(np.float16).view(np.uint16), g["probs_errors"])
    sums = np.loadtxt(d + "/out/individual_sum_probs_errors.csv", delimiter=",", ndmin=1)
    assert np.array_equal(sums.astype(np.float16).view(np.uint16), g["sums"])


def test_rank_rowsum_float64_partial_sums():
    """gfx_rank_rowsum: per-animal float64 sums of E (the quantity SNP-sharded ranks all-reduce) against
    numpy on the same E matrix to 1e-12 relative and repeatable bit for bit."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    rows = synth.make_pedigree(1200, 6, sire_frac=0.1, dam_frac=0.5, seed=9)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 1003, seed=2), 0.02, seed=3), 0.03, seed=4)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    eng.correct(obs, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    for mv in (1.0, -1.0):
        a = eng.rank_rowsum_device(mv).cpu().numpy()
        b = eng.rank_rowsum_device(mv).cpu().numpy()
        assert np.array_equal(a, b, equal_nan=True)
        E = eng.E_host()
        E[obs == 9] = mv
        ref = E.sum(axis=1)
        nan = np.isnan(ref)           # half-sib matings give loopy blankets: contradictory evidence there is a NaN rank (D16)
        assert np.array_equal(np.isnan(a), nan)
        assert np.all(np.abs(a[~nan] - ref[~nan]) <= 1e-12 * np.maximum(1.0, np.abs(ref[~nan])))
    assert (~nan).sum() > 600 and float(np.nanmax(a)) > 0


# ---- empirical/preindex.py: Mendel-filtered LD index ----------------------------------------------------------------
def _preidx_inputs(g, cand):
    import pandas as pd
    row = {int(k): i for i, k in enumerate(g["ids"])}
    names = ["SNP%d" % i for i in range(g["obs"].shape[1])]
    df = pd.DataFrame(g["obs"][[row[k] for k in cand]], index=cand, columns=names)
    c2s = {"0": ["pad"]}
    for n, c in zip(names, g["chrom"]):
        c2s.setdefault(str(c), []).append(n)
    return df, c2s


@pytest.mark.parametrize("name", ["preidx_a", "preidx_b"])
def test_preindex_matches_reference_golden(name):
    """The indexes the UNMODIFIED reference pickled (preindex.py:339-346): rank cut, cells under the cut and every
    window frequency, bit for bit: k_mendel_rank<HIST> -> histogram quantile -> mask -> k_window_counts."""
    from genofix_b200.empirical.preindex import index_chunk, seed_chunks
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    g = load_golden(name)
    ped = PedigreeDAG.from_table(g["ped_rows"])
    chunks = seed_chunks(ped, g["ids"], int(g["seed_chunk"]))
    assert len(chunks) == 1
    df, c2s = _preidx_inputs(g, chunks[0][1])
    res = index_chunk(df, ped, c2s, int(g["surround"]), float(g["q"]))
    assert sorted(res) == [str(c) for c in g["chroms"]]
    for ci, c in enumerate(sorted(res)):
        empC, info = res[c]
        assert info["quantQ"] == g["quantQ"][ci]
        assert (info["n_unmasked"], len(info["eval_ids"]) * len(c2s[c])) == tuple(g["kept"][ci])
        for i, s in enumerate(empC.snp_ordered):
            w = int(g["wlen_" + c][i])
            assert len(empC.windows[s]) == w
            if w:
                assert np.array_equal(empC._freq[s], g["freq_" + c][i][:3 ** w])


def test_preindex_matches_oracle_on_several_seed_chunks():
    """Seeded pedigree cut into seed chunks of 60 kids, three chromosomes of different length (one below the
    2-SNP minimum), surround 3 (7-SNP windows): every chunk's index against the oracle."""
    from genofix_b200 import synth
    from genofix_b200.empirical.preindex import index_chunk, seed_chunks
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    import pandas as pd
    rows = synth.make_pedigree(240, 6, sire_frac=0.2, dam_frac=0.5, seed=77, related_mating_p=0.2)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 41, seed=8), 0.05, seed=9), 0.03, seed=10)
    ids = rows[:, 0]
    names = ["S%d" % i for i in range(41)]
    chrom = ["2"] * 24 + ["10"] * 16 + ["7"]
    cc, c2s = {}, {}
    for i, c in enumerate(chrom):
        cc.setdefault(c, []).append(i)
        c2s.setdefault(c, []).append(names[i])
    ped = PedigreeDAG.from_table(rows)
    oped = go.OraclePedigree(rows)
    chunks = seed_chunks(ped, ids, 60)
    assert len(chunks) >= 3 and chunks == go.preindex_chunks(oped, ids, 60)
    row = {int(k): i for i, k in enumerate(ids)}
    for kid_ids, cand in chunks[:3]:
        df = pd.DataFrame(obs[[row[k] for k in cand]], index=cand, columns=names)
        res = index_chunk(df, ped, c2s, 3, 0.85)
        ref = go.preindex_chunk(oped, ids, obs, cand, cc, 3, 0.85)
        assert sorted(res) == sorted(ref) == ["10", "2"]
        for c in res:
            empC, info = res[c]
            assert info["eval_ids"] == ref[c]["eval_ids"]
            assert info["quantQ"] == ref[c]["quantQ"] and info["n_unmasked"] == ref[c]["n_unmasked"]
            assert np.array_equal(info["individualSumProbs"].view(np.uint16), ref[c]["sums"].view(np.uint16))
            for i, s in enumerate(empC.snp_ordered):
                f = ref[c]["freq"][i]
                assert len(empC.windows[s]) == (0 if len(f) == 0 else round(np.log(len(f)) / np.log(3)))
                if len(f):
                    assert np.array_equal(empC._freq[s], f)


def test_preindex_cli_writes_the_reference_indexes(tmp_path):
    """src/empirical/preindex.py with the reference's command line; the pickled indexes hold the golden frequencies."""
    import gzip
    import pickle
    import subprocess
    import sys
    from conftest import ROOT
    g = load_golden("preidx_a")
    d = str(tmp_path)
    S = g["obs"].shape[1]
    snps = ["SNP%d" % i for i in range(S)]
    with open(d + "/gens.txt", "w") as f:
        f.write("id " + " ".join(snps) + " pad\n")
        for k, r in zip(g["ids"], g["obs"]):
            f.write(str(int(k)) + " " + " ".join(str(int(x)) for x in r) + " 0\n")
    with open(d + "/ped.txt", "w") as f:
        for r in g["ped_rows"]:
            f.write(" ".join(str(int(x)) for x in r[:4]) + "\n")
    with open(d + "/map.csv", "w") as f:
        for i, s in enumerate(snps):
            f.write("%s,%s,0,%d\n" % (g["chrom"][i], s, 100 * i + 1))
        f.write("0,pad,0,1\n")
    subprocess.run([sys.executable, ROOT + "/src/empirical/preindex.py", "-g", d + "/gens.txt", "-p", d + "/ped.txt",
                    "-s", d + "/map.csv", "-w", str(int(g["surround"])), "-q", str(float(g["q"])), "-o", d + "/out"], check=True)
    for c in (str(x) for x in g["chroms"]):
        with gzip.open(d + "/out/%s/empiricalIndex.idx.gz" % c, "rb") as f:
            empC = pickle.load(f)
        for i, s in enumerate(empC.snp_ordered):
            w = int(g["wlen_" + c][i])
            assert len(empC.windows[s]) == w
            if w:
                tab = empC.snp2frequency[s]
                import itertools
                vals = np.array([tab[tuple(zip(empC.windows[s], v))] for v in itertools.product([0, 1, 2], repeat=w)])
                assert np.array_equal(vals, g["freq_" + c][i][:3 ** w])


# ---- the reference's static call surface (what its own worker pools call) -------------------------------------------
@pytest.mark.timeout(180)
def test_static_call_surface_matches_oracle():
    """initializer(matrix, pedigree, probs_errors) + CorrectGenotypes.mendelProbsSingle / correctPair(back + 1) /
    correctSingle(back), called the way correct_genotypes3.py:569-590,684-693,851-859 calls them, against the oracle on the
    ORIGINAL matrix (the statics only compute results, they apply nothing)."""
    import pandas as pd
    from genofix_b200 import correct_genotypes3 as cg
    from genofix_b200 import synth
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle.genofix_oracle import Oracle
    rows = synth.make_pedigree(150, 5, sire_frac=0.25, dam_frac=0.5, seed=21, related_mating_p=0.3)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 40, seed=3), 0.04, seed=4), 0.03, seed=5)
    ids = [int(x) for x in rows[:, 0]]
    names = ["S%d" % i for i in range(obs.shape[1])]
    df = pd.DataFrame(obs, index=ids, columns=names)
    ped = PedigreeDAG.from_table(rows)
    o = Oracle(rows, ids)
    # rank scores, one animal at a time (:569-590)
    cg.initializer(df, ped, None)
    raw = np.vstack([np.squeeze(cg.CorrectGenotypes.mendelProbsSingle(k, 2)[1]) for k in ids])
    oraw = o.mendel_rank(obs, 2)
    assert _same_f16(raw, oraw)
    E, test_p = o.transform(oraw, obs, 0.9)
    if np.isnan(test_p):
        pytest.skip("NaN rank in the seeded case")
    cg.initializer(df, ped, pd.DataFrame(E, index=ids, columns=names))
    pairs = sorted(set(p for p in (o.ped.parents(k) for k in ids) if p[0] is not None and p[1] is not None))
    checked = 0
    for sire, dam in pairs:
        ref = o.correct_pair(obs, E, sire, dam, 3, test_p)
        got = cg.CorrectGenotypes.correctPair(sire, dam, back=3, tiethreshold=0.05, test_p=test_p, suspect_t=0.2)
        assert (ref is None) == (got is None)
        if ref is None:
            continue
        assert set(got.prob_sire.keys()) == set(names[j] for j in ref)
        for j, (ps, pd_, _ops, _opd) in ref.items():
            assert _log_close(got.prob_sire[names[j]], ps) and _log_close(got.prob_dam[names[j]], pd_)
            checked += 1
    paired = set(x for p in pairs for x in p)
    for kid in (k for k in ids if k not in paired):
        ref = o.correct_single(obs, E, kid, 2, test_p)
        got = cg.CorrectGenotypes.correctSingle(kid, back=2, tiethreshold=0.05, test_p=test_p, suspect_t=0.2)
        assert (ref is None) == (got is None)
        if ref is None:
            continue
        assert set(got.prob.keys()) == set(names[j] for j in ref)
        for j, (p, _op) in ref.items():
            assert _log_close(got.prob[names[j]], p)
            checked += 1
    assert checked > 50


@pytest.mark.timeout(180)
def test_sharded_error_tally_single_rank_matches_float64_oracle():
    """mendel_error_tally_sharded with world = 1 and 7-SNP chunks (histogram pass, table, row-sum pass) against the
    float64 form of error_checker.py:193-204 computed from the oracle's scores."""
    from genofix_b200.pedigree.error_checker import mendel_error_tally_sharded
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    g = load_golden("errchk_b")
    wp, sums = mendel_error_tally_sharded(_errchk_frame(g), PedigreeDAG.from_table(g["ped_rows"]), g["kids"], chunk_snps=7)
    assert [int(x) for x in wp] == [int(x) for x in g["with_parents"]]
    oped = go.OraclePedigree(g["ped_rows"])
    owp, cand = go.trio_candidates(oped, g["ids"], g["kids"])
    row = {int(k): i for i, k in enumerate(g["ids"])}
    Gc = g["obs"][[row[k] for k in cand]]
    raw = go.Oracle(oped, cand).mendel_rank(Gc, 2)
    with np.errstate(all="ignore"):
        E = np.log(np.log(raw.astype(np.float64) + 1) + 1)
        E = E / np.nanmax(E)
    E[Gc == 9] = -1.0
    pos = {k: i for i, k in enumerate(cand)}
    ref = E[[pos[k] for k in owp]].sum(axis=1)
    ref = ref / np.nanmax(ref)
    assert np.allclose(sums, ref, rtol=1e-12, atol=1e-15)


def test_error_checker_matches_reference_on_example_pedigree():
    """error_checker.py on the reference's own examples/simulate/example.fam (4 560 animals, 4 235 kids with both parents
    genotyped) x 12 SNPs: 104 s of reference time, every float16 rank and every per-animal sum bit for bit."""
    test_error_checker_matches_reference_golden("errchk_example")


@pytest.mark.parametrize("sur", [1, 2, 3])
def test_count_joint_frq_all_matches_reference_index(sur):
    """countJointFrqAll on the GPU against the frequencies the reference's own countJointFrqAll stored for the same
    table (kats2.npz; jalleledist3.py:143-174), then getCountTable on the device-built index."""
    import pandas as pd
    from genofix_b200.empirical.jalleledist3 import JointAllellicDistribution
    g = load_golden("kats2")
    G = g["emp_G"]
    n, s = G.shape
    names = ["SNP%d" % i for i in range(s)]
    emp = JointAllellicDistribution("1", names, surround_size=sur)
    emp.countJointFrqAll(pd.DataFrame(G, index=np.arange(1, n + 1), columns=names))
    for i, snp in enumerate(names):
        w = int(g["emp_wlen_%d" % sur][i])
        if w:
            assert np.array_equal(emp._freq[snp], g["emp_freq_%d" % sur][i][:3 ** w]), snp
    ct = g["emp_counttable_%d" % sur]
    for a in range(0, ct.shape[0], 7):
        ev = {snp: int(G[a, j]) for j, snp in enumerate(names)}
        for j, snp in enumerate(names):
            assert np.array_equal(np.asarray(emp.getCountTable(ev, snp), dtype=float), ct[a, j])


def test_linked_gene_drop_follows_quickgsim_meiosis():
    """gfx_gene_drop_linked (quickgsim/genome.py:74-115; the reference's generator is unseeded, so properties): with parents
    whose two strands are all-0 / all-1 the transmitted alleles ARE the strand sequence: per gamete and chromosome at least one
    switch, never at the first or the last marker, their number distributed like max(1, Poisson(Morgans)), their positions
    like the genetic distances; Mendelian consistency; determinism in the seed."""
    import torch
    from genofix_b200.simulate_device import DeviceSimulator, linkage_arrays
    n_kids, s = 4000, 1000
    rows = np.zeros((2 + n_kids, 4), dtype=np.int64)
    rows[:, 0] = np.arange(1, 3 + n_kids)
    rows[0, 3], rows[1, 3] = 1, 2
    rows[2:, 1], rows[2:, 2], rows[2:, 3] = 1, 2, 1
    # three chromosomes: 400 markers / 1.5 M evenly spaced, 350 markers / 0.6 M with a hot interval, 250 markers / 2.5 M
    chrom = np.array([1] * 400 + [2] * 350 + [3] * 250)
    cm2 = np.concatenate([np.linspace(0.1, 10, 175), np.linspace(50, 60, 175)])       # 40 cM between markers 174 and 175
    cm = np.concatenate([np.linspace(0.375, 150, 400), cm2, np.linspace(1, 250, 250)])
    start, morgans, cum_p = linkage_arrays(chrom, cm)
    assert list(start) == [0, 400, 750, 1000] and np.allclose(morgans, [1.5, 0.6, 2.5])
    sim = DeviceSimulator(rows, s, seed=77)
    het = torch.full((sim.wpr,), 0xAAAAAAAA - (1 << 32), dtype=torch.int32, device=sim.hap.device)   # strand 0 all 0, strand 1 all 1
    sim.hap[0] = het
    sim.hap[1] = het
    truth = sim.gene_drop(linkage_map=(chrom, cm), first_level=1)
    hap = sim.hap[2:].cpu().numpy().view(np.uint32)
    j = np.arange(s)
    bits = (hap[:, j // 16] >> (2 * (j % 16)).astype(np.uint32)) & 3
    for side in (0, 1):
        strand = (bits >> side) & 1                                  # [kids, s]: which strand of the parent
        for c, (a, b) in enumerate(zip(start[:-1], start[1:])):
            seg = strand[:, a:b]
            sw = seg[:, 1:] != seg[:, :-1]                           # sw[:, i]: switch AT marker i + 1
            nsw = sw.sum(axis=1)
            assert nsw.min() >= 1                                    # obligatory crossover
            assert not sw[:, -1].any()                               # never at the last marker; the first cannot show by construction
            lam = float(morgans[c])
            k = np.arange(60)
            from math import exp, factorial
            pk = np.array([exp(-lam) * lam ** i / factorial(i) for i in k])
            expect = (np.maximum(k, 1) * pk).sum()
            assert abs(nsw.mean() - expect) < 5 * np.sqrt(max(lam, 0.3) / n_kids) + 0.02, (side, c, nsw.mean(), expect)
            assert 0.45 < seg[:, 0].mean() < 0.55                    # random start strand
        # the hot interval of chromosome 2 (40 of the 59.8 cM between its interior markers): a gamete with n crossovers on
        # distinct markers holds it with probability ~ 1 - (1 - p)^n, so its share of all crossovers is E[that] / E[n]
        seg = strand[:, start[1]:start[2]]
        sw = seg[:, 1:] != seg[:, :-1]
        p_hot = 40.0 / (cm2[-2] - cm2[0])
        lam = float(morgans[1])
        pn = {n: exp(-lam) * lam ** n / factorial(n) for n in range(2, 40)}
        pn[1] = exp(-lam) * (1 + lam)
        share = sum(q * (1 - (1 - p_hot) ** n) for n, q in pn.items()) / sum(q * n for n, q in pn.items())
        assert abs(sw[:, 174].sum() / sw.sum() - share) < 0.03
    codes = truth[2:].cpu().numpy().view(np.uint32)
    g = (codes[:, j // 16] >> (2 * (j % 16)).astype(np.uint32)) & 3
    assert g.max() <= 2                                              # het x het parents: every genotype possible, none missing
    sim2 = DeviceSimulator(rows, s, seed=77)
    sim2.hap[0] = het
    sim2.hap[1] = het
    sim2.gene_drop(linkage_map=(chrom, cm), first_level=1)
    assert torch.equal(sim.hap, sim2.hap)
    # a multi-generation drop with random founders is Mendelian-consistent
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    prow = synth.make_pedigree(600, 5, seed=3)
    sim3 = DeviceSimulator(prow, s, seed=5)
    t3 = sim3.gene_drop(linkage_map=(chrom, cm))
    eng = Engine(PedigreeDAG.from_table(prow), prow[:, 0], 2)
    eng._alloc(s)
    eng.use_packed(t3[:, :eng.wpr].contiguous(), s)
    inc, tested = eng.trio_scan()
    assert int(inc.sum()) == 0 and int(tested.sum()) > 0


# ---- the empirical (LD-window) blend: correctMatrix WITH an index ----------------
def _ld_index(obs, names, surround, pseudocount):
    import pandas as pd
    from genofix_b200.empirical.jalleledist3 import JointAllellicDistribution
    emp = JointAllellicDistribution("1", names, surround_size=surround, pseudocount=pseudocount)
    emp.countJointFrqAll(pd.DataFrame(obs, index=np.arange(1, obs.shape[0] + 1), columns=names))
    return emp


@pytest.mark.parametrize("name", ["emp_a", "emp_b"])
def test_empirical_blend_matches_reference_golden(name):
    """The reference's own correctMatrix run with an LD index it built itself (oracle/ref_harness.run_correct_matrix(emp=...)):
    index, threshold, corrected calls and error count, through the reference-facing class."""
    import pandas as pd
    from genofix_b200.correct_genotypes3 import CorrectGenotypes
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    g = load_golden(name)
    obs = g["obs"]
    names = ["SNP%d" % i for i in range(obs.shape[1])]
    emp = _ld_index(obs, names, int(g["emp_surround"]), float(g["emp_pseudocount"]))
    for j, snp in enumerate(names):
        w = int(g["emp_wlen"][j])
        if w:
            assert np.array_equal(emp._freq[snp][:3 ** w], g["emp_freq"][j, :3 ** w]), snp
    tp, ts, back, q, tie, et = g["params"]
    cg = CorrectGenotypes()
    df = pd.DataFrame(obs, index=[int(x) for x in g["ids"]], columns=names)
    out = cg.correctMatrix(df, PedigreeDAG.from_table(g["ped_rows"]), emp, tp, ts, back=int(back), init_filter_p=q,
                           tiethreshold=tie, err_thresh=et, weight_empirical=float(g["emp_weight"]))
    assert cg.last_stats["test_p"] == float(g["test_p"])
    got = out.to_numpy()
    assert np.array_equal(got, g["corrected"]), "%d corrected calls differ" % int((got != g["corrected"]).sum())
    assert int(cg.last_stats["corrected_per_animal"].sum()) == int(g["errors_all"])
    # and the run without an index gives something else: the blend is really in the decision
    plain = cg.correctMatrix(df, PedigreeDAG.from_table(g["ped_rows"]), None, tp, ts, back=int(back), init_filter_p=q,
                             tiethreshold=tie, err_thresh=et).to_numpy()
    assert not np.array_equal(plain, got)


@pytest.mark.parametrize("seed,surround,miss", [(5, 1, 0.0), (6, 2, 0.03), (7, 3, 0.02)])
def test_empirical_blend_matches_oracle_on_seeded_inputs(seed, surround, miss):
    """Seeded matrices, also with missing calls inside the windows (summed over their completions: the oracle defines that
    case, the reference raises) and with a tie band, which writes 9s that later windows then see."""
    from genofix_b200 import synth
    from genofix_b200.correct_genotypes3 import ld_index_arrays
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    rows = synth.make_pedigree(150, 6, sire_frac=0.25, dam_frac=0.5, seed=seed, related_mating_p=0.3)
    s = 48
    obs = synth.inject_errors(synth.gene_drop(rows, s, seed=seed), 0.05, seed=seed)
    if miss:
        obs = synth.inject_missing(obs, miss, seed=seed)
    names = ["SNP%d" % i for i in range(s)]
    emp = _ld_index(obs, names, surround, 1e-4)
    oemp = go.emp_index(obs, surround, 1e-4, weight=3)
    for j, snp in enumerate(names):
        w = int(oemp["wlen"][j])
        if w:
            assert np.array_equal(emp._freq[snp][:3 ** w], oemp["freq"][j]), snp
    params = dict(back=2, init_filter_p=0.8, tiethreshold=0.1, err_thresh=1)
    want = go.Oracle(rows, rows[:, 0]).correct_matrix(obs, 0.4, 0.5, emp=oemp, **params)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    got = eng.correct_emp(obs, ld_index_arrays(emp, names, 3), 0.4, 0.5, init_filter_p=0.8, tiethreshold=0.1, err_thresh=1)
    assert eng.stats["test_p"] == want["test_p"]
    assert np.array_equal(got, want["corrected"]), "%d corrected calls differ" % int((got != want["corrected"]).sum())
    assert np.array_equal(eng.stats["corrected_per_animal"], want["corrected_per_animal"])
    assert int((got != obs).sum()) > 10


# ---- entry points end to end and the multi-GPU product path (item 1d / 5) -----------------------
def _write_case(tmp, n=600, gens=6, s=1300, seed=3):
    import os
    from genofix_b200 import synth
    from genofix_b200.gfutils.packedgens import PackedGens
    rows = synth.make_pedigree(n, gens, sire_frac=0.1, dam_frac=0.5, seed=seed)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, s, seed=seed), 0.01, seed=seed), 0.01, seed=seed + 1)
    np.savetxt(os.path.join(tmp, "ped.fam"), rows, fmt="%d")
    names = ["SNP%d" % (i + 1) for i in range(s)]
    with open(os.path.join(tmp, "snps.map"), "w") as f:
        for i, nm in enumerate(names):
            f.write("%d %s %.2f %d\n" % (1 if i < 700 else 2, nm, 0.01 * i, 1000 * i))      # chromosome 2 starts at column 700
    with open(os.path.join(tmp, "geno.ssv"), "w") as f:
        f.write(" ".join(["id"] + names) + "\n")
        for k, r in zip(rows[:, 0], obs):
            f.write(" ".join([str(int(k))] + [str(int(v)) for v in r]) + "\n")
    PackedGens.write(os.path.join(tmp, "geno.g2b"), rows[:, 0], names, obs)
    return rows, obs, names


def _run(cmd, **kw):
    import os
    import subprocess
    import sys
    from conftest import ROOT
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable] + cmd, capture_output=True, text=True, cwd=ROOT, env=env, timeout=900, **kw)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    return r


def _expected_by_chunks(rows, obs, chunk, bounds):
    """Engine.correct on the chunks genofix.py forms: per chromosome [a, b), chunk() of `chunk` SNPs"""
    from genofix_b200.chunking import chunk as chunk_fn
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    want = obs.copy()
    tall = np.zeros(len(rows), dtype=np.int64)
    for a, b in bounds:
        for cols in chunk_fn(list(range(a, b)), chunk):
            want[:, cols] = eng.correct(np.ascontiguousarray(obs[:, cols]), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
            tall += eng.stats["corrected_per_animal"]
    return want, tall


@pytest.mark.timeout(900)
def test_genofix_entry_point_text_and_container_inputs(tmp_path):
    """src/genofix.py end to end on a text matrix and on the 2-bit container (chunks that start inside a word: chromosome 2
    begins at column 700): both outputs equal Engine.correct chunk by chunk, the corrected container equals the text."""
    import pandas as pd
    from genofix_b200.gfutils.packedgens import PackedGens
    tmp = str(tmp_path)
    rows, obs, names = _write_case(tmp)
    want, tall = _expected_by_chunks(rows, obs, 300, [(0, 700), (700, 1300)])
    assert (want != obs).any()
    for inp, out in (("geno.ssv", "out_text"), ("geno.g2b", "out_g2b")):
        _run(["src/genofix.py", "-p", tmp + "/ped.fam", "-s", tmp + "/snps.map", "-o", tmp + "/" + out, "-i", tmp + "/" + inp, "-C", "300"])
        got = pd.read_csv(tmp + "/" + out + "/simulatedgenome_corrected_errors_threshold.ssv", sep=" ", index_col=0)
        assert list(got.columns) == names and [int(x) for x in got.index] == [int(x) for x in rows[:, 0]]
        assert np.array_equal(got.to_numpy(dtype=np.uint8), want), inp
        t = pd.read_csv(tmp + "/" + out + "/corrections_per_animal.tsv", sep="\t", header=None, index_col=0)
        assert np.array_equal(t.iloc[:, 0].to_numpy(), tall)
    g = PackedGens(tmp + "/out_g2b/simulatedgenome_corrected_errors_threshold.g2b")
    assert np.array_equal(g.getMatrix(rows[:, 0]), want)


@pytest.mark.timeout(900)
def test_genofix_entry_point_with_an_ld_index(tmp_path):
    """src/genofix.py -P <index folder>: indexes pickled per chromosome (the layout empirical/preindex.py writes), every
    chromosome corrected as one matrix with the LD blend; equal to Engine.correct_emp per chromosome."""
    import gzip
    import os
    import pickle
    import sys
    import pandas as pd
    from conftest import ROOT
    from genofix_b200.correct_genotypes3 import ld_index_arrays
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    tmp = str(tmp_path)
    rows, obs, names = _write_case(tmp, n=300, gens=5, s=90, seed=9)          # chromosome 2 would start at column 700: one chromosome
    sys.path.insert(0, os.path.join(ROOT, "src"))
    try:
        from empirical.jalleledist3 import JointAllellicDistribution
    finally:
        sys.path.pop(0)
    emp = JointAllellicDistribution(1, names, surround_size=2)
    emp.countJointFrqAll(pd.DataFrame(obs, index=[int(x) for x in rows[:, 0]], columns=names))
    os.makedirs(tmp + "/idx/1")
    with gzip.open(tmp + "/idx/1/empiricalIndex.idx.gz", "wb") as f:
        pickle.dump(emp, f)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    want = eng.correct_emp(obs, ld_index_arrays(emp, names, 1.9), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    tall = eng.stats["corrected_per_animal"].copy()
    assert (want != obs).any()
    _run(["src/genofix.py", "-p", tmp + "/ped.fam", "-s", tmp + "/snps.map", "-o", tmp + "/out", "-i", tmp + "/geno.ssv",
          "-P", tmp + "/idx"])
    got = pd.read_csv(tmp + "/out/simulatedgenome_corrected_errors_threshold.ssv", sep=" ", index_col=0)
    assert np.array_equal(got.to_numpy(dtype=np.uint8), want)
    t = pd.read_csv(tmp + "/out/corrections_per_animal.tsv", sep="\t", header=None, index_col=0)
    assert np.array_equal(t.iloc[:, 0].to_numpy(), tall)


@pytest.mark.timeout(900)
def test_simulate_and_founder_entry_points(tmp_path):
    """src/simulate/simulate.py (host and --device_sim) and src/gfutils/extract_founders.py run end to end."""
    import os
    tmp = str(tmp_path)
    rows, _obs, _names = _write_case(tmp, n=500, s=600)
    for extra, out in (([], "sim_host"), (["--device_sim"], "sim_dev")):
        r = _run(["src/simulate/simulate.py", "-p", tmp + "/ped.fam", "-s", tmp + "/snps.map", "-o", tmp + "/" + out, "-c", "300",
                  "-t", "0.4", "-l", "0.64"] + extra)
        assert os.path.exists(tmp + "/" + out + "/statistics.tsv"), r.stdout[-800:]
    _run(["src/gfutils/extract_founders.py", "-p", tmp + "/ped.fam", "-o", tmp + "/founders.txt"])
    founders = set(int(x) for x in open(tmp + "/founders.txt").read().split())
    assert founders == set(int(k) for k, s, d, _ in rows if s == 0 or d == 0)


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


@pytest.mark.timeout(900)
def test_two_rank_genofix_equals_one_rank(tmp_path):
    """torchrun --nproc-per-node 2 src/genofix.py on the container input: same files as the single-process run (needs 2 GPUs)."""
    import filecmp
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    tmp = str(tmp_path)
    _write_case(tmp)
    _run(["src/genofix.py", "-p", tmp + "/ped.fam", "-s", tmp + "/snps.map", "-o", tmp + "/one", "-i", tmp + "/geno.g2b", "-C", "300"])
    _run(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29611",
          "src/genofix.py", "-p", tmp + "/ped.fam", "-s", tmp + "/snps.map", "-o", tmp + "/two", "-i", tmp + "/geno.g2b", "-C", "300"])
    for f in ("simulatedgenome_corrected_errors_threshold.ssv", "simulatedgenome_corrected_errors_threshold.g2b", "corrections_per_animal.tsv"):
        assert filecmp.cmp(tmp + "/one/" + f, tmp + "/two/" + f, shallow=False), f


_NCCL_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch
from genofix_b200 import parallel, synth
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
rank, world = parallel.init_from_env("nccl")
rows = synth.make_pedigree(700, 6, sire_frac=0.1, dam_frac=0.5, seed=21, related_mating_p=0.2)
obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 900, seed=21), 0.01, seed=21), 0.02, seed=22)
eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
def correct_chunk(cols):
    out = eng.correct(np.ascontiguousarray(obs[:, cols]), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    return out, eng.stats["corrected_per_animal"]
shards, tall = parallel.correct_sharded(obs.shape[1], 128, correct_chunk, obs.shape[0], rank, world, device="cuda")
np.savez(os.path.join(sys.argv[2], "rank%d.npz" % rank), tall=tall.cpu().numpy(),
         **{"c%d" % ci: c for ci, (_, c) in shards.items()}, **{"i%d" % ci: np.array(cols) for ci, (cols, _) in shards.items()})
torch.distributed.destroy_process_group()
"""


@pytest.mark.timeout(900)
def test_two_rank_nccl_sharding_with_the_cuda_engine(tmp_path):
    """parallel.correct_sharded on 2 NCCL ranks with the CUDA engine: the union of the shards equals the single-rank
    result, the all-reduced tallies are identical on both ranks (needs 2 GPUs)."""
    import os
    from conftest import ROOT
    from genofix_b200 import parallel, synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    tmp = str(tmp_path)
    script = os.path.join(tmp, "worker.py")
    with open(script, "w") as f:
        f.write(_NCCL_WORKER)
    _run(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29612",
          script, ROOT, tmp])
    rows = synth.make_pedigree(700, 6, sire_frac=0.1, dam_frac=0.5, seed=21, related_mating_p=0.2)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 900, seed=21), 0.01, seed=21), 0.02, seed=22)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)

    def correct_chunk(cols):
        out = eng.correct(np.ascontiguousarray(obs[:, cols]), 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
        return out, eng.stats["corrected_per_animal"]
    ref_shards, ref_tall = parallel.correct_sharded(obs.shape[1], 128, correct_chunk, obs.shape[0], 0, 1)
    got = np.full_like(obs, 255)
    talls, seen = [], set()
    for r in range(2):
        z = np.load(os.path.join(tmp, "rank%d.npz" % r))
        talls.append(z["tall"])
        for k in z.files:
            if k.startswith("c"):
                ci = int(k[1:])
                assert ci % 2 == r and ci not in seen
                seen.add(ci)
                got[:, z["i%d" % ci]] = z[k]
    ref = np.zeros_like(obs)
    for ci, (cols, c) in ref_shards.items():
        ref[:, cols] = c
    assert seen == set(ref_shards) and np.array_equal(got, ref)
    assert np.array_equal(talls[0], talls[1]) and np.array_equal(talls[0], ref_tall.numpy()) and talls[0].sum() > 0


# ---- parity at the sizes BASELINE.json names ------------------------------------------------------
@pytest.mark.timeout(1500)
def test_config1_example_pedigree_1000_snps_matches_oracle():
    """BASELINE config 1 at its own size: the reference's example pedigree (4 560 animals) x 1 000 SNPs, 1 % errors, one
    1 000-SNP chunk (simulate.py -c 1000), against the oracle port run on all host cores."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    g = load_golden("example_fam")
    drop_rows = synth.complete_rows(g["ped_rows"])
    ids = drop_rows[:, 0]
    obs = synth.inject_errors(synth.gene_drop(drop_rows, 1000, seed=1234), 0.01, seed=1234)
    eng = Engine(PedigreeDAG.from_table(g["ped_rows"]), ids, 2)
    out = eng.correct(obs, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    o = go.correct_matrix_parallel(g["ped_rows"], ids, obs, 0.5, 0.64, back=2, init_filter_p=0.95, tiethreshold=0.4)
    assert _same_f16(eng.raw_host(), o["raw"])
    assert eng.stats["test_p"] == o["test_p"]
    assert np.array_equal(out, o["corrected"]), int((out != o["corrected"]).sum())
    assert np.array_equal(eng.stats["corrected_per_animal"], o["corrected_per_animal"])
    assert np.array_equal(eng.stats["excluded_per_animal"], o["excluded_per_animal"])
    assert (out != obs).sum() > 1000


@pytest.mark.timeout(1500)
def test_config4_named_size_matches_oracle():
    """BASELINE config 4 at the named pedigree size: 15 generations x 1 000 animals, 2 % sires / 30 % dams, related matings,
    10 % of the non-founders with both parents unknown, 5 % missing calls; a 64-SNP chunk against the oracle port."""
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    rows = synth.make_pedigree(15000, 15, sire_frac=0.02, dam_frac=0.3, seed=1234, related_mating_p=0.25, unknown_parents_frac=0.10)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 64, seed=1234), 0.01, seed=1234), 0.05, seed=1235)
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    assert eng.schedule.info["n_loopy"] > 500
    out = eng.correct(obs, 0.5, 0.64, init_filter_p=0.95, tiethreshold=0.4)
    o = go.correct_matrix_parallel(rows, rows[:, 0], obs, 0.5, 0.64, back=2, init_filter_p=0.95, tiethreshold=0.4)
    assert _same_f16(eng.raw_host(), o["raw"])
    assert eng.stats["test_p"] == o["test_p"] or (np.isnan(eng.stats["test_p"]) and np.isnan(o["test_p"]))
    assert np.array_equal(out, o["corrected"]), int((out != o["corrected"]).sum())
    assert np.array_equal(eng.stats["corrected_per_animal"], o["corrected_per_animal"])
    assert np.array_equal(eng.stats["excluded_per_animal"], o["excluded_per_animal"])


@pytest.mark.timeout(1500)
def test_config5_trio_scan_one_million_animals():
    """BASELINE config 5 pedigree (1M animals, 10 generations) x 1 024 SNPs, 2 % missing: k_trio_scan against the oracle's
    scan on 12 000 sampled trios, tested counts included; founders against the oracle's founder list."""
    import torch
    from genofix_b200 import synth
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    n, s = 1000000, 1024
    rows = synth.make_pedigree(n, 10, seed=1234)
    rs = np.random.default_rng(9)
    # the scan does not care whether the data are Mendelian: random calls exercise every (kid, sire, dam) combination
    obs = rs.integers(0, 3, size=(n, s), dtype=np.uint8)
    obs[rs.random((n, s)) < 0.02] = 9
    eng = Engine(PedigreeDAG.from_table(rows), rows[:, 0], 2)
    inc, tested = eng.trio_scan(obs)
    has = np.nonzero((rows[:, 1] > 0) & (rows[:, 2] > 0))[0]
    sel = has[rs.choice(len(has), 12000, replace=False)]
    sub_ids = np.unique(np.concatenate([rows[sel, 0], rows[sel, 1], rows[sel, 2]]))
    sub_rows = sub_ids - 1                                   # ids are 1 .. n in row order
    ri, rt = go.trio_scan(go.OraclePedigree(rows[np.isin(rows[:, 0], rows[sel, 0])]), [int(x) for x in sub_ids], obs[sub_rows])
    pos = {int(k): i for i, k in enumerate(sub_ids)}
    idx = np.array([pos[int(k)] for k in rows[sel, 0]])
    assert np.array_equal(inc[sel], ri[idx]) and np.array_equal(tested[sel], rt[idx])
    assert inc[sel].sum() > 0 and (tested[sel] < s).any()
    nopar = np.nonzero((rows[:, 1] == 0) | (rows[:, 2] == 0))[0]
    assert (inc[nopar] == 0).all() and (tested[nopar] == 0).all()
    f = eng.schedule.founders()
    assert np.array_equal(np.sort(f), np.sort(rows[nopar, 0]))
    del eng
    torch.cuda.empty_cache()


_TMA_WORKER = r"""
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from genofix_b200.engine import Engine
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
z = np.load(sys.argv[2])
eng = Engine(PedigreeDAG.from_table(z["rows"]), z["ids"], 2)
inc, tested = eng.trio_scan(z["obs"])
np.savez(sys.argv[3], inc=inc, tested=tested)
"""


@pytest.mark.timeout(600)
def test_tma_staged_trio_scan_variant_matches_the_default_kernel(tmp_path):
    """GFX_TRIO_TMA=1 selects k_trio_scan_tma (row segments staged by cp.async.bulk + mbarrier, SASS UBLKCP): same tallies as
    the default kernel and as the oracle, on ragged widths (last segment shorter than 512 bytes, last word partial)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    from genofix_b200.engine import Engine
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from oracle import genofix_oracle as go
    for s in (1003, 2048, 5000):
        rows, ids, obs = _random_case(31 + s, n=500, gens=6, s=s, miss=0.03, drop=0.1)
        inp, out = os.path.join(str(tmp_path), "in%d.npz" % s), os.path.join(str(tmp_path), "out%d.npz" % s)
        np.savez(inp, rows=rows, ids=ids, obs=obs)
        script = os.path.join(str(tmp_path), "w.py")
        with open(script, "w") as f:
            f.write(_TMA_WORKER)
        env = dict(os.environ, GFX_TRIO_TMA="1")
        r = subprocess.run([sys.executable, script, ROOT, inp, out], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        z = np.load(out)
        inc, tested = Engine(PedigreeDAG.from_table(rows), ids, 2).trio_scan(obs)
        assert np.array_equal(z["inc"], inc) and np.array_equal(z["tested"], tested)
        ri, rt = go.trio_scan(go.OraclePedigree(rows), [int(x) for x in ids], obs)
        assert np.array_equal(z["inc"], ri) and np.array_equal(z["tested"], rt)
