"""The oracle restatement against outputs of the UNMODIFIED reference (tests/golden/*.npz, made by
oracle/make_golden.py in the build container).  Bit-exact: rank matrix, threshold, corrected calls,
error counts and every posterior."""
import numpy as np
import pytest

from conftest import load_golden
from oracle.genofix_oracle import Oracle
from oracle import genofix_oracle as go

CASES = ["toy_a", "toy_b", "deep", "nan_loop", "example_fam"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    g = load_golden(name)
    tp, ts, back, q, tie, et = g["params"]
    o = Oracle(g["ped_rows"], g["ids"])
    r = o.correct_matrix(g["obs"], tp, ts, back=int(back), init_filter_p=q, tiethreshold=tie, err_thresh=et,
                         keep_posteriors=True)
    assert np.array_equal(r["raw"].view(np.uint16), g["raw"].view(np.uint16))
    assert r["test_p"] == g["test_p"] or (np.isnan(r["test_p"]) and np.isnan(g["test_p"]))
    assert np.array_equal(r["corrected"], g["corrected"])
    assert int(r["corrected_per_animal"].sum()) == int(g["errors_all"])
    occ, mine = {}, {}
    for (gen, s, d) in sorted(r["pair_posteriors"].keys()):
        k = occ.get((s, d), 0)
        occ[(s, d)] = k + 1
        for j, (ps, pd_, _a, _b) in r["pair_posteriors"][(gen, s, d)].items():
            mine[(k, s, d, j)] = np.concatenate([ps, pd_])
    ref = {(int(x[0]), int(x[1]), int(x[2]), int(x[3])): x[4:] for x in g["pair_post"]}
    assert set(mine) == set(ref)
    for k, v in ref.items():
        assert np.array_equal(mine[k], v, equal_nan=True)
    smine = {(k, j): v[0] for k, res in r["single_posteriors"].items() for j, v in res.items()}
    sref = {(int(x[0]), int(x[1])): x[2:] for x in g["single_post"]}
    assert set(smine) == set(sref)
    for k, v in sref.items():
        assert np.array_equal(smine[k], v, equal_nan=True)


def test_nan_case_is_really_nan():
    g = load_golden("nan_loop")
    assert np.isnan(g["raw"].astype(float)).sum() == 3
    assert np.isnan(g["test_p"])
    assert np.array_equal(g["corrected"], g["obs"])


def test_kats_model_functions():
    from oracle import genofix_oracle as go
    g = load_golden("kats")
    for c in range(len(g["kid"])):
        n = int((g["kid"][c] >= 0).sum())
        ks = list(g["kid"][c, :n])
        ps = [None if p < 0 else int(p) for p in g["par"][c, :n]]
        assert np.array_equal(np.asarray(go.probs_kids(ks, ps), dtype=float), g["probs_kids"][c])
    # commented examples inside the reference (bayesnet/model.py:91,136-142)
    assert np.allclose(go.probs_kids([1, 0, 1, 0, 1], [2, 2, 2, 2, 2]), [2 / 3, 1 / 3, 0])
    assert go.probs_differences_parent(2, 2, 1) == 1.0
    assert go.probs_differences_parent(1, 1, 0) == 0.25
    assert go.probs_differences_parent(2, 9, 0) == 0.5
    assert go.probs_differences_parent(9, 2, 0) == 0.5
    assert go.probs_differences_parent(9, 9, 1) == 0
    for i, p1 in enumerate([0, 1, 2, 9]):
        for j, p2 in enumerate([0, 1, 2, 9]):
            for o in range(3):
                assert go.probs_differences_parent(p1, p2, o) == g["diff_parent"][i, j, o]


def test_kats_empirical_counts_and_windows():
    from oracle import genofix_oracle as go
    g = load_golden("kats")
    for w in (3, 5, 7):
        res = go.count_joint_frq(g["cj_tab_%d" % w], g["cj_mask_%d" % w], w, 1e-4)
        import itertools
        vals = np.array([res[v] for v in itertools.product([0, 1, 2], repeat=w)])
        assert np.array_equal(vals, g["cj_val_%d" % w])
    for n, sur, i, a, b in g["windows"]:
        assert go.ld_window(int(n), int(i), int(sur)) == (int(a), int(b))


ERRCHK = ["errchk_a", "errchk_b", "errchk_long", "errchk_example"]


@pytest.mark.parametrize("name", ERRCHK)
def test_oracle_error_checker_matches_reference(name):
    """pedigree/error_checker.py main() run unmodified (oracle/ref_harness.run_error_checker): the selected kids, the
    float16 probs_errors.csv matrix and individual_sum_probs_errors.csv, bit for bit."""
    from oracle import genofix_oracle as go
    g = load_golden(name)
    wp, _cand, E, sums = go.error_checker(g["ped_rows"], g["ids"], g["obs"], g["kids"])
    assert np.array_equal(np.array(wp, dtype=np.int64), g["with_parents"])
    assert np.array_equal(E.view(np.uint16), g["probs_errors"])
    assert np.array_equal(sums.view(np.uint16), g["sums"])


def test_error_checker_golden_pins_the_float16_accumulator():
    """errchk_long is long enough that the float16 accumulator loses low bits: a float64 sum rounded at the end does
    not reproduce the reference's file (35 of 36 animals differ), the column-by-column float16 chain does."""
    from oracle import genofix_oracle as go
    g = load_golden("errchk_long")
    E = g["probs_errors"].view(np.float16)
    ref = g["sums"]
    wide = E.astype(np.float64).sum(axis=1).astype(np.float16)
    chain = go.half_rowsum(E)
    with np.errstate(all="ignore"):
        wide = wide / np.nanmax(wide)
        chain = chain / np.nanmax(chain)
    assert (wide.view(np.uint16) != ref).sum() > 30
    assert np.array_equal(chain.view(np.uint16), ref)


def _chrom_cols(g):
    cc = {}
    for i, c in enumerate(g["chrom"]):
        cc.setdefault(str(c), []).append(i)
    return cc


@pytest.mark.parametrize("name", ["preidx_a", "preidx_b", "preidx_c"])
def test_oracle_preindex_matches_reference(name):
    """empirical/preindex.py main() run unmodified (oracle/ref_harness.run_preindex): the rank cut it printed and the
    number of cells under the cut for every (seed chunk, chromosome), and every window frequency of the pickled
    indexes (the last chunk's: each chunk overwrites the file, preindex.py:339-346), bit for bit."""
    from oracle import genofix_oracle as go
    g = load_golden(name)
    ped = go.OraclePedigree(g["ped_rows"])
    chunks = go.preindex_chunks(ped, g["ids"], int(g["seed_chunk"]))
    assert len(chunks) == (4 if name == "preidx_c" else 1)
    k = 0
    for ci, (_kid_ids, cand) in enumerate(chunks):
        res = go.preindex_chunk(ped, g["ids"], g["obs"], cand, _chrom_cols(g), int(g["surround"]), float(g["q"]))
        assert sorted(res) == [str(c) for c in g["chroms"]]
        for c in sorted(res):
            r = res[c]
            assert r["quantQ"] == g["quantQ"][k]
            assert (r["n_unmasked"], len(r["eval_ids"]) * len(_chrom_cols(g)[c])) == tuple(g["kept"][k])
            k += 1
            if ci == len(chunks) - 1:
                for i, f in enumerate(r["freq"]):
                    w = int(g["wlen_" + c][i])
                    assert len(f) == (3 ** w if w else 0)
                    assert np.array_equal(f, g["freq_" + c][i][:len(f)])
    assert k == len(g["quantQ"])


def test_parallel_oracle_driver_equals_the_single_process_oracle():
    """correct_matrix_parallel (columns dealt to processes, the matrix-wide maximum / quantile taken once in between) is
    the checker of the BASELINE-size GPU tests: it must return exactly what Oracle.correct_matrix returns."""
    from genofix_b200 import synth
    from oracle import genofix_oracle as go
    rows = synth.make_pedigree(160, 6, sire_frac=0.2, dam_frac=0.5, seed=3, related_mating_p=0.3, unknown_parents_frac=0.05)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 30, seed=3), 0.03, seed=3), 0.02, seed=4)
    a = go.Oracle(rows, rows[:, 0]).correct_matrix(obs, 0.5, 0.64, back=2, init_filter_p=0.9, tiethreshold=0.4)
    b = go.correct_matrix_parallel(rows, rows[:, 0], obs, 0.5, 0.64, back=2, init_filter_p=0.9, tiethreshold=0.4, procs=3)
    assert np.array_equal(a["corrected"], b["corrected"]) and (a["corrected"] != obs).any()
    assert a["test_p"] == b["test_p"]
    assert np.array_equal(a["raw"].view(np.uint16), b["raw"].view(np.uint16))
    assert np.array_equal(a["corrected_per_animal"], b["corrected_per_animal"])
    assert np.array_equal(a["excluded_per_animal"], b["excluded_per_animal"])


@pytest.mark.parametrize("name", ["emp_a", "emp_b"])
def test_oracle_empirical_blend_equals_reference(name):
    """correctMatrix WITH an LD index (correct_genotypes3.py:619-667 ranks, :746-779 / :883-908 decisions): the oracle's index,
    every getEmpProbs result, the threshold the reference printed, the corrected calls and the error count."""
    g = load_golden(name)
    obs = g["obs"]
    emp = go.emp_index(obs, int(g["emp_surround"]), float(g["emp_pseudocount"]), float(g["emp_weight"]))
    assert np.array_equal(emp["wstart"], g["emp_wstart"]) and np.array_equal(emp["wlen"], g["emp_wlen"])
    for j in range(obs.shape[1]):
        assert np.array_equal(emp["freq"][j], g["emp_freq"][j, :3 ** int(emp["wlen"][j])])
    for r, j, o, a, b, c in g["emp_probs"]:
        assert np.array_equal(go.emp_state_probs(obs[int(r)], int(j), emp), [a, b, c])
    P = g["params"]
    res = go.Oracle(g["ped_rows"], g["ids"]).correct_matrix(obs, P[0], P[1], back=int(P[2]), init_filter_p=P[3], tiethreshold=P[4],
                                                            err_thresh=P[5], emp=emp)
    assert res["test_p"] == float(g["test_p"])
    assert np.array_equal(res["corrected"], g["corrected"])
    assert int(res["corrected_per_animal"].sum()) == int(g["errors_all"])
    assert int((res["corrected"] != obs).sum()) > 0
