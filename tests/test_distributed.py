"""N > 1 host logic on CPU: two gloo ranks share the SNP chunks of one matrix, the per-animal tallies are
all-reduced, and the union of the shards equals the single-process result (the oracle stands in for the
CUDA engine here; the sharding / collective code is the product's)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from genofix_b200 import parallel, synth
    from oracle.genofix_oracle import Oracle
    r, w = parallel.init_from_env("gloo")
    rows = synth.make_pedigree(80, 5, sire_frac=0.3, dam_frac=0.6, seed=5)
    obs = synth.inject_errors(synth.gene_drop(rows, 40, seed=5), 0.03, seed=5)
    o = Oracle(rows, rows[:, 0])

    def correct_chunk(cols):
        res = o.correct_matrix(obs[:, cols], 0.5, 0.64, back=2, init_filter_p=0.9, tiethreshold=0.4)
        return res["corrected"], res["corrected_per_animal"]
    shards, tall = parallel.correct_sharded(obs.shape[1], 8, correct_chunk, obs.shape[0], r, w)
    np.savez(os.path.join(tmp, "rank%d.npz" % r), tall=tall.numpy(),
             **{"c%d" % ci: c for ci, (_, c) in shards.items()}, **{"i%d" % ci: np.array(cols) for ci, (cols, _) in shards.items()})
    import torch.distributed as dist
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    from genofix_b200 import parallel, synth
    from oracle.genofix_oracle import Oracle
    rows = synth.make_pedigree(80, 5, sire_frac=0.3, dam_frac=0.6, seed=5)
    obs = synth.inject_errors(synth.gene_drop(rows, 40, seed=5), 0.03, seed=5)
    o = Oracle(rows, rows[:, 0])

    def correct_chunk(cols):
        res = o.correct_matrix(obs[:, cols], 0.5, 0.64, back=2, init_filter_p=0.9, tiethreshold=0.4)
        return res["corrected"], res["corrected_per_animal"]
    ref_shards, ref_tall = parallel.correct_sharded(obs.shape[1], 8, correct_chunk, obs.shape[0], 0, 1)
    ref = np.zeros_like(obs)
    for ci, (cols, c) in ref_shards.items():
        ref[:, cols] = c
    got = np.full_like(obs, 255)
    talls = []
    seen = set()
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        talls.append(z["tall"])
        for k in z.files:
            if k.startswith("c"):
                ci = int(k[1:])
                assert ci % 2 == r          # round-robin ownership
                assert ci not in seen
                seen.add(ci)
                got[:, z["i%d" % ci]] = z[k]
    assert seen == set(ref_shards)
    assert np.array_equal(got, ref)
    assert np.array_equal(talls[0], talls[1])            # all-reduced: identical on every rank
    assert np.array_equal(talls[0], ref_tall.numpy())
    assert talls[0].sum() > 0


def test_chunks_for_rank_partition():
    from genofix_b200.chunking import chunks_for_rank
    for n in (1, 5, 8, 60):
        for w in (1, 2, 4, 8):
            owned = [chunks_for_rank(n, r, w) for r in range(w)]
            flat = sorted(x for o in owned for x in o)
            assert flat == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def _tally_inputs():
    from genofix_b200 import synth
    from oracle.genofix_oracle import Oracle
    rows = synth.make_pedigree(90, 5, sire_frac=0.3, dam_frac=0.6, seed=12)
    obs = synth.inject_missing(synth.inject_errors(synth.gene_drop(rows, 44, seed=6), 0.05, seed=7), 0.04, seed=8)
    raw = Oracle(rows, rows[:, 0]).mendel_rank(obs, 2)          # stands in for k_mendel_rank
    chunks = [list(range(i, min(i + 8, 44))) for i in range(0, 44, 8)]
    return obs, raw, chunks


def _tally_callbacks(obs, raw, chunks):
    from genofix_b200.engine import rank_transform_table

    def hist_of_chunk(ci):
        cols = chunks[ci]
        pat = np.where(obs[:, cols] == 9, 65536, raw[:, cols].view(np.uint16).astype(np.int64))
        return np.bincount(pat.ravel(), minlength=65537)

    def table_from_hist(hist):
        return rank_transform_table(hist.astype(np.uint64), 0.9)[0]     # float64 E of every score pattern

    def rowsum_of_chunk(ci, table):
        cols = chunks[ci]
        E = table[raw[:, cols].view(np.uint16)]
        E[obs[:, cols] == 9] = -1.0                                     # error_checker.py:196-197
        return E.sum(axis=1)
    return hist_of_chunk, rowsum_of_chunk, table_from_hist


def _tally_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from genofix_b200 import parallel
    r, w = parallel.init_from_env("gloo")
    obs, raw, chunks = _tally_inputs()
    h, rs, tf = _tally_callbacks(obs, raw, chunks)
    hist, sums = parallel.mendel_tally_sharded(len(chunks), h, rs, tf, obs.shape[0], r, w)
    np.savez(os.path.join(tmp, "tally%d.npz" % r), hist=hist, sums=sums)
    import torch.distributed as dist
    dist.destroy_process_group()


def test_two_rank_mendel_tally_matches_single_process(tmp_path):
    """The error_checker tally over SNP shards: histogram all-reduce -> one transform table, row-sum all-reduce."""
    port = _free_port()
    mp.spawn(_tally_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    from genofix_b200 import parallel
    from oracle import genofix_oracle as go
    obs, raw, chunks = _tally_inputs()
    h, rs, tf = _tally_callbacks(obs, raw, chunks)
    hist1, sums1 = parallel.mendel_tally_sharded(len(chunks), h, rs, tf, obs.shape[0], 0, 1)
    z = [np.load(os.path.join(str(tmp_path), "tally%d.npz" % r)) for r in range(2)]
    assert np.array_equal(z[0]["hist"], z[1]["hist"]) and np.array_equal(z[0]["hist"], hist1)
    assert np.array_equal(z[0]["sums"], z[1]["sums"])                    # all-reduced: identical on every rank
    assert np.allclose(z[0]["sums"], sums1, rtol=1e-12, atol=0)
    # against the whole-matrix float64 transform (what error_checker.py computes, in float64 instead of float16)
    with np.errstate(all="ignore"):
        E = np.log(np.log(raw.astype(np.float64) + 1) + 1)
        E = E / np.nanmax(E)
    E[obs == 9] = -1.0
    assert np.allclose(sums1, E.sum(axis=1), rtol=1e-12, atol=0)
    # and it ranks the animals like the reference's float16 tally does, up to float16 resolution
    half = go.half_rowsum(go.half_transform(raw, obs, -1.0)).astype(np.float64)
    assert np.max(np.abs(half - sums1)) < 0.05 * max(1.0, np.max(np.abs(sums1)))
