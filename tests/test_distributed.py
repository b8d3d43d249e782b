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
