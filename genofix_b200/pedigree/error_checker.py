"""pedigree.error_checker -- per-animal Mendel error tallies (reference: src/pedigree/error_checker.py).

The reference selects the kids whose sire and dam are both genotyped (:138-144), restricts the genotype matrix to
those kids and their parents (:146-157), runs mendelProbsSingle(kid, back=2) on every one of them (:164-187),
transforms the FLOAT16 rank matrix with log(log(x + 1) + 1) / max (:193-195), sets missing calls to -1 (:196-197),
writes the rows of the selected kids (:199) and their row sums divided by the largest sum (:201-204).

Here the ranks come from `k_mendel_rank<HIST>` (one pass, fused with the score histogram), the float16 transform is a
65 536-entry table built with numpy from that histogram, and the row sums are taken by `k_rank_rowsum_f16`, which
reproduces pandas' float16 accumulation order so that the result is bit-identical to the reference's.
There is no CPU path: without the CUDA library / a GPU this raises.
"""
import warnings

import numpy as np
import pandas as pd

from ..engine import Engine


def trio_candidates(pedigree, all_ids, kids):
    """error_checker.py:138-153.  Returns (animalswithparents in genotype-file order, candidate ids sorted).
    The reference keeps the candidates in a set; their order does not influence any result."""
    genotyped = set(int(x) for x in all_ids)
    wanted = set(int(x) for x in kids)
    with_parents = []
    cand = set()
    for k in (int(x) for x in all_ids):
        if k not in wanted:
            continue
        sire, dam = pedigree.get_parents(k)
        if sire in genotyped and dam in genotyped:
            with_parents.append(k)
            cand.update((k, sire, dam))
    return with_parents, sorted(cand)


def mendel_error_tally(genotypes, pedigree, kids, missing_value=-1.0, back=2):
    """genotypes: DataFrame (index = animal ids, columns = SNP ids, values 0/1/2/9), already restricted to the
    chromosome to test.  Returns (probs_errors, individualSumProbs): the float16 DataFrame the reference writes to
    probs_errors.csv (rows = kids with both parents genotyped) and the float16 vector of
    individual_sum_probs_errors.csv (error_checker.py:199-204).  missing_value = 1 gives preindex.py:262-263."""
    all_ids = [int(x) for x in genotypes.index]
    with_parents, cand = trio_candidates(pedigree, all_ids, kids)
    if len(cand) == 0:
        return (pd.DataFrame(np.zeros((0, genotypes.shape[1]), dtype=np.float16), columns=genotypes.columns),
                np.zeros(0, dtype=np.float16))
    Gc = np.ascontiguousarray(genotypes.loc[cand, :].to_numpy(), dtype=np.uint8)
    Gc[Gc > 2] = 9
    eng = Engine(pedigree, np.asarray(cand, dtype=np.int64), back)
    pos = {k: i for i, k in enumerate(cand)}
    rows = np.array([pos[k] for k in with_parents], dtype=np.int32)
    E16, sums, _ = eng.mendel_tally(Gc, rows, missing_value)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sums = sums / np.nanmax(sums) if len(sums) else sums          # :202 (float16 / float16)
    return pd.DataFrame(E16, index=with_parents, columns=genotypes.columns), sums


def mendel_error_tally_sharded(genotypes, pedigree, kids, chunk_snps=10000, missing_value=-1.0, back=2, rank=0, world=1):
    """The same tally with the SNP chunks of the matrix dealt round-robin to `world` ranks (one process per GPU,
    `parallel.init_from_env()` first): histogram all-reduce -> one transform table on every rank, float64 row-sum
    all-reduce (`parallel.mendel_tally_sharded`).  Returns (animalswithparents, individualSumProbs float64), identical
    on every rank.  float64 sums: the float16 accumulation chain of `mendel_error_tally` cannot be split over ranks.
    NOTE: composed of GPU-tested calls (gfx_mendel_rank_hist, gfx_rank_rowsum); the host logic is covered by the gloo test
    in tests/test_distributed.py; this composition itself has not been run on a multi-GPU box yet."""
    from .. import parallel
    from ..engine import rank_transform_table
    all_ids = [int(x) for x in genotypes.index]
    with_parents, cand = trio_candidates(pedigree, all_ids, kids)
    if not cand:
        return with_parents, np.zeros(0)
    Gc = np.ascontiguousarray(genotypes.loc[cand, :].to_numpy(), dtype=np.uint8)
    Gc[Gc > 2] = 9
    n_snps = Gc.shape[1]
    chunks = [(a, min(a + chunk_snps, n_snps)) for a in range(0, n_snps, max(1, chunk_snps))]
    eng = Engine(pedigree, np.asarray(cand, dtype=np.int64), back)
    torch = eng.torch

    def rank_chunk(ci):
        a, b = chunks[ci]
        eng.load_host(Gc[:, a:b])
        eng.mendel_rank_device(fused=True)

    def hist_of_chunk(ci):
        rank_chunk(ci)
        eng.host_hist.copy_(eng.hist, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return eng.host_hist.numpy().view(np.uint64).copy()

    def table_from_hist(hist):
        return rank_transform_table(hist.astype(np.uint64), 0.5)[0]

    def rowsum_of_chunk(ci, table):
        rank_chunk(ci)                                  # 93 us per 10k x 10k chunk: recomputed rather than kept
        eng.host_elut.numpy()[:] = table
        eng.elut.copy_(eng.host_elut, non_blocking=True)
        return eng.rank_rowsum_device(missing_value).cpu().numpy()

    _hist, sums = parallel.mendel_tally_sharded(len(chunks), hist_of_chunk, rowsum_of_chunk, table_from_hist, len(cand),
                                                rank, world, device="cuda")
    pos = {k: i for i, k in enumerate(cand)}
    sel = sums[[pos[k] for k in with_parents]]
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return with_parents, (sel / np.nanmax(sel) if len(sel) else sel)
