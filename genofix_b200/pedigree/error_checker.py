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
