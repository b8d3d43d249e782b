"""empirical.preindex -- Mendel-filtered LD index (reference: src/empirical/preindex.py:166-346).

Per chunk of kid seeds (kids with both parents genotyped, :166-189) and per chromosome (:197-346) the reference
  1. keeps the candidates whose sire and dam are candidates too and ranks ONLY those animals (:216-253),
  2. transforms the float16 rank matrix, missing calls = 1 (:259-263), sums it per animal (:265-266),
  3. takes the 0.95 / 0.99 / init_filter_p quantiles of all ranks ('interpolated_inverted_cdf') on the FIRST
     chromosome of the chunk and reuses them for the others (:299-302),
  4. counts the joint genotype frequencies of every +-surround window over the cells whose rank is <= the cut
     (:331-338) and pickles the index (:339-346).
The Gaussian-mixture cut of :272-279 only feeds a plot (`filteredIndividualsQuant` is never set, :289-292) and is
not computed.

Here 1-2 are `Engine.mendel_tally` (k_mendel_rank<HIST> + k_rank_rowsum_f16), 3 is an exact weighted quantile over
the 65 537-bin score histogram the rank kernel produced, and 4 is `k_window_counts` through
`JointAllellicDistribution.countJointFrqAll`.  No CPU path: without the CUDA library / a GPU this raises.
"""
import warnings

import numpy as np
import pandas as pd

from ..chunking import chunk
from ..engine import Engine, weighted_nanquantile
from .jalleledist3 import JointAllellicDistribution

SKIP_CHROMOSOMES = ("", "0", "-999", "-9")     # preindex.py:198


def seed_chunks(pedigree, all_ids, seedskidschunk=10000):
    """preindex.py:166-189: [(kid_ids, candidate ids sorted)] per chunk of `seedskidschunk` kids with two genotyped parents."""
    genotyped = set(int(x) for x in all_ids)
    with_parents = []
    for k in (int(x) for x in all_ids):
        sire, dam = pedigree.get_parents(k)
        if sire in genotyped and dam in genotyped:
            with_parents.append(k)
    out = []
    for kid_ids in chunk(with_parents, seedskidschunk, 0, 0):
        cand = set(kid_ids)
        for k in kid_ids:
            cand.update(pedigree.get_parents(k))
        out.append((kid_ids, sorted(cand)))
    return out


def rank_quantiles(hist, table, qs):
    """np.nanquantile(E.flatten(), qs, method='interpolated_inverted_cdf') (:301) from the histogram of score patterns:
    E = table[pattern] for the 65 536 patterns, bin 65536 (missing cells) carries table[0xFFFF]."""
    hist = np.asarray(hist, dtype=np.uint64).astype(np.int64)
    counts = hist[:65536].copy()
    counts[0xFFFF] += hist[65536]
    return weighted_nanquantile(table, counts, qs, 0.0, 1.0)


def index_chunk(genotypes, pedigree, chromosome2snp, surround_size=3, init_filter_p=0.9, back=2):
    """genotypes: DataFrame of ONE seed chunk's candidates (index = animal ids, columns = SNP ids).  Returns
    {chromosome: (JointAllellicDistribution, info)} for every chromosome the reference indexes, info = dict(eval_ids,
    quantQ, quant95, quant99, n_unmasked, individualSumProbs)."""
    cand = [int(x) for x in genotypes.index]
    cset = set(cand)
    eval_ids = []
    for k in cand:                                                                     # :216-223
        sire, dam = pedigree.get_parents(k)
        if sire in cset and dam in cset:
            eval_ids.append(k)
    out = {}
    if not eval_ids:
        return out
    eng = Engine(pedigree, np.asarray(eval_ids, dtype=np.int64), back)
    quants = None
    for chromosome in sorted(chromosome2snp.keys()):                                   # :197-199
        snps = list(chromosome2snp[chromosome])
        if chromosome in SKIP_CHROMOSOMES or len(snps) < 2:
            continue
        sub = genotypes.loc[eval_ids, snps]                                            # :224
        G = np.ascontiguousarray(sub.to_numpy(), dtype=np.uint8)
        G[G > 2] = 9
        E16, sums, _ = eng.mendel_tally(G, None, 1.0)                                   # :230-265
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sums = sums / np.nanmax(sums)                                              # :266
        if quants is None:                                                             # :299-302
            quants = rank_quantiles(eng.tally_hist, eng.tally_table, [0.95, 0.99, init_filter_p])
        quantQ = quants[2]
        mask = np.array(E16 <= quantQ, dtype=bool)                                     # :331
        empC = JointAllellicDistribution(chromosome, snps, chromosome2snp=chromosome2snp, surround_size=surround_size)
        empC.countJointFrqAll(pd.DataFrame(G, index=eval_ids, columns=snps), mask)     # :338
        out[chromosome] = (empC, dict(eval_ids=eval_ids, quantQ=float(quantQ), quant95=float(quants[0]),
                                      quant99=float(quants[1]), n_unmasked=int(np.count_nonzero(mask)),
                                      individualSumProbs=sums))
    return out
