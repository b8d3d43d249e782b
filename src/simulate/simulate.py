#!/usr/bin/env python3
"""simulate -- gene-drop a genome down a pedigree, inject errors, correct them on the GPU and score the result.

Flag set of the reference entry point (src/simulate/simulate.py:89-113) with its two argparse collisions
resolved: `-T/--threads` once; `-M` stays `--minimum_cluster_size`, the boolean microarray switch
is `--microarray_errors` only.  The reference imports correct_genotypes2, which cannot run at HEAD (D2); this
entry point drives correct_genotypes3 with empC=None.  The gene drop uses independent SNPs (free
recombination): with the empirical LD model off, linkage does not enter the correction.
"""
import math
import multiprocessing
import pathlib
import sys
import time
from argparse import ArgumentParser

import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # src/ (the reference sets PYTHONPATH=src)
import _bootstrap  # noqa: F401,E402
import numpy as np
import pandas as pd

from genofix_b200 import synth
from genofix_b200.chunking import chunk
from genofix_b200.correct_genotypes3 import CorrectGenotypes
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG


def build_parser():
    p = ArgumentParser(description="simulate errors on a pedigree and correct them (CUDA)")
    p.add_argument("-H", "--founders_haplo", dest="founders_haplo_file", required=False)
    p.add_argument("-G", "--founders_geno", dest="founders_geno_file", required=False)
    p.add_argument("-p", "--pedigree", dest="pedigree", required=True)
    p.add_argument("-s", "--snps", dest="snps", required=True)
    p.add_argument("-c", "--chunk", dest="chunk", default=1000, type=int)
    p.add_argument("-l", "--thresholdsingles", dest="threshold_singles", type=float, default=0.7)
    p.add_argument("-m", "--thresholdpairs", dest="threshold_pairs", type=float, default=0.5)
    p.add_argument("-w", "--surroundsnps", dest="surround_size", type=int, default=4)
    p.add_argument("-o", "--outdir", dest="outdir", required=True)
    p.add_argument("-d", "--lddisttype", dest="lddisttype", default="global", choices=["none", "global", "local"])
    p.add_argument("-b", "--lookback", dest="lookback", type=int, default=2)
    p.add_argument("-t", "--tiethreshold", dest="tiethreshold", type=float, default=0.05)
    p.add_argument("-e", "--errorrate", dest="error_rate", type=float, default=1)
    p.add_argument("-et", "--error_thresh", dest="err_thresh", type=float, default=1)
    p.add_argument("-g", "--popgenome", dest="prior_pop_genome", type=str, required=False)
    p.add_argument("-z", "--popgenomeerrors", dest="prior_genome_errors", type=str, required=False)
    p.add_argument("-n", "--firstnsnps", dest="first_n_snps", type=int, default=None)
    p.add_argument("-q", "--initquantilefilter", dest="initquantilefilter", type=float, default=0.9)
    p.add_argument("-Q", "--filter_e", dest="filter_e", type=float, default=0.8)
    p.add_argument("-W", "--weightempirical", dest="weight_empirical", type=float, default=2)
    p.add_argument("-M", "--minimum_cluster_size", dest="minimum_cluster_size", type=float, required=False)
    p.add_argument("-E", "--elimination_order", dest="elimination_order", type=str, default="weightedminfill",
                   choices=["weightedminfill", "minneighbors", "minweight", "minfill"])
    p.add_argument("-T", "--threads", dest="threads", type=int, default=multiprocessing.cpu_count())
    p.add_argument("--microarray_errors", dest="microarray_errors", action="store_true", default=False)
    p.add_argument("--device_sim", dest="device_sim", action="store_true", default=False,
                   help="gene drop, error injection, correction and scoring on the GPU (2-bit packed, nothing but the "
                        "statistics comes back); the chunk size is rounded to a multiple of 16 SNPs")
    return p


def _read_matrix(path):
    return pd.read_csv(path, sep=" ", compression="gzip" if path.endswith(".gz") else None, header=0, index_col=0)


def drop_rows(pedigree):
    """(kid, sire, dam, sex) rows, parents first, founders included, for synth.gene_drop."""
    order = np.argsort(pedigree.generation, kind="stable")
    # generation (min line length) is not a topological order in general; do a proper one
    n = len(pedigree.ids)
    depth = np.zeros(n, dtype=np.int64)
    for _ in range(n):
        ds = np.where(pedigree.sire_idx >= 0, depth[np.maximum(pedigree.sire_idx, 0)] + 1, 0)
        dd = np.where(pedigree.dam_idx >= 0, depth[np.maximum(pedigree.dam_idx, 0)] + 1, 0)
        new = np.maximum(ds, dd)
        if np.array_equal(new, depth):
            break
        depth = new
    order = np.argsort(depth, kind="stable")
    ids = pedigree.ids
    rows = np.zeros((n, 4), dtype=np.int64)
    rows[:, 0] = ids[order]
    rows[:, 1] = np.where(pedigree.sire_idx[order] >= 0, ids[np.maximum(pedigree.sire_idx[order], 0)], 0)
    rows[:, 2] = np.where(pedigree.dam_idx[order] >= 0, ids[np.maximum(pedigree.dam_idx[order], 0)], 0)
    rows[:, 3] = pedigree.sex[order]
    return rows


def run_on_device(args, pedigree, snps, out_dir, genomein=None):
    """The whole loop on the device: simulate.py:187-340 (drop + errors), the correction, :426-515 (scoring).  With the map's
    chromosome and cM columns the drop follows quickgsim's meiosis (crossovers per chromosome), else free recombination."""
    import torch
    from genofix_b200.engine import Engine
    from genofix_b200.simulate_device import DeviceSimulator
    rows = drop_rows(pedigree)
    n, s = len(rows), len(snps)
    sim = DeviceSimulator(rows, s, seed=1234)
    linkage = None
    if genomein is not None:
        chrom, cm = np.asarray(genomein["chrom"])[:s], np.asarray(genomein["cm"], dtype=float)[:s]
        grouped = len(set(chrom[[0] + [i for i in range(1, s) if chrom[i] != chrom[i - 1]]])) == 1 + int(np.count_nonzero(chrom[1:] != chrom[:-1]))
        if grouped and np.all(np.isfinite(cm)) and cm.max() > 0:
            linkage = (chrom, cm)
        else:
            print("genetic map not usable (SNPs not grouped by chromosome or no cM positions): free recombination")
    truth = sim.gene_drop(linkage_map=linkage)
    obs = sim.inject_errors(truth, args.error_rate / 100)
    fixed = obs.clone()
    eng = Engine(pedigree, rows[:, 0], args.lookback)
    step = max(16, (args.chunk // 16) * 16)
    tic = time.time()
    for a in range(0, s, step):
        b = min(s, a + step)
        eng._alloc(b - a)
        nw = (b - a + 15) // 16
        blk = torch.full((n, eng.wpr), -1, dtype=torch.int32, device=obs.device)
        blk[:, :nw] = obs[:, a // 16:a // 16 + nw]
        eng.use_packed(blk, b - a)
        eng.correct_resident(args.threshold_pairs, args.threshold_singles, init_filter_p=args.initquantilefilter,
                             tiethreshold=args.tiethreshold, err_thresh=args.err_thresh)
        fixed[:, a // 16:a // 16 + nw] = eng.packed_live[:, :nw]
    eng.check_status()
    secs = time.time() - tic
    stats = sim.confusion(truth, obs, fixed)
    stats.update(animals=n, snps=s, seconds=secs, threshold_singles=args.threshold_singles, threshold_pairs=args.threshold_pairs,
                 tiethreshold=args.tiethreshold, lookback=args.lookback, initquantilefilter=args.initquantilefilter)
    pd.DataFrame([stats]).to_csv("%s/statistics.tsv" % out_dir, sep="\t", index=False)
    print(stats)
    return 0


def main(argv=None):
    args = build_parser().parse_args(argv)
    out_dir = args.outdir
    pathlib.Path(out_dir).mkdir(parents=True, exist_ok=True)
    pedigree = PedigreeDAG.from_file(args.pedigree)
    genomein = pd.read_csv(args.snps, sep=" ", names=["chrom", "snpid", "cm", "pos"], engine="c")
    snps = [str(x) for x in genomein["snpid"]]
    if args.first_n_snps is not None:
        snps = snps[:args.first_n_snps]
    if args.device_sim:
        return run_on_device(args, pedigree, snps, out_dir, genomein)
    if args.founders_haplo_file or args.founders_geno_file:
        raise NotImplementedError("founder genotype import is part of the simulator, outside this path")
    if args.prior_pop_genome is None:
        rows = drop_rows(pedigree)
        truth = synth.gene_drop(rows, len(snps), seed=1234)
        genomematrix = pd.DataFrame(truth, index=[int(x) for x in rows[:, 0]], columns=snps)
        genomematrix.to_csv("%s/simulated_genotype_genome.ssv.gz" % out_dir, sep=" ", compression="gzip")
    else:
        genomematrix = _read_matrix(args.prior_pop_genome).iloc[:, :len(snps)].astype(np.uint8)
    if args.prior_genome_errors is None:
        g = genomematrix.to_numpy(dtype=np.uint8)
        if args.microarray_errors:
            rs = np.random.Generator(np.random.PCG64(1234))
            k = int(math.ceil(g.size * (args.error_rate / 100)))
            r, c = rs.integers(0, g.shape[0], k), rs.integers(0, g.shape[1], k)
            e = g.copy()
            e[r, c] = np.where(g[r, c] == 1, rs.choice([0, 2], size=k), 1)   # hom <-> het switches
        else:
            e = synth.inject_errors(g, args.error_rate / 100, seed=1234)
        with_errors = pd.DataFrame(e, index=genomematrix.index, columns=genomematrix.columns)
        with_errors.to_csv("%s/simulatedgenome_with_%serrors.ssv.gz" % (out_dir, args.error_rate), sep=" ", compression="gzip")
    else:
        with_errors = _read_matrix(args.prior_genome_errors).iloc[:, :len(snps)].astype(np.uint8)
    n_err = int((genomematrix.to_numpy() != with_errors.to_numpy()).sum())
    print("%s errors %1.2fpc in array" % (n_err, 100.0 * n_err / genomematrix.size))
    c = CorrectGenotypes(elimination_order=args.elimination_order)
    corrected = with_errors.copy()
    tic = time.time()
    cols = list(with_errors.columns)
    for part in chunk(cols, args.chunk):
        res = c.correctMatrix(with_errors.loc[:, part], pedigree, None, args.threshold_pairs, args.threshold_singles,
                              back=args.lookback, tiethreshold=args.tiethreshold, init_filter_p=args.initquantilefilter,
                              filter_e=args.filter_e, weight_empirical=args.weight_empirical, threads=args.threads,
                              err_thresh=args.err_thresh, DEBUGDIR=out_dir)
        corrected.loc[:, part] = res.to_numpy()
    secs = time.time() - tic
    print("done correct matrix in %s minutes" % (secs / 60))
    corrected.to_csv("%s/simulatedgenome_corrected_errors.ssv.gz" % out_dir, sep=" ", compression="gzip")
    # evaluation (reference: simulate.py:426-515): confusion of error cells vs changed cells
    t, o, k = genomematrix.to_numpy(), with_errors.to_numpy(), corrected.to_numpy()
    was_err, changed = t != o, k != o
    tp = int((was_err & changed & (k == t)).sum())            # error fixed to the truth
    tp_missing = int((was_err & changed & (k == 9)).sum())    # error blanked
    wrong_fix = int((was_err & changed & (k != t) & (k != 9)).sum())
    fp = int((~was_err & changed).sum())                      # good call altered
    fn = int((was_err & ~changed).sum())
    flagged = tp + tp_missing + wrong_fix
    precision = flagged / max(1, flagged + fp)
    recall = flagged / max(1, flagged + fn)
    stats = dict(animals=t.shape[0], snps=t.shape[1], errors=int(was_err.sum()), fixed=tp, blanked=tp_missing,
                 wrong_fix=wrong_fix, new_errors=fp, missed=fn, precision=precision, recall=recall,
                 f1=2 * precision * recall / max(1e-12, precision + recall), seconds=secs,
                 threshold_singles=args.threshold_singles, threshold_pairs=args.threshold_pairs,
                 tiethreshold=args.tiethreshold, lookback=args.lookback, initquantilefilter=args.initquantilefilter)
    pd.DataFrame([stats]).to_csv("%s/statistics.tsv" % out_dir, sep="\t", index=False)
    print(stats)
    return 0


if __name__ == "__main__":
    sys.exit(main())
