#!/usr/bin/env python3
"""genofix -- correct genotype calls with Mendelian likelihoods on a B200.

Same command line as the reference entry point (src/genofix.py:94-112).  Differences, all forced by the
state of the reference at HEAD:
  * -P/--empC is optional and ignored with a warning: the empirical (LD) blend cannot run in the reference
    (D3/D4), so correction uses the Mendelian model only (the reference's `empC=None` path);
  * every chunk is corrected from its own columns and written back by column (the reference passes the whole
    chromosome and slices rows, D5);
  * chunks are dealt round-robin to the GPUs of the box when launched under torchrun.
"""
import multiprocessing
import pathlib
import sys
import time
from argparse import ArgumentParser
from collections import Counter, defaultdict

import _bootstrap  # noqa: F401
import numpy as np
import pandas as pd

from genofix_b200 import parallel
from genofix_b200.chunking import chunk, chunks_for_rank
from genofix_b200.correct_genotypes3 import CorrectGenotypes
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
from gfutils.gensrandaccess import GensCache


def build_parser():
    p = ArgumentParser(description="genofix: Mendelian genotype correction (CUDA)")
    p.add_argument("-p", "--pedigree", dest="pedigree", required=True, help="pedigree file")
    p.add_argument("-s", "--snps", dest="snps", required=True, help="snp map file")
    p.add_argument("-l", "--thresholdsingles", dest="threshold_singles", type=float, default=0.64)
    p.add_argument("-m", "--thresholdpairs", dest="threshold_pairs", type=float, default=0.5)
    p.add_argument("-o", "--outdir", dest="outdir", required=True, help="outputdir")
    p.add_argument("-d", "--lddisttype", dest="lddisttype", default="global", choices=["none", "global", "local"])
    p.add_argument("-b", "--lookback", dest="lookback", type=int, default=2)
    p.add_argument("-t", "--tiethreshold", dest="tiethreshold", type=float, default=0.4)
    p.add_argument("-i", "--genotypes_input_file", dest="genotypes_input_file", type=str, required=True)
    p.add_argument("-n", "--firstnsnps", dest="first_n_snps", type=int, default=None)
    p.add_argument("-q", "--initquantilefilter", dest="initquantilefilter", type=float, default=0.95)
    p.add_argument("-Q", "--filter_e", dest="filter_e", type=float, default=0.99)
    p.add_argument("-W", "--weightempirical", dest="weight_empirical", type=float, default=1.9)
    p.add_argument("-E", "--elimination_order", dest="elimination_order", type=str, default="weightedminfill",
                   choices=["weightedminfill", "minneighbors", "minweight", "minfill"])
    p.add_argument("-T", "--threads", dest="threads", type=int, default=multiprocessing.cpu_count())
    p.add_argument("-P", "--empC", dest="empC", required=False, default=None,
                   help="prior empirical distribution folder (ignored: see module docstring)")
    p.add_argument("-C", "--chunksize", dest="chunksize", type=int, default=10000)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    out_dir = args.outdir
    pathlib.Path(out_dir).mkdir(parents=True, exist_ok=True)
    rank, world = parallel.init_from_env()
    pedigree = PedigreeDAG.from_file(args.pedigree)
    genomein = pd.read_csv(args.snps, sep=" ", names=["chrom", "snpid", "cm", "pos"], engine="c")
    if args.empC is not None and rank == 0:
        print("WARNING: -P/--empC is ignored; the empirical blend is not executable in the reference at HEAD")
    print("building cache index from %s" % args.genotypes_input_file)
    if args.genotypes_input_file.endswith(".g2b"):
        # 2-bit packed container (genofix_b200.gfutils.packedgens): no text parsing, same GensCache surface
        from genofix_b200.gfutils.packedgens import PackedGens
        g_cache = PackedGens(args.genotypes_input_file)
    else:
        g_cache = GensCache(args.genotypes_input_file, header=True)
    print("Detected %s individuals and %s snps in input file" % (len(g_cache.all_ids), len(g_cache.snps)))
    snps = [str(x) for x in genomein["snpid"]]
    chromosome2snp = defaultdict(list)
    for c, s in zip(genomein["chrom"], snps):
        chromosome2snp[c].append(s)
    c = CorrectGenotypes(chromosome2snp=chromosome2snp, elimination_order=args.elimination_order)
    tic = time.time()
    corrected = pd.DataFrame(data=g_cache.getMatrix(list(g_cache.all_ids)), index=list(g_cache.all_ids),
                             columns=g_cache.snps)
    if args.first_n_snps is not None:
        corrected = corrected.iloc[:, :args.first_n_snps]
    print("pre-corrected array: %s" % Counter(corrected.to_numpy().flatten()))
    have = set(corrected.columns)
    jobs = []
    for chromosome in sorted(chromosome2snp, key=lambda x: str(x)):
        if str(chromosome) in ("0", "Z"):
            print("chromosome skipped %s" % chromosome)
            continue
        found = [x for x in chromosome2snp[chromosome] if x in have]
        for cols in chunk(found, args.chunksize):
            jobs.append((chromosome, cols))
    tall = np.zeros(len(corrected.index), dtype=np.int64)
    mine = chunks_for_rank(len(jobs), rank, world)
    results = {}
    for ji in mine:
        print("[rank %d] chromosome %s: chunk of %d snps" % (rank, jobs[ji][0], len(jobs[ji][1])))
    for k, (cols, res) in enumerate(c.correctChunks(corrected, [jobs[ji][1] for ji in mine], pedigree, None,
                                                    args.threshold_pairs, args.threshold_singles, back=args.lookback,
                                                    tiethreshold=args.tiethreshold, init_filter_p=args.initquantilefilter,
                                                    filter_e=args.filter_e, weight_empirical=args.weight_empirical,
                                                    threads=args.threads, DEBUGDIR=out_dir)):
        results[mine[k]] = res
        tall += c.last_stats[k]["corrected_per_animal"]
    if world > 1:
        import torch
        import torch.distributed as dist
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        t = torch.as_tensor(tall, device=dev)
        parallel.allreduce_tallies(t)
        tall = t.cpu().numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, {ji: r.to_numpy() for ji, r in results.items()})
        if rank == 0:
            for part in gathered:
                for ji, arr in part.items():
                    corrected.loc[:, jobs[ji][1]] = arr
    else:
        for ji, r in results.items():
            corrected.loc[:, jobs[ji][1]] = r.to_numpy()
    if rank == 0:
        print("done correct matrix in %s minutes" % ((time.time() - tic) / 60))
        corrected.to_csv("%s/simulatedgenome_corrected_errors_threshold.ssv" % out_dir, sep=" ")
        pd.Series(tall, index=corrected.index).to_csv("%s/corrections_per_animal.tsv" % out_dir, sep="\t", header=False)
        print("corrected array: %s" % Counter(corrected.to_numpy().flatten()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
