#!/usr/bin/env python3
"""empirical.preindex -- builds the Mendel-filtered LD index; same module path and command line as the reference
(empirical/preindex.py:106-113), computed on the GPU.  One gzip-pickled JointAllellicDistribution per chromosome in
<outdir>/<chromosome>/empiricalIndex.idx.gz (a later seed chunk overwrites an earlier one, like the reference, :339-346)."""
import gzip
import os
import pathlib
import pickle
import sys
from argparse import ArgumentParser
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # src/ (the reference sets PYTHONPATH=src)
import _bootstrap  # noqa: F401,E402
import pandas as pd  # noqa: E402
from gfutils.gensrandaccess import GensCache  # noqa: E402
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG  # noqa: E402
from genofix_b200.empirical.preindex import seed_chunks, index_chunk  # noqa: E402


def main(argv=None):
    parser = ArgumentParser(description="index of empirical LD-window genotype frequencies after a Mendel-rank filter")
    parser.add_argument("-g", "--genotypes", dest="genotypes_input_file", required=True)
    parser.add_argument("-p", "--pedigree", dest="pedigree", required=True)
    parser.add_argument("-w", "--surroundsnps", dest="surround_size", type=int, default=3)
    parser.add_argument("-o", "--outdir", dest="outdir", type=str, required=True)
    parser.add_argument("-s", "--snps", dest="snps", required=True, help="snp map file: chrom,snpid,cM,pos without header")
    parser.add_argument("-T", "--threads", dest="threads", type=int, required=False, default=os.cpu_count())
    parser.add_argument("-q", "--initquantilefilter", dest="initquantilefilter", type=float, required=False, default=0.9)
    parser.add_argument("-c", "--seedskidschunk", dest="seedskidschunk", type=int, required=False, default=10000)
    args = parser.parse_args(argv)
    pathlib.Path(args.outdir).mkdir(parents=True, exist_ok=True)
    g_cache = GensCache(args.genotypes_input_file, header=True)
    snp_names = [s.rstrip("\n") for s in g_cache.snps]
    pedigree = PedigreeDAG.from_file(args.pedigree)
    genomein = pd.read_csv(args.snps, sep=",", names=["chrom", "snpid", "cM", "pos"], skiprows=0)
    genomein = genomein.sort_values(by=["chrom", "pos"], ascending=[True, True])
    chromosome2snp = defaultdict(list)
    for chrom, snp in zip(genomein["chrom"].astype(str), genomein["snpid"]):
        if snp in chromosome2snp[chrom]:
            raise Exception("Error in map file %s appears multiple times in chromosome %s " % (snp, chrom))
        chromosome2snp[chrom].append(snp)
    chunks = seed_chunks(pedigree, g_cache.all_ids, args.seedskidschunk)
    for i, (kid_ids, cand) in enumerate(chunks):
        print("chunk %s of %s: %s kids, %s with their genotyped parents" % (i, len(chunks), len(kid_ids), len(cand)))
        genotypes = pd.DataFrame(g_cache.getMatrix(cand), index=cand, columns=snp_names)
        for chromosome, (empC, info) in index_chunk(genotypes, pedigree, chromosome2snp, args.surround_size,
                                                    args.initquantilefilter).items():
            print("chromosome %s: cut %s, %s cells under the cut" % (chromosome, info["quantQ"], info["n_unmasked"]))
            pathlib.Path("%s/%s" % (args.outdir, chromosome)).mkdir(parents=True, exist_ok=True)
            with gzip.open("%s/%s/empiricalIndex.idx.gz" % (args.outdir, chromosome), "wb") as fout:
                pickle.dump(empC, fout)
    return 0


if __name__ == "__main__":
    sys.exit(main())
