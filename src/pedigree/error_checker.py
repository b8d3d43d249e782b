#!/usr/bin/env python3
"""pedigree.error_checker -- highlights possible pedigree errors by Mendel statistics; same module path and
command line as the reference (pedigree/error_checker.py:85-93,199-204), computed on the GPU."""
import os
import pathlib
import sys
from argparse import ArgumentParser

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # src/ (the reference sets PYTHONPATH=src)
import _bootstrap  # noqa: F401,E402
import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
from gfutils.gensrandaccess import GensCache  # noqa: E402
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG  # noqa: E402
from genofix_b200.pedigree.error_checker import trio_candidates, mendel_error_tally  # noqa: F401,E402


def main(argv=None):
    """Same command line as the reference (error_checker.py:85-93); writes probs_errors.csv and
    individual_sum_probs_errors.csv into --outdir."""
    parser = ArgumentParser(description="highlights possible pedigree errors by Mendel statistics")
    parser.add_argument("-g", "--genotypes", dest="genotypes_input_file", required=True)
    parser.add_argument("-p", "--pedigree", dest="pedigree", required=True)
    parser.add_argument("-k", "--kids", dest="kids", required=True)
    parser.add_argument("-s", "--snps", dest="snps", required=True)
    parser.add_argument("-T", "--threads", dest="threads", type=int, required=False, default=os.cpu_count())
    parser.add_argument("-c", "--chromosome", dest="chromosome", type=str, required=True, default="1")
    parser.add_argument("-o", "--outdir", dest="outdir", type=str, required=True)
    parser.add_argument("-d", "--delimiter", dest="delimiter_pedigree", default=" ", type=str, required=False)
    args = parser.parse_args(argv)
    pathlib.Path(args.outdir).mkdir(parents=True, exist_ok=True)
    kids = pd.read_csv(args.kids, sep=args.delimiter_pedigree).iloc[:, 0].to_list()
    g_cache = GensCache(args.genotypes_input_file, header=True)
    pedigree = PedigreeDAG.from_file(args.pedigree, delimiter=args.delimiter_pedigree)
    genomein = pd.read_csv(args.snps, sep=",", names=["snpid", "chrom", "pos", "topAllele", "B"], skiprows=1)
    genomein = genomein.sort_values(by=["chrom", "pos"], ascending=[True, True])
    on_chrom = genomein[genomein["chrom"].astype(str) == str(args.chromosome)]["snpid"].tolist()
    if len(set(on_chrom)) != len(on_chrom):
        raise Exception("Error in map file: a snp appears multiple times in chromosome %s" % args.chromosome)
    print("chromsome has %s snps" % len(on_chrom))
    snp_names = [s.rstrip("\n") for s in g_cache.snps]
    all_ids = list(g_cache.all_ids)
    with_parents, cand = trio_candidates(pedigree, all_ids, kids)
    print("Found %s out of %s kids have genotypes and trios" % (len(with_parents), len(kids)))
    genotypes = pd.DataFrame(g_cache.getMatrix(cand), index=cand, columns=snp_names).loc[:, on_chrom]
    probs_errors, sums = mendel_error_tally(genotypes, pedigree, kids)
    probs_errors.to_csv("%s/probs_errors.csv" % args.outdir, index=True, header=True)
    np.savetxt("%s/individual_sum_probs_errors.csv" % args.outdir, sums, delimiter=",")
    return 0


if __name__ == "__main__":
    sys.exit(main())
