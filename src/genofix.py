#!/usr/bin/env python3
"""genofix -- correct genotype calls with Mendelian likelihoods on a B200.

Same command line as the reference entry point (src/genofix.py:94-112).  Differences, all forced by the
state of the reference at HEAD:
  * -P/--empC is optional.  Without it correction uses the Mendelian model only (the reference's `empC=None` path, the
    one that completes at HEAD).  With it, <folder>/<chromosome>/empiricalIndex.idx.gz (written by empirical/preindex.py) is
    loaded per chromosome (:209) and blended into ranks and decisions (correct_genotypes3.py:619-667, :746-779, :883-908; the
    reference dies in its pool initializer there, D3).  A chromosome is then ONE matrix: the LD windows must not be cut, so
    -C/--chunksize does not apply (the reference passes the whole chromosome to every chunk call as well, D5);
  * every chunk is corrected from its own columns and written back by column (the reference passes the whole
    chromosome and slices rows, D5);
  * chunks are dealt round-robin to the GPUs of the box when launched under torchrun: every rank reads only the
    SNP blocks it corrects (2-bit container input) and writes them as shard files; the one collective is the
    all-reduce of the per-animal tallies, after which rank 0 assembles the outputs.
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
                   help="folder with prior empirical distribution for ld snps /chr/empiricalIndex.idx.gz (optional)")
    p.add_argument("-C", "--chunksize", dest="chunksize", type=int, default=10000)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    out_dir = args.outdir
    pathlib.Path(out_dir).mkdir(parents=True, exist_ok=True)
    rank, world = parallel.init_from_env()
    pedigree = PedigreeDAG.from_file(args.pedigree)
    genomein = pd.read_csv(args.snps, sep=" ", names=["chrom", "snpid", "cm", "pos"], engine="c")
    print("building cache index from %s" % args.genotypes_input_file)
    packed_input = args.genotypes_input_file.endswith(".g2b")
    if packed_input:
        # 2-bit packed container (genofix_b200.gfutils.packedgens), memory mapped: a rank only touches the pages of
        # the SNP blocks it corrects; nothing is parsed
        from genofix_b200.gfutils.packedgens import PackedGens
        g_cache = PackedGens(args.genotypes_input_file)
    else:
        g_cache = GensCache(args.genotypes_input_file, header=True)      # text has no random column access: parsed by every rank
    ids = [int(x) for x in g_cache.all_ids]
    n_file_snps = len(g_cache.snps) if args.first_n_snps is None else min(len(g_cache.snps), args.first_n_snps)
    file_snps = list(g_cache.snps[:n_file_snps])
    print("Detected %s individuals and %s snps in input file" % (len(ids), len(g_cache.snps)))
    snps = [str(x) for x in genomein["snpid"]]
    chromosome2snp = defaultdict(list)
    for c, s in zip(genomein["chrom"], snps):
        chromosome2snp[c].append(s)
    c = CorrectGenotypes(chromosome2snp=chromosome2snp, elimination_order=args.elimination_order)
    tic = time.time()
    pos = {s: i for i, s in enumerate(file_snps)}
    jobs = []          # (chromosome, column names, column positions in the file)
    for chromosome in sorted(chromosome2snp, key=lambda x: str(x)):
        if str(chromosome) in ("0", "Z"):
            print("chromosome skipped %s" % chromosome)
            continue
        found = [x for x in chromosome2snp[chromosome] if x in pos]
        if args.empC is not None:
            jobs.append((chromosome, found, np.array([pos[x] for x in found], dtype=np.int64)))   # windows must not be cut
            continue
        for cols in chunk(found, args.chunksize):
            jobs.append((chromosome, cols, np.array([pos[x] for x in cols], dtype=np.int64)))
    mine = chunks_for_rank(len(jobs), rank, world)
    for ji in mine:
        print("[rank %d] chromosome %s: chunk of %d snps" % (rank, jobs[ji][0], len(jobs[ji][1])))

    def contiguous(idx):
        return len(idx) > 0 and int(idx[-1]) - int(idx[0]) + 1 == len(idx) and bool(np.all(np.diff(idx) == 1))
    # 2-bit blocks straight from the container when every chunk of this rank is a run of consecutive file columns
    use_packed = packed_input and all(contiguous(jobs[ji][2]) for ji in mine) and args.empC is None
    from genofix_b200.engine import engine_for
    eng = engine_for(pedigree, np.asarray(ids, dtype=np.int64), args.lookback)

    def inputs():
        for ji in mine:
            idx = jobs[ji][2]
            if use_packed:
                yield (g_cache.packed_columns(int(idx[0]), int(idx[-1]) + 1), len(idx))
            elif packed_input:
                yield np.ascontiguousarray(g_cache.getMatrix(ids)[:, idx])
            else:
                yield np.ascontiguousarray(g_cache._matrix[:, idx])
    # every rank writes the blocks it corrected as shard files; the only collective is the all-reduce of the tallies
    shard_dir = pathlib.Path(out_dir) / "shards"
    shard_dir.mkdir(parents=True, exist_ok=True)
    for ji in mine:                                      # leftovers of an interrupted run
        for f in shard_dir.glob("chunk_%06d_*.npy" % ji):
            f.unlink()
    tall = np.zeros(len(ids), dtype=np.int64)
    if args.empC is not None:
        import gzip
        import pickle
        from genofix_b200.correct_genotypes3 import ld_index_arrays
        for k, G in enumerate(inputs()):
            chromosome, cols, _ = jobs[mine[k]]
            empC = pickle.load(gzip.open("%s/%s/empiricalIndex.idx.gz" % (args.empC, chromosome)))       # :209
            print("surround size is %s " % empC.surround_size)
            out = eng.correct_emp(G, ld_index_arrays(empC, cols, args.weight_empirical), args.threshold_pairs,
                                  args.threshold_singles, init_filter_p=args.initquantilefilter, tiethreshold=args.tiethreshold)
            np.save(shard_dir / ("chunk_%06d_u.npy" % mine[k]), out)
            tall += eng.stats["corrected_per_animal"]
    else:
        for k, (out, stats) in enumerate(eng.correct_chunks(inputs(), args.threshold_pairs, args.threshold_singles,
                                                            init_filter_p=args.initquantilefilter, tiethreshold=args.tiethreshold,
                                                            packed_io=use_packed)):
            np.save(shard_dir / ("chunk_%06d_%s.npy" % (mine[k], "p" if use_packed else "u")), out)
            tall += stats["corrected_per_animal"]
    if world > 1:
        import torch
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        t = torch.as_tensor(tall, device=dev)
        parallel.allreduce_tallies(t)      # also orders the shard files of every rank before rank 0 reads them
        tall = t.cpu().numpy()
    if rank == 0:
        from genofix_b200.gfutils.packedgens import PackedGens, pack_codes, unpack_codes
        print("done correct matrix in %s minutes" % ((time.time() - tic) / 60))
        shards = {}
        for f in sorted(shard_dir.glob("chunk_*.npy")):
            ji, kind = int(f.name.split("_")[1]), f.name.split("_")[2][0]
            shards[ji] = (f, kind)
        missing_shards = [ji for ji in range(len(jobs)) if ji not in shards]
        if missing_shards:
            raise RuntimeError("no result for chunks %s" % missing_shards[:5])
        n_cells = len(ids) * n_file_snps
        if packed_input:
            # corrected container next to the text output: the input file's rows with the corrected blocks stored back
            import shutil
            out_g2b = "%s/simulatedgenome_corrected_errors_threshold.g2b" % out_dir
            shutil.copyfile(args.genotypes_input_file, out_g2b)
            dst = PackedGens(out_g2b, mode="r+")
            for ji, (f, kind) in shards.items():
                idx = jobs[ji][2]
                blk = np.load(f)
                if kind == "p":
                    dst.store_columns(int(idx[0]), int(idx[-1]) + 1, blk)
                else:
                    for a in range(len(idx)):          # scattered columns: one at a time
                        dst.store_columns(int(idx[a]), int(idx[a]) + 1, pack_codes(blk[:, a:a + 1]))
            dst.packed.flush()
            print("corrected container: %s" % out_g2b)
        if n_cells <= 500_000_000:
            if packed_input:
                corrected = pd.DataFrame(g_cache.getMatrix(ids)[:, :n_file_snps], index=ids, columns=file_snps)
            else:
                corrected = pd.DataFrame(np.array(g_cache._matrix[:, :n_file_snps]), index=ids, columns=file_snps)
            print("pre-corrected array: %s" % Counter(corrected.to_numpy().flatten()))
            M = corrected.to_numpy(copy=True)
            for ji, (f, kind) in shards.items():
                blk = np.load(f)
                M[:, jobs[ji][2]] = unpack_codes(blk, len(jobs[ji][2])) if kind == "p" else blk
            corrected = pd.DataFrame(M, index=ids, columns=file_snps)
            corrected.to_csv("%s/simulatedgenome_corrected_errors_threshold.ssv" % out_dir, sep=" ")
            print("corrected array: %s" % Counter(M.flatten()))
        else:
            print("text output skipped for %d cells (the 2-bit container holds the result)" % n_cells)
        pd.Series(tall, index=ids).to_csv("%s/corrections_per_animal.tsv" % out_dir, sep="\t", header=False)
        for f, _kind in shards.values():
            f.unlink()
        try:
            shard_dir.rmdir()
        except OSError:
            pass
    return 0


if __name__ == "__main__":
    sys.exit(main())
