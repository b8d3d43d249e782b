#!/usr/bin/env python3
"""gfutils.extract_founders -- founders of a pedigree, optionally restricted to sequenced ids
(same CLI as the reference, gfutils/extract_founders.py:85-118)."""
import csv
import gzip
import sys
from argparse import ArgumentParser

import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # src/ (the reference sets PYTHONPATH=src)
import _bootstrap  # noqa: F401,E402
import numpy as np
from genofix_b200.pedigree.pedigree_dag import PedigreeDAG


def founders(pedigree, sequenced=None):
    """Animals lacking a sire or a dam inside the (restricted) pedigree."""
    ids = pedigree.ids
    if sequenced is None:
        keep = np.ones(len(ids), bool)
    else:
        keep = np.isin(ids, np.asarray(list(sequenced), dtype=np.int64))
    s, d = pedigree.sire_idx, pedigree.dam_idx
    has_s = (s >= 0) & keep[np.maximum(s, 0)]
    has_d = (d >= 0) & keep[np.maximum(d, 0)]
    sexed = pedigree.sex > 0
    return ids[keep & sexed & ~(has_s & has_d)]


def main(argv=None):
    parser = ArgumentParser(description="find founders in pedigree")
    parser.add_argument("-p", "--pedigree", dest="pedigree", required=True, help="pedigree file")
    parser.add_argument("-s", "--sequencedtsv", dest="sequencedtsv", required=False,
                        help="space separated file whose first column are the sequenced ids")
    parser.add_argument("-o", "--outfile", dest="outfile", required=False, help="one founder id per line")
    args = parser.parse_args(argv)
    pedigree = PedigreeDAG.from_file(args.pedigree)
    seq = None
    if args.sequencedtsv is not None:
        seq = []
        opener = gzip.open if args.sequencedtsv.endswith(".gz") else open
        with opener(args.sequencedtsv, "rt") as f:
            for row in csv.reader(f, delimiter=" "):
                try:
                    seq.append(int(row[0]))
                except (ValueError, IndexError) as e:
                    sys.stderr.write(str(e))
    out = founders(pedigree, seq)
    if args.outfile is not None:
        with open(args.outfile, "wt") as f:
            for x in out:
                f.write(str(int(x)) + "\n")
    else:
        for x in out:
            print(int(x))
    return 0


if __name__ == "__main__":
    sys.exit(main())
