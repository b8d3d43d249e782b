"""gfutils.gensrandaccess -- genotype text matrix reader with the reference's surface
(gfutils/gensrandaccess.py:10-50): first column animal id, then one 0/1/2/9 code per SNP.

The reference indexes line numbers and re-reads lines through `linecache`; here the file is parsed once
into a uint8 matrix (the GPU path wants the whole matrix anyway) and `getMatrix(ids)` is a row gather.
The trailing newline the reference leaves on the last SNP name is stripped."""
import gzip

import numpy as np


class GensCache(object):
    def __init__(self, gensfile, header=False, delimiter=' '):
        self.gensfile = gensfile
        self.snps = None
        opener = gzip.open if gensfile.endswith(".gz") else open
        ids, rows = [], []
        with opener(gensfile, "rt") as f:
            if header:
                self.snps = f.readline().rstrip("\n").split(delimiter)[1:]
            for line in f:
                parts = line.rstrip("\n").split(delimiter)
                if not parts or parts[0] == "":
                    continue
                ids.append(int(parts[0]))
                rows.append(np.array(parts[1:], dtype=np.uint8))
        self.all_ids = ids
        self.kid2index = {k: i for i, k in enumerate(ids)}
        self._matrix = np.vstack(rows) if rows else np.zeros((0, 0), np.uint8)

    def getMatrix(self, ids):
        idx = [self.kid2index[int(x)] for x in ids if int(x) in self.kid2index]
        return self._matrix[idx]
