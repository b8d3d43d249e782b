"""Gene drop, error injection and evaluation on the device.

The reference simulates with quickgsim (simulate.py:187-340) and scores with pandas (simulate.py:426-515); with the
empirical LD model off only the marginal Mendelian structure of the data matters, so SNPs are dropped independently
(free recombination).  The generators are counter based (splitmix64 of seed, row, word): results depend on the seed
only, not on the launch geometry.  They are NOT the numpy generators of genofix_b200.synth (different streams)."""
import ctypes as C

import numpy as np

from . import _lib


def levels_from_parents(sire_row, dam_row):
    """(order, level_off): rows sorted by depth (founders = level 0; a row is one level below its deeper parent)."""
    sire_row = np.asarray(sire_row, dtype=np.int64)
    dam_row = np.asarray(dam_row, dtype=np.int64)
    n = len(sire_row)
    depth = np.zeros(n, dtype=np.int64)
    for _ in range(n + 1):
        ds = np.where(sire_row >= 0, depth[np.maximum(sire_row, 0)] + 1, 0)
        dd = np.where(dam_row >= 0, depth[np.maximum(dam_row, 0)] + 1, 0)
        new = np.maximum(ds, dd)
        if np.array_equal(new, depth):
            break
        depth = new
    else:
        raise ValueError("pedigree has a cycle")
    order = np.argsort(depth, kind="stable").astype(np.int32)
    level_off = np.concatenate([[0], np.cumsum(np.bincount(depth))]).astype(np.int32)
    return order, level_off


class DeviceSimulator(object):
    """rows: int64 [n, 4] (kid, sire, dam, sex), every animal gets a genotype row (row i = rows[i])."""

    def __init__(self, rows, n_snps, seed=1234):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("genofix_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.torch = torch
        self.L = _lib.lib()
        rows = np.asarray(rows, dtype=np.int64)
        pos = {int(k): i for i, k in enumerate(rows[:, 0])}
        self.n = len(rows)
        self.s = int(n_snps)
        self.wpr = (((self.s + 15) // 16) + 3) // 4 * 4
        sire = np.array([pos.get(int(x), -1) if x > 0 else -1 for x in rows[:, 1]], dtype=np.int32)
        dam = np.array([pos.get(int(x), -1) if x > 0 else -1 for x in rows[:, 2]], dtype=np.int32)
        order, self.level_off = levels_from_parents(sire, dam)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.sire = torch.from_numpy(sire).to(dev)
        self.dam = torch.from_numpy(dam).to(dev)
        self.order = torch.from_numpy(order).to(dev)
        self.seed = int(seed)
        self.hap = torch.zeros((self.n, self.wpr), dtype=torch.int32, device=dev)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def gene_drop(self):
        """packed truth genotypes, int32 cuda tensor [n, wpr] (the kernels' 2-bit layout)"""
        t = self.torch
        _lib.check(self.L.gfx_gene_drop(self.n, self.sire.data_ptr(), self.dam.data_ptr(), self.order.data_ptr(),
                                        self.level_off.ctypes.data_as(C.c_void_p), len(self.level_off) - 1, self.s, self.wpr,
                                        C.c_uint64(self.seed), self.hap.data_ptr(), self._stream()))
        truth = t.empty((self.n, self.wpr), dtype=t.int32, device=self.hap.device)
        _lib.check(self.L.gfx_haps_to_codes(self.hap.data_ptr(), self.n, self.s, self.wpr, truth.data_ptr(), self._stream()))
        return truth

    def inject_errors(self, packed, rate, seed=None):
        """copy of `packed` with every observed cell replaced by another genotype with probability `rate`"""
        out = packed.clone()
        _lib.check(self.L.gfx_inject_errors(out.data_ptr(), self.n, self.s, self.wpr, float(rate),
                                            C.c_uint64(self.seed + 1 if seed is None else int(seed)), self._stream()))
        return out

    def confusion(self, truth, observed, corrected):
        """dict of the evaluation counts of simulate.py:426-515 (+ precision / recall / f1)"""
        t = self.torch
        out = t.zeros(8, dtype=t.int64, device=truth.device)
        _lib.check(self.L.gfx_confusion(truth.data_ptr(), observed.data_ptr(), corrected.data_ptr(), self.n, self.s, self.wpr,
                                        out.data_ptr(), self._stream()))
        c = [int(x) for x in out.cpu().numpy()]
        flagged = c[1] + c[2] + c[3]
        precision = flagged / max(1, flagged + c[4])
        recall = flagged / max(1, flagged + c[5])
        return dict(errors=c[0], fixed=c[1], blanked=c[2], wrong_fix=c[3], new_errors=c[4], missed=c[5], cells=c[6],
                    precision=precision, recall=recall, f1=2 * precision * recall / max(1e-12, precision + recall))
