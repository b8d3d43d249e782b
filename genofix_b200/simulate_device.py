"""Gene drop, error injection and evaluation on the device.

The reference simulates with quickgsim (simulate.py:187-340) and scores with pandas (simulate.py:426-515).  Without a genetic
map SNPs are dropped independently (free recombination: all the Mendelian path sees); with one (`linkage_map`) the drop follows
quickgsim's meiosis (quickgsim/genome.py:74-115): per gamete and chromosome max(1, Poisson(length in Morgans)) crossovers on
interior markers, placed with probability proportional to the genetic distance to the previous marker, random start strand.
The generators are counter based (splitmix64 of seed, row, word): results depend on the seed
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


def linkage_arrays(chrom, cm):
    """(chrom_start int32 [C + 1], morgans float32 [C], cum_p float32 [S]) of a genetic map given per SNP in map order:
    quickgsim.genome.Chrom.finalise_chrom_configuration / cm_to_p (:21-41): a chromosome is max(cM) / 100 Morgans long; the
    crossover probability of an interior marker is its genetic distance to the previous marker over the sum of those of the
    interior markers (first and last marker excluded); cum_p is the running sum per chromosome."""
    chrom = np.asarray(chrom)
    cm = np.asarray(cm, dtype=np.float64)
    if len(chrom) != len(cm):
        raise ValueError("one chromosome and one cM position per SNP")
    cuts = [0] + [i for i in range(1, len(chrom)) if chrom[i] != chrom[i - 1]] + [len(chrom)]
    if len(set(chrom[a] for a in cuts[:-1])) != len(cuts) - 1:
        raise ValueError("SNPs must be grouped by chromosome (map order)")
    start = np.array(cuts, dtype=np.int32)
    morgans = np.zeros(len(cuts) - 1, dtype=np.float32)
    cum_p = np.zeros(len(chrom), dtype=np.float32)
    for c in range(len(cuts) - 1):
        a, b = cuts[c], cuts[c + 1]
        pos = cm[a:b]
        m = float(pos.max()) / 100.0
        morgans[c] = m
        if b - a >= 3:
            d = np.diff(pos, prepend=0.0)[1:-1]                  # interior markers
            tot = d.sum()
            p = d / tot if tot > 0 else np.full(len(d), 1.0 / len(d))
            cp = np.cumsum(p)
            cp[-1] = 1.0
            cum_p[a + 1:b - 1] = cp
            cum_p[b - 1] = 1.0
    return start, morgans, cum_p


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

    def gene_drop(self, linkage_map=None, first_level=0):
        """packed truth genotypes, int32 cuda tensor [n, wpr] (the kernels' 2-bit layout).  linkage_map = (chromosome of every
        SNP, cM position of every SNP), SNPs in map order: quickgsim's meiosis instead of free recombination.
        first_level > 0 keeps the haplotypes of the shallower levels as they are in self.hap (tests set founders by hand)."""
        t = self.torch
        lo = self.level_off[first_level:]
        if linkage_map is None:
            _lib.check(self.L.gfx_gene_drop(self.n, self.sire.data_ptr(), self.dam.data_ptr(), self.order.data_ptr(),
                                            lo.ctypes.data_as(C.c_void_p), len(lo) - 1, self.s, self.wpr,
                                            C.c_uint64(self.seed), self.hap.data_ptr(), self._stream()))
        else:
            start, morgans, cum_p = linkage_arrays(*linkage_map)
            if int(start[-1]) != self.s:
                raise ValueError("the genetic map must cover every SNP")
            dev = self.hap.device
            keep = [t.from_numpy(start).to(dev), t.from_numpy(morgans).to(dev), t.from_numpy(cum_p).to(dev)]
            _lib.check(self.L.gfx_gene_drop_linked(self.n, self.sire.data_ptr(), self.dam.data_ptr(), self.order.data_ptr(),
                                                   lo.ctypes.data_as(C.c_void_p), len(lo) - 1, self.s, self.wpr,
                                                   C.c_uint64(self.seed), len(morgans), keep[0].data_ptr(), keep[1].data_ptr(),
                                                   keep[2].data_ptr(), self.hap.data_ptr(), self._stream()))
            t.cuda.current_stream().synchronize()
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
