"""LD-window joint genotype frequencies with the reference's call surface (empirical/jalleledist3.py).

The counting (countJointFrq / countJointFrqAll, jalleledist3.py:120-174) is one CUDA kernel over all
windows (gfx_window_counts): per window a 3^w-bin shared-memory histogram of the base-3 state index of
every animal whose window cells all pass the mask, plus one smaller histogram per nested "inner window"
of :133-138 (a row with a missing call in an OUTER cell still counts there, as in the reference).  The
float recurrence of :133-138 (including its i = 0 quirk) is applied on the host in float64,
in the reference's operation order.
"""
import ctypes as C
import itertools
import sys

import numpy as np

from .. import _lib


def window_bounds(n_snps, pos, surround):
    """[start, stop) of the window of SNP `pos` (jalleledist3.py:62-75).  The reference clips the end
    to len-1 *exclusive*, so windows touching the chromosome end lose the last SNP."""
    start = max(0, pos - surround)
    stop = pos + surround + 1
    if stop >= n_snps:
        stop = n_snps - 1
    return start, max(stop, start)


def window_counts_device(G, mask, starts, lens):
    """Integer joint counts for arbitrary windows.  G uint8 [n, s] (0,1,2, anything else = missing),
    mask bool [n, s] or None.  Returns int64 [n_windows, counters_per_window(max_len)]: per window the 3^len
    full histogram followed by its inner-window histograms (include/genofix_b200.h)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("genofix_b200 needs a CUDA device; there is no CPU fallback")
    L = _lib.lib()
    G = np.ascontiguousarray(G, dtype=np.uint8)
    n, s = G.shape
    starts = np.ascontiguousarray(starts, dtype=np.int32)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    nw = len(starts)
    stride = counters_per_window(max(int(lens.max()), 1))
    ld = ((s + 15) // 16) * 16
    host = np.full((n, ld), 9, np.uint8)
    host[:, :s] = G
    codes = torch.from_numpy(host).cuda()
    wpr = (((s + 15) // 16 + 3) // 4) * 4
    packed = torch.empty((n, wpr), dtype=torch.int32, device="cuda")
    _lib.check(L.gfx_pack2(codes.data_ptr(), n, s, ld, packed.data_ptr(), wpr, None))
    mdev, mwpr = None, 0
    if mask is not None:
        m = np.asarray(mask, dtype=bool)
        mwpr = (s + 31) // 32
        bits = np.zeros((n, mwpr * 32), dtype=bool)
        bits[:, :s] = m
        words = np.packbits(bits.reshape(n, mwpr, 32), axis=2, bitorder="little").view(np.uint32).reshape(n, mwpr)
        mdev = torch.from_numpy(words.view(np.int32).copy()).cuda()
    counts = torch.zeros((nw, stride), dtype=torch.int32, device="cuda")
    ws = torch.from_numpy(starts).cuda()
    wl = torch.from_numpy(lens).cuda()
    _lib.check(L.gfx_window_counts(packed.data_ptr(), mdev.data_ptr() if mdev is not None else None, n, s, wpr, mwpr,
                                   ws.data_ptr(), wl.data_ptr(), nw, stride, counts.data_ptr(), None))
    torch.cuda.synchronize()
    return counts.cpu().numpy().view(np.uint32).astype(np.int64)


def inner_levels(w):
    """Number of nested windows gfx_window_counts keeps per window: the full one + int((w-2)/2) - 1 inner ones."""
    return max(1, int((w - 2) / 2))


def counters_per_window(w):
    return int(sum(3 ** (w - 2 * l) for l in range(inner_levels(w))))


def frequencies_from_counts(counts, w, pseudocount, sum_inner_count=True):
    """Apply jalleledist3.py:130-138 to the counters of one window: 3^w entries of the full window (first SNP =
    most significant digit) followed by the histograms of the inner windows 1, 2, ... (cells i .. w-1-i, counted
    over the rows whose whole window passes the mask, whatever the outer cells hold).  Returns float64 [3^w]."""
    c = np.asarray(counts[:3 ** w], dtype=np.int64).reshape((3,) * w) if w > 0 else np.asarray(counts[:1])
    val = c.astype(np.float64) + pseudocount
    if sum_inner_count and w > 0:
        off = 3 ** w
        for i in range(int((w - 2) / 2)):
            if i == 0:
                obs = np.ones_like(val)          # np.count_nonzero(np.logical_and.reduce([])) == 1
            else:
                wi = w - 2 * i
                inner = np.asarray(counts[off:off + 3 ** wi], dtype=np.int64).reshape((1,) * i + (3,) * wi + (1,) * i)
                off += 3 ** wi
                obs = np.broadcast_to(inner, c.shape).astype(np.float64)
            obs = obs - val
            val = (obs / (i + 1)) + val
    return val.reshape(-1)


class _FrequencyView(object):
    """dict-like snp -> {state key: value} built on demand (the reference keeps 3^w tuple keys per SNP)."""

    def __init__(self, owner):
        self.o = owner

    def __getitem__(self, snp):
        o = self.o
        win = o.windows[snp]
        vals = o._freq[snp]
        keys = [tuple(zip(win, v)) for v in itertools.product(o.conditions_index, repeat=len(win))]
        return dict(zip(keys, vals[:len(keys)]))

    def __contains__(self, snp):
        return snp in self.o._freq

    def keys(self):
        return self.o._freq.keys()


class JointAllellicDistribution(object):

    def __init__(self, chromosome, snp_ordered, chromosome2snp=None, pseudocount=0.0001, surround_size=1,
                 conditions_index=[0, 1, 2]):
        self.chromosome2snp = chromosome2snp
        self.pseudocount = pseudocount
        self.surround_size = surround_size
        self.chromosome = chromosome
        self.window_size = (surround_size * 2) + 1
        self.conditions_index = list(conditions_index)
        self.snp_ordered = [sys.intern(x) for x in snp_ordered
                            if chromosome2snp is None or x in chromosome2snp[chromosome]]
        n = len(self.snp_ordered)
        self._bounds = {snp: window_bounds(n, i, surround_size) for i, snp in enumerate(self.snp_ordered)}
        self.windows = {snp: self.snp_ordered[a:b] for snp, (a, b) in self._bounds.items()}
        self.state_values = list(itertools.product(self.conditions_index, repeat=self.window_size))
        self._freq = {snp: np.full(3 ** len(self.windows[snp]), pseudocount, dtype=np.float64)
                      for snp in self.snp_ordered}
        self.snp2frequency = _FrequencyView(self)

    def getWindow(self, targetSnp):
        return self.windows.get(targetSnp)

    def getObservations(self, targetSnp):
        return self.getCountTable({x: 9 for x in self.getWindow(targetSnp)}, targetSnp)

    def getCountTable(self, observedstates, targetSnp):
        """Stored frequency of each state of targetSnp given the animal's window genotypes; unknown (9 or
        absent) window SNPs are summed over their completions.  (The reference's own version raises
        AttributeError whenever a 9 is present, jalleledist3.py:84-96 -- SURVEY D4 -- so the
        missing-data case has no reference behaviour to match.)"""
        win = self.windows[targetSnp]
        w = len(win)
        if targetSnp not in win:
            # D11: the last SNP's window does not contain it; every key is unchanged by the substitution
            pass
        table = self._freq[targetSnp].reshape((3,) * w) if w > 0 else self._freq[targetSnp]
        out = []
        for state in (0, 1, 2):
            idx = []
            for snp in win:
                if snp == targetSnp:
                    idx.append(state)
                else:
                    v = observedstates.get(snp, 9)
                    idx.append(int(v) if v in (0, 1, 2) else slice(None))
            out.append(np.sum(table[tuple(idx)]))
        return out

    @staticmethod
    def countJointFrq(table, mask, column_names, s_values, pseudocount, sum_inner_count=True):
        """{state key: frequency} of one window (jalleledist3.py:120-141).  table uint8 [N, w]."""
        table = np.asarray(table)
        w = table.shape[1]
        counts = window_counts_device(table, np.asarray(mask, dtype=bool), [0], [w])[0]
        vals = frequencies_from_counts(counts, w, pseudocount, sum_inner_count)
        names = list(column_names)
        order = {v: i for i, v in enumerate(itertools.product([0, 1, 2], repeat=w))}
        return {tuple(zip(names, v)): vals[order[tuple(v)]] for v in s_values}

    def countJointFrqAll(self, table, mask=None, threads=None):
        """Fill the frequencies of every SNP window from a DataFrame (columns = SNP ids) in one kernel
        launch (jalleledist3.py:143-174)."""
        cols = list(table.columns)
        pos = {c: i for i, c in enumerate(cols)}
        G = table.to_numpy(dtype=np.uint8)
        if mask is not None and hasattr(mask, "to_numpy"):
            mask = mask.to_numpy(dtype=bool)
        snps = [s for s in self.snp_ordered if len(self.windows[s]) > 0]
        starts, lens = [], []
        for s in snps:
            win = self.windows[s]
            a = pos[win[0]]
            if [pos[x] for x in win] != list(range(a, a + len(win))):
                raise ValueError("table columns must be in map order for window %s" % s)
            starts.append(a)
            lens.append(len(win))
        if not snps:
            return
        counts = window_counts_device(G, mask, starts, lens)
        for i, s in enumerate(snps):
            self._freq[s] = frequencies_from_counts(counts[i], lens[i], self.pseudocount, True)
        print("empirical count done")
