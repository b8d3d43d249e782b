"""Device pipeline behind CorrectGenotypes.correctMatrix: schedule + kernels through the C ABI.

PyTorch is used only for device memory, pinned host buffers and streams.  Every compute step is a
kernel of libgenofix_b200.so; the one host-side numeric step is the rank transform table
(log(log(x+1)+1)/max over the <= 65536 distinct float16 scores), done with numpy so that it is
bit-identical to the reference's own numpy calls (correct_genotypes3.py:600-603).
"""
import ctypes as C
import os
import warnings

import numpy as np

from . import _lib
from .pedigree.pedigree_dag import PedigreeDAG

SUSPECT_T = 0.2  # hard-coded in the reference driver (correct_genotypes3.py:691,858)
POST_SENTINEL = -1.0  # posterior buffers are prefilled with it; evaluated cells hold probabilities or NaN


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("genofix_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


_POOL = None


def _parallel_copy(dst, src, threads=None):
    """dst[...] = src for large 2-D arrays, row blocks on a few threads (numpy releases the GIL while copying):
    staging a 100 MB chunk through pinned memory is otherwise the longest step of the end-to-end call."""
    global _POOL
    n = dst.shape[0]
    if threads is None:
        # one rank per GPU: share the host cores with the other ranks of the box
        threads = max(2, min(16, (os.cpu_count() or 2) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    if dst.size < (1 << 22) or n < threads:
        dst[...] = src
        return
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=threads)
    step = (n + threads - 1) // threads

    def work(i):
        dst[i:i + step] = src[i:i + step]
    list(_POOL.map(work, range(0, n, step)))


class Schedule(object):
    """Flattened pedigree for a given list of genotyped ids (matrix row order) and look-back depth."""

    def __init__(self, pedigree, ids, back=2, host_only=False):
        if not isinstance(pedigree, PedigreeDAG):
            raise TypeError("pedigree must be a genofix_b200 PedigreeDAG")
        L = _lib.lib()
        ids = np.asarray(ids, dtype=np.int64)
        pidx = pedigree.index_of(ids)
        if (pidx < 0).any():
            # the reference fails with NetworkXError when a genotyped id is missing from the pedigree
            bad = ids[pidx < 0][:5]
            try:
                import networkx as nx
                raise nx.NetworkXError("The node %s is not in the digraph." % bad[0])
            except ImportError:
                raise KeyError(int(bad[0]))
        n_ped = len(pedigree.ids)
        row = np.full(n_ped, -1, dtype=np.int32)
        row[pidx] = np.arange(len(ids), dtype=np.int32)
        if len(np.unique(pidx)) != len(pidx):
            raise ValueError("duplicate ids in the genotype index")
        kid_off, kid = pedigree.children_csr_reference_order()
        self._keep = dict(ped_id=np.ascontiguousarray(pedigree.ids, dtype=np.int64),
                          sire=np.ascontiguousarray(pedigree.sire_idx, dtype=np.int32),
                          dam=np.ascontiguousarray(pedigree.dam_idx, dtype=np.int32),
                          generation=np.ascontiguousarray(pedigree.generation, dtype=np.int32),
                          row=row, kid_off=np.ascontiguousarray(kid_off, dtype=np.int32),
                          kid=np.ascontiguousarray(kid, dtype=np.int32))
        k = self._keep
        self.desc = _lib.PedigreeDesc(n_ped, len(ids), int(back), _ptr(k["ped_id"]), _ptr(k["sire"]), _ptr(k["dam"]),
                                      _ptr(k["generation"]), _ptr(k["row"]), _ptr(k["kid_off"]), _ptr(k["kid"]))
        self.handle = C.c_void_p()
        fn = L.gfx_schedule_create_host if host_only else L.gfx_schedule_create
        _lib.check(fn(C.byref(self.desc), C.byref(self.handle)))
        info = _lib.ScheduleInfo()
        _lib.check(L.gfx_schedule_info(self.handle, C.byref(info)))
        self.info = {f: getattr(info, f) for f, _ in _lib.ScheduleInfo._fields_}
        self.n_rows = len(ids)
        self.back = int(back)
        self.ids = ids

    def array(self, which):
        cap = {0: 2 * self.info["n_pairs"], 1: self.info["n_generations"] + 1, 2: self.info["n_singles"],
               3: self.info["n_pair_jobs"], 4: self.n_rows}[which]
        out = np.zeros(max(cap, 1), dtype=np.int32)
        n = _lib.check(_lib.lib().gfx_schedule_copy(self.handle, which, _ptr(out), cap))
        return out[:n]

    def founders(self, restrict_ids=None):
        """Pedigree ids lacking a sire or a dam (gfutils/extract_founders.py:106-110)."""
        n_ped = self.desc.n_ped
        flags = None
        if restrict_ids is not None:
            flags = np.zeros(n_ped, dtype=np.uint8)
            idx = np.searchsorted(self._keep["ped_id"], np.asarray(restrict_ids, dtype=np.int64))
            idx = idx[(idx < n_ped)]
            idx = idx[np.isin(self._keep["ped_id"][idx], np.asarray(restrict_ids, dtype=np.int64))]
            flags[idx] = 1
        out = np.zeros(n_ped, dtype=np.int32)
        n = C.c_int32(0)
        _lib.check(_lib.lib().gfx_founders(C.byref(self.desc), _ptr(flags) if flags is not None else None, _ptr(out),
                                           C.byref(n)))
        return self._keep["ped_id"][out[:n.value]]

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            _lib.lib().gfx_schedule_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------
# rank transform + quantile on the histogram (host, numpy)
# ---------------------------------------------------------------------------------------------------
_ALL_HALF = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float64)


def _lerp(a, b, t):
    # numpy.lib._function_base_impl._lerp
    d = b - a
    r = a + d * t
    if t >= 0.5:
        r = b - d * (1 - t)
    return r


def weighted_quantile_linear(values, counts, q):
    """np.quantile(x, q) (default 'linear' method) for x = values repeated counts times; values need
    not be sorted.  Bit-identical to numpy (checked in tests/test_host_logic.py)."""
    values = np.asarray(values, dtype=np.float64)
    counts = np.asarray(counts, dtype=np.int64)
    if np.isnan(values[counts > 0]).any():
        return np.float64(np.nan)
    order = np.argsort(values, kind="stable")
    v = values[order]
    cum = np.cumsum(counts[order])
    n = int(cum[-1])
    vi = np.float64(n - 1) * np.float64(q)
    lo = np.floor(vi)
    if vi >= n - 1:
        lo_i = hi_i = n - 1
    elif vi < 0:
        lo_i = hi_i = 0
    else:
        lo_i = int(lo)
        hi_i = lo_i + 1
    gamma = vi - lo
    a = v[np.searchsorted(cum, lo_i, side="right")]
    b = v[np.searchsorted(cum, hi_i, side="right")]
    return np.float64(_lerp(a, b, gamma))


def weighted_nanquantile(values, counts, qs, alpha=0.0, beta=1.0):
    """np.nanquantile(x, qs, method=...) for x = values repeated counts times, for the continuous methods numpy
    defines through _compute_virtual_index: (alpha, beta) = (0, 1) is 'interpolated_inverted_cdf'
    (empirical/preindex.py:270,301); 'linear' uses another index formula, see weighted_quantile_linear.  values keep
    their dtype: for float16 data numpy subtracts the two neighbouring order statistics in float16 and interpolates in
    float64 (numpy.lib._function_base_impl._lerp), which is reproduced here.  Returns float64 [len(qs)]."""
    values = np.asarray(values)
    counts = np.asarray(counts, dtype=np.int64)
    keep = (counts > 0) & ~np.isnan(values.astype(np.float64))
    v, c = values[keep], counts[keep]
    qs = np.asarray(qs, dtype=np.float64)
    if len(v) == 0:
        return np.full(qs.shape, np.nan)
    order = np.argsort(v.astype(np.float64), kind="stable")
    v = v[order]
    cum = np.cumsum(c[order])
    n = int(cum[-1])
    vi = n * qs + (alpha + qs * (1 - alpha - beta)) - 1            # _compute_virtual_index
    lo = np.floor(vi)
    hi = lo + 1
    above, below = vi >= n - 1, vi < 0                              # _get_indexes
    lo[above], hi[above] = n - 1, n - 1
    lo[below], hi[below] = 0, 0
    gamma = vi - np.floor(vi)
    a = v[np.searchsorted(cum, lo.astype(np.int64), side="right")]
    b = v[np.searchsorted(cum, hi.astype(np.int64), side="right")]
    with np.errstate(all="ignore"):
        d = np.subtract(b, a)                                       # in the data dtype
        r = np.asanyarray(np.add(a, d * gamma), dtype=np.float64)
        alt = np.subtract(b, d * (1 - gamma))
        r = np.where(gamma >= 0.5, alt, r)
    return np.asarray(r, dtype=np.float64)


_E_ALL = None


def _e_all():
    """log(log(x + 1) + 1) of every float16 value (constant; numpy's log so that it matches the reference)."""
    global _E_ALL
    if _E_ALL is None:
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _E_ALL = np.log(np.log(_ALL_HALF + 1) + 1)
    return _E_ALL


def rank_transform_table(hist, init_filter_p):
    """hist: uint64[65537] from gfx_rank_hist.  Returns (elut float64[65536], test_p, n_nan_cells).

    E = log(log(raw+1)+1) / nanmax(.)  with missing cells set to 1 afterwards
    (correct_genotypes3.py:600-603); test_p = np.quantile(E, init_filter_p) (:666-667)."""
    hist = np.asarray(hist, dtype=np.uint64)
    counts = hist[:65536].astype(np.int64)
    n_missing = int(hist[65536])
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e_all = _e_all()
        present = counts > 0
        if n_missing > 0:
            present = present.copy()
            present[0x3C00] = True  # missing cells carry raw = 1.0 until they are overwritten with E = 1
        m = np.nanmax(e_all[present]) if present.any() else np.float64(np.nan)
        elut = e_all / m
    n_nan = int(counts[np.isnan(_ALL_HALF)].sum())
    vals = np.concatenate([elut, [1.0]])
    cnts = np.concatenate([counts, [n_missing]])
    keep = cnts > 0
    test_p = weighted_quantile_linear(vals[keep], cnts[keep], init_filter_p) if keep.any() else np.float64(np.nan)
    return elut, test_p, n_nan


# ---------------------------------------------------------------------------------------------------
# engine
# ---------------------------------------------------------------------------------------------------
class Engine(object):
    """One schedule + device buffers; `correct()` is the host-buffer (end-to-end) call,
    the `*_device` methods work on resident data."""

    def __init__(self, pedigree, ids, back=2, schedule=None):
        self.torch = _torch()
        self.L = _lib.lib()
        self.schedule = schedule if schedule is not None else Schedule(pedigree, ids, back)
        self.n = self.schedule.n_rows
        self.back = int(back)
        self._shape = None
        self.stats = {}

    # ---- buffers -----------------------------------------------------------------------------
    def _alloc(self, s):
        if self._shape == s:
            return
        t = self.torch
        dev = t.device("cuda", t.cuda.current_device())
        n = self.n
        self.s = int(s)
        self.nwords = (self.s + 15) // 16
        self.wpr = ((self.nwords + 3) // 4) * 4
        self.ld_raw = 16 * self.wpr
        self.ldc = ((self.s + 15) // 16) * 16
        self.codes = t.empty((n, self.ldc), dtype=t.uint8, device=dev)
        self.packed_orig = t.empty((n, self.wpr), dtype=t.int32, device=dev)
        self.packed_live = t.empty((n, self.wpr), dtype=t.int32, device=dev)
        self.raw = t.empty((n, self.ld_raw), dtype=t.int16, device=dev)
        self.hist = t.empty(65537, dtype=t.int64, device=dev)
        self.elut = t.empty(65536, dtype=t.float64, device=dev)
        # snapshot + fallback bitmap + lists of parked cells / deferred jobs (~76 bytes per parked cell)
        # (>= 64 MB of lists: with many-mate sires a small matrix parks most of its suspect cells, and a cell that finds no
        # slot is redone by the slow whole-cell kernel -- 90 % of the step on the reference's example pedigree before)
        self.snapshot = t.empty(n * (self.wpr + 32 * ((self.s + 1023) // 1024)) + 64 + 4 * n + max(1 << 24, n * self.s // 4),
                                dtype=t.int32, device=dev)
        self.tallies = t.zeros(2 * n, dtype=t.int32, device=dev)
        self.host_in = t.empty((n, self.ldc), dtype=t.uint8, pin_memory=True)
        self.host_out = t.empty((n, self.ldc), dtype=t.uint8, pin_memory=True)
        self.host_hist = t.empty(65537, dtype=t.int64, pin_memory=True)
        self.host_elut = t.empty(65536, dtype=t.float64, pin_memory=True)
        self.cutraw = t.empty(65537, dtype=t.int16, device=dev)
        self.host_cutraw = t.empty(65537, dtype=t.int16, pin_memory=True)
        self._shape = s

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    # ---- resident-data steps -------------------------------------------------------------------
    def load_codes_device(self, codes_dev):
        """codes_dev: uint8 cuda tensor [n, >=s] (values 0,1,2,9)."""
        _lib.check(self.L.gfx_pack2(codes_dev.data_ptr(), self.n, self.s, codes_dev.stride(0), self.packed_orig.data_ptr(),
                                    self.wpr, self._stream()))

    def use_packed(self, packed_dev, n_snps):
        """Adopt an already packed device matrix [n, wpr] int32 (wpr as computed by `_alloc(n_snps)`)."""
        self._alloc(n_snps)
        if tuple(packed_dev.shape) != (self.n, self.wpr):
            raise ValueError("packed matrix must be [%d, %d]" % (self.n, self.wpr))
        self.packed_orig = packed_dev

    def correct_resident(self, threshold_pair, threshold_single, init_filter_p=0.9, tiethreshold=0.05, err_thresh=1):
        """Whole path on resident data: rank -> histogram -> thresholds -> pairs -> singles.
        Result stays in self.packed_live; per-animal tallies in self.tallies."""
        self.mendel_rank_device()
        self.thresholds_device(init_filter_p)
        self.correct_device(threshold_pair, threshold_single, tiethreshold, err_thresh)

    def mendel_rank_device(self, fused=True):
        """Scores of every cell.  fused (the pipeline's form): one pass that also builds the histogram and
        puts the 0xFFFF marker on missing cells; otherwise the plain scores (thresholds_device then runs
        gfx_rank_hist as a second pass)."""
        if fused:
            _lib.check(self.L.gfx_mendel_rank_hist(self.schedule.handle, self.packed_orig.data_ptr(), self.s, self.wpr,
                                                   self.raw.data_ptr(), self.ld_raw, self.hist.data_ptr(), self._stream()))
        else:
            _lib.check(self.L.gfx_mendel_rank(self.schedule.handle, self.packed_orig.data_ptr(), self.s, self.wpr,
                                              self.raw.data_ptr(), self.ld_raw, self._stream()))
        self._hist_ready = bool(fused)

    def thresholds_device(self, init_filter_p):
        if not getattr(self, "_hist_ready", False):
            _lib.check(self.L.gfx_rank_hist(self.raw.data_ptr(), self.ld_raw, self.packed_orig.data_ptr(), self.wpr, self.n,
                                            self.s, self.hist.data_ptr(), self._stream()))
        self._hist_ready = False
        self.host_hist.copy_(self.hist, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        hist = self.host_hist.numpy().view(np.uint64)
        # the next chunks of this matrix: k_mendel_rank3 needs words without missing calls for its fast paths; with more than
        # 0.5 % missing cells (8 % of the words) the second form is the leaner kernel (gfx_rank_form)
        form = 2 if int(hist[65536]) * 200 > self.n * self.s else 0
        if form != getattr(self.schedule, "_rank_form", 0):
            _lib.check(self.L.gfx_rank_form(self.schedule.handle, form))
            self.schedule._rank_form = form
        elut, test_p, n_nan = rank_transform_table(hist, init_filter_p)
        self.host_elut.numpy()[:] = elut
        self.elut.copy_(self.host_elut, non_blocking=True)
        tp_raw = C.c_uint32(0)
        _lib.check(self.L.gfx_build_cut_table(self.host_elut.data_ptr(), float(test_p), SUSPECT_T,
                                              self.host_cutraw.data_ptr(), C.byref(tp_raw)))
        self.cutraw.copy_(self.host_cutraw, non_blocking=True)
        self.tp_raw = int(tp_raw.value)
        self.test_p = float(test_p)
        self.n_nan = n_nan
        self.elut_host = elut
        return self.test_p

    def correct_device(self, threshold_pair, threshold_single, tiethreshold, err_thresh, keep_posteriors=False):
        t = self.torch
        self.packed_live.copy_(self.packed_orig)
        self.tallies.zero_()
        posts = {"pairs": [], "singles": None}
        if not np.isnan(self.test_p):  # D16: NaN threshold => nothing is suspect
            info = self.schedule.info
            pp = _lib.CorrectParams(self.test_p, SUSPECT_T, float(tiethreshold), float(threshold_pair), float(err_thresh),
                                    self.cutraw.data_ptr(), self.tp_raw, 0)
            gen_off = self.schedule.array(1) if keep_posteriors else None
            for g in range(info["n_generations"]):
                post = None
                if keep_posteriors:
                    nj = int(gen_off[g + 1] - gen_off[g])
                    post = t.full((max(nj, 1) * 2, self.s, 3), POST_SENTINEL, dtype=t.float64, device=self.codes.device)
                _lib.check(self.L.gfx_correct_pairs(
                    self.schedule.handle, g, self.packed_live.data_ptr(), self.s, self.wpr,
                    self.raw.data_ptr(), self.ld_raw, self.elut.data_ptr(), C.byref(pp), self.snapshot.data_ptr(), self.snapshot.numel(),
                    post.data_ptr() if post is not None else None, self.tallies.data_ptr(), self._stream()))
                if keep_posteriors:
                    ph = post.cpu().numpy()
                    posts["pairs"].append((ph, ph[:, :, 0] != POST_SENTINEL))
            ps = _lib.CorrectParams(self.test_p, SUSPECT_T, float(tiethreshold), float(threshold_single), float(err_thresh),
                                    self.cutraw.data_ptr(), self.tp_raw, 0)
            post = None
            ns = info["n_singles"]
            if keep_posteriors:
                post = t.full((max(ns, 1), self.s, 3), POST_SENTINEL, dtype=t.float64, device=self.codes.device)
            _lib.check(self.L.gfx_correct_singles(
                self.schedule.handle, self.packed_live.data_ptr(), self.s, self.wpr,
                self.raw.data_ptr(), self.ld_raw, self.elut.data_ptr(), C.byref(ps), self.snapshot.data_ptr(), self.snapshot.numel(),
                post.data_ptr() if post is not None else None, self.tallies.data_ptr(), self._stream()))
            if keep_posteriors:
                ph = post.cpu().numpy()
                posts["singles"] = (ph, ph[:, :, 0] != POST_SENTINEL)
        return posts

    def unpack_device(self, out_codes_dev):
        _lib.check(self.L.gfx_unpack2(self.packed_live.data_ptr(), self.n, self.s, self.wpr, out_codes_dev.data_ptr(),
                                      out_codes_dev.stride(0), self._stream()))

    def check_status(self):
        _lib.check(self.L.gfx_check_device_status(self.schedule.handle, self._stream()))

    def raw_host(self):
        """float16 [n, s] rank matrix (what mendelProbsSingle returns per animal).  gfx_rank_hist marks
        missing cells with 0xFFFF; their score is 1.0 (correct_genotypes3.py:102)."""
        r = self.raw[:, :self.s].cpu().numpy().view(np.uint16).copy()
        r[r == 0xFFFF] = 0x3C00
        return r.view(np.float16)

    def E_host(self):
        """float64 [n, s] transformed ranks, missing cells of the input = 1 (correct_genotypes3.py:600-603)."""
        raw = self.raw[:, :self.s].cpu().numpy().view(np.uint16)
        E = self.elut_host[np.where(raw == 0xFFFF, 0x3C00, raw)]
        p = self.packed_orig.cpu().numpy().view(np.uint32)
        j = np.arange(self.s)
        codes = (p[:, j // 16] >> (2 * (j % 16)).astype(np.uint32)) & 3
        E[codes == 3] = 1.0
        return E

    # ---- correctMatrix with an LD index (correct_genotypes3.py:619-667, :746-779, :883-908) ---------------------
    def emp_index_device(self, wstart, wlen, freq, weight):
        """Device copy of an LD index for the s columns of the loaded chunk: per column its window [wstart, wstart + wlen) and
        its 3^wlen stored frequencies (float64, itertools.product order)."""
        t = self.torch
        dev = self.codes.device
        wstart = np.ascontiguousarray(wstart, dtype=np.int32)
        wlen = np.ascontiguousarray(wlen, dtype=np.int32)
        if len(wstart) != self.s or len(wlen) != self.s:
            raise ValueError("the LD index must cover every column of the matrix")
        if wlen.max(initial=0) > 9:
            raise ValueError("LD windows of more than 9 SNPs are not supported")
        sizes = np.array([3 ** int(w) if w > 0 else 0 for w in wlen], dtype=np.int64)
        foff = np.zeros(self.s, dtype=np.int64)
        foff[1:] = np.cumsum(sizes)[:-1]
        flat = np.zeros(max(int(sizes.sum()), 1), dtype=np.float64)
        for j in range(self.s):
            if sizes[j]:
                f = np.asarray(freq[j], dtype=np.float64).reshape(-1)
                flat[foff[j]:foff[j] + sizes[j]] = f[:sizes[j]]
        keep = dict(wstart=t.from_numpy(wstart).to(dev), wlen=t.from_numpy(wlen).to(dev), foff=t.from_numpy(foff).to(dev),
                    freq=t.from_numpy(flat).to(dev))
        idx = _lib.EmpIndex(keep["wstart"].data_ptr(), keep["wlen"].data_ptr(), keep["foff"].data_ptr(), keep["freq"].data_ptr(),
                            float(weight), int(wlen.max(initial=0)), 0)
        return idx, keep

    def correct_emp(self, G, index, threshold_pair, threshold_single, init_filter_p=0.9, tiethreshold=0.05, err_thresh=1):
        """correctMatrix with an LD index.  index = (wstart, wlen, freq, weight) over the columns of G.  The ranks are blended
        with the window frequencies on the device (gfx_emp_rank_blend), re-normalised and re-coded as dense ranks of their
        distinct values (the correction kernels compare 16-bit codes), the posteriors of every generation come from the
        inference kernels run on a scratch copy, and gfx_emp_apply walks each focal animal's jobs and SNPs in order."""
        t = self.torch
        G = np.asarray(G)
        if G.ndim != 2 or G.shape[0] != self.n:
            raise ValueError("genotype matrix must be [%d, n_snps]" % self.n)
        _lib.check(self.L.gfx_reset_device_status(self.schedule.handle))
        self._alloc(G.shape[1])
        hin = self.host_in.numpy()
        _parallel_copy(hin[:, :self.s], G)
        if self.ldc > self.s:
            hin[:, self.s:] = 9
        self.codes.copy_(self.host_in, non_blocking=True)
        self.load_codes_device(self.codes)
        self.mendel_rank_device()
        self.thresholds_device(init_filter_p)                      # elut of the Mendelian ranks (:593-603)
        idx, keep = self.emp_index_device(*index)
        dev = self.codes.device
        E = t.empty((self.n, self.s), dtype=t.float64, device=dev)
        _lib.check(self.L.gfx_emp_rank_blend(self.packed_orig.data_ptr(), self.n, self.s, self.wpr, self.raw.data_ptr(), self.ld_raw,
                                             self.elut.data_ptr(), C.byref(idx), E.data_ptr(), self.s, self._stream()))
        missing = self.codes[:, :self.s] > 2
        nan = t.isnan(E)
        mx = t.where(nan, t.full_like(E, -np.inf), E).max()         # np.nanmax (:658)
        E = E / mx
        E[missing] = 1.0                                             # :660
        flat = E.flatten()
        self.packed_live.copy_(self.packed_orig)
        self.tallies.zero_()
        if bool(t.isnan(flat).any()):
            self.test_p = float("nan")                               # D16: a NaN rank poisons the quantile, nothing is suspect
        else:
            srt = t.sort(flat).values
            pos = init_filter_p * (flat.numel() - 1)                 # np.quantile, method 'linear' (:667)
            lo = int(np.floor(pos))
            hi = min(lo + 1, flat.numel() - 1)
            ab = srt[[lo, hi]].cpu().numpy()
            self.test_p = float(_lerp(ab[0], ab[1], pos - lo))
            vals, inv = t.unique(flat, sorted=True, return_inverse=True)
            if vals.numel() > 0x7C01:
                raise ValueError("more than 31745 distinct blended rank values in one matrix: correct it in narrower column chunks")
            elut = np.full(65536, np.nan)
            elut[:vals.numel()] = vals.cpu().numpy()
            self.host_elut.numpy()[:] = elut
            self.elut.copy_(self.host_elut)
            self.raw[:, :self.s] = inv.reshape(self.n, self.s).to(t.int16)
            tp_raw = C.c_uint32(0)
            _lib.check(self.L.gfx_build_cut_table(self.host_elut.data_ptr(), self.test_p, SUSPECT_T, self.host_cutraw.data_ptr(),
                                                  C.byref(tp_raw)))
            self.cutraw.copy_(self.host_cutraw)
            self.tp_raw = int(tp_raw.value)
            self.elut_host = elut
            info = self.schedule.info
            gen_off = self.schedule.array(1)
            scratch_live = t.empty_like(self.packed_live)
            scratch_tallies = t.zeros_like(self.tallies)
            for g in range(info["n_generations"] + 1):
                pairs = g < info["n_generations"]
                rows = int(gen_off[g + 1] - gen_off[g]) * 2 if pairs else info["n_singles"]
                if rows <= 0:
                    continue
                pp = _lib.CorrectParams(self.test_p, SUSPECT_T, float(tiethreshold), float(threshold_pair if pairs else threshold_single),
                                        float(err_thresh), self.cutraw.data_ptr(), self.tp_raw, 0)
                post = t.full((rows, self.s, 3), POST_SENTINEL, dtype=t.float64, device=dev)
                scratch_live.copy_(self.packed_live)
                if pairs:
                    _lib.check(self.L.gfx_correct_pairs(
                        self.schedule.handle, g, scratch_live.data_ptr(), self.s, self.wpr, self.raw.data_ptr(), self.ld_raw,
                        self.elut.data_ptr(), C.byref(pp), self.snapshot.data_ptr(), self.snapshot.numel(), post.data_ptr(),
                        scratch_tallies.data_ptr(), self._stream()))
                else:
                    _lib.check(self.L.gfx_correct_singles(
                        self.schedule.handle, scratch_live.data_ptr(), self.s, self.wpr, self.raw.data_ptr(), self.ld_raw,
                        self.elut.data_ptr(), C.byref(pp), self.snapshot.data_ptr(), self.snapshot.numel(), post.data_ptr(),
                        scratch_tallies.data_ptr(), self._stream()))
                # the snapshot region of the scratch buffer holds the matrix at the start of the generation
                _lib.check(self.L.gfx_emp_apply(self.schedule.handle, g if pairs else -1, self.packed_live.data_ptr(),
                                                self.snapshot.data_ptr(), self.s, self.wpr, self.raw.data_ptr(), self.ld_raw,
                                                self.elut.data_ptr(), C.byref(idx), C.byref(pp), post.data_ptr(), POST_SENTINEL,
                                                self.tallies.data_ptr(), self._stream()))
        self.unpack_device(self.codes)
        self.host_out.copy_(self.codes, non_blocking=True)
        self.check_status()
        tall = self.tallies.cpu().numpy()
        self.n_nan = int(nan.sum())
        self.stats = dict(test_p=self.test_p, n_nan=self.n_nan, corrected_per_animal=tall[:self.n].copy(),
                          excluded_per_animal=tall[self.n:].copy(), posteriors=None)
        out = np.empty((self.n, self.s), dtype=np.uint8)
        _parallel_copy(out, self.host_out.numpy()[:, :self.s])
        del keep
        return out

    # ---- end-to-end on host buffers ----------------------------------------------------------------
    def correct(self, G, threshold_pair, threshold_single, init_filter_p=0.9, tiethreshold=0.05, err_thresh=1,
                keep_posteriors=False):
        """G: uint8 [n, s] numpy (0,1,2,9).  Returns the corrected uint8 [n, s] matrix (new array)."""
        G = np.asarray(G)
        if G.ndim != 2 or G.shape[0] != self.n:
            raise ValueError("genotype matrix must be [%d, n_snps]" % self.n)
        _lib.check(self.L.gfx_reset_device_status(self.schedule.handle))
        self._alloc(G.shape[1])
        hin = self.host_in.numpy()
        _parallel_copy(hin[:, :self.s], G)
        if self.ldc > self.s:
            hin[:, self.s:] = 9
        self.codes.copy_(self.host_in, non_blocking=True)
        self.load_codes_device(self.codes)
        self.mendel_rank_device()
        self.thresholds_device(init_filter_p)
        posts = self.correct_device(threshold_pair, threshold_single, tiethreshold, err_thresh, keep_posteriors)
        self.unpack_device(self.codes)
        self.host_out.copy_(self.codes, non_blocking=True)
        self.check_status()  # synchronises the stream
        tall = self.tallies.cpu().numpy()
        self.stats = dict(test_p=self.test_p, n_nan=self.n_nan, corrected_per_animal=tall[:self.n].copy(),
                          excluded_per_animal=tall[self.n:].copy(), posteriors=posts)
        out = np.empty((self.n, self.s), dtype=np.uint8)
        _parallel_copy(out, self.host_out.numpy()[:, :self.s])
        return out

    # ---- end-to-end on a sequence of chunks (host buffers), copies overlapped with compute -----------------
    def _alloc_pipe(self):
        if getattr(self, "_pipe_shape", None) == self._shape:
            return
        t = self.torch
        dev = self.codes.device
        n = self.n
        self.p_in = [t.empty((n, self.ldc), dtype=t.uint8, pin_memory=True) for _ in range(2)]
        self.p_out = [t.empty((n, self.ldc), dtype=t.uint8, pin_memory=True) for _ in range(2)]
        self.p_codes_in = [t.full((n, self.ldc), 9, dtype=t.uint8, device=dev) for _ in range(2)]   # pad columns stay 9
        self.p_codes_out = [t.empty((n, self.ldc), dtype=t.uint8, device=dev) for _ in range(2)]
        self.p_packed = [t.empty((n, self.wpr), dtype=t.int32, device=dev) for _ in range(2)]
        self.p_packed_out = [t.empty((n, self.wpr), dtype=t.int32, device=dev) for _ in range(2)]
        self.p_pin_packed_in = self.p_pin_packed_out = None     # pinned staging for packed I/O, allocated on first use
        self.p_tall_dev = [t.empty(2 * n, dtype=t.int32, device=dev) for _ in range(2)]
        self.p_tall = [t.empty(2 * n, dtype=t.int32, pin_memory=True) for _ in range(2)]
        if not hasattr(self, "s_in"):
            self.s_in, self.s_c, self.s_out = t.cuda.Stream(), t.cuda.Stream(), t.cuda.Stream()
        self._pipe_shape = self._shape

    def correct_chunks(self, chunks, threshold_pair, threshold_single, init_filter_p=0.9, tiethreshold=0.05, err_thresh=1,
                       outs=None, packed_io=False):
        """Generator over (corrected uint8 [n, s_i], stats) for an iterable of uint8 [n, s_i] chunks -- the chunk loop of
        genofix.py:207-237 with every chunk an independent correctMatrix call.  Same results as calling correct()
        chunk by chunk; here three streams and two buffer sets overlap a chunk's kernels with the host staging +
        H2D copy of the next chunk and the D2H copy + un-staging of the previous one.  `outs`: optional list of
        caller-owned uint8 [n, s_i] arrays to receive the results.  Chunks and outs that live in pinned host memory
        (pinned_array()) are read / written by the copy engines directly; anything else is staged through pinned
        buffers with a threaded host copy.

        packed_io=True: every chunk is a pair (P, n_snps) with P the 2-bit packed block uint32 [n, wpr] in the kernels'
        own layout (gfutils.packedgens.PackedGens.packed_columns / pack_codes) and the results are packed blocks of
        the same shape: a quarter of the bytes cross PCIe and the pack / unpack kernels are skipped."""
        t = self.torch
        if not getattr(self, "_lane", False):       # a lane of MultiLane: the owner has reset the shared status word
            _lib.check(self.L.gfx_reset_device_status(self.schedule.handle))
        it = iter(chunks)
        dt_io = np.uint32 if packed_io else np.uint8

        def width(item):
            return int(item[1]) if packed_io else np.asarray(item).shape[1]

        def io_shape(s_):
            return (self.n, ((((s_ + 15) // 16) + 3) // 4) * 4) if packed_io else (self.n, s_)
        state = {"ev_comp": [None, None], "ev_out": [None, None], "hold": [None, None], "staged": 0}

        def pinned_view(arr, s):
            """torch view of a caller array the copy engines can reach directly, else None (staged copy)"""
            if not (isinstance(arr, np.ndarray) and arr.dtype == dt_io and arr.flags.c_contiguous and arr.flags.writeable
                    and arr.shape == io_shape(s)):
                return None
            tv = t.from_numpy(arr.view(np.int32) if packed_io else arr)
            return tv if tv.is_pinned() else None

        def stage(item, slot):
            s_new = width(item)
            G = np.asarray(item[0] if packed_io else item)
            if G.ndim != 2 or G.shape[0] != self.n or (packed_io and (G.dtype != np.uint32 or G.shape != io_shape(s_new))):
                raise ValueError("chunk must be %s %s" % ("uint32" if packed_io else "uint8", (io_shape(s_new),)))
            if self._shape != s_new:
                t.cuda.synchronize()          # chunk width changed (last, ragged chunk): drain, then re-allocate
                self._alloc(s_new)
                state["ev_comp"] = [None, None]
                state["ev_out"] = [None, None]
            self._alloc_pipe()
            src = pinned_view(G, self.s)
            if src is None:
                if packed_io:
                    if self.p_pin_packed_in is None:
                        self.p_pin_packed_in = [t.empty((self.n, self.wpr), dtype=t.int32, pin_memory=True) for _ in range(2)]
                    _parallel_copy(self.p_pin_packed_in[slot].numpy().view(np.uint32), G)
                else:
                    _parallel_copy(self.p_in[slot].numpy()[:, :self.s], G)
            state["hold"][slot] = (G, src)    # keep the caller's memory alive until the copy has run
            order = getattr(self, "_copy_order", None)   # a lane of MultiLane: (CopyOrder, global chunk numbers of this lane)
            kglob = order[1][state["staged"]] if order is not None else -1
            state["staged"] += 1
            before = order[0].wait_for(kglob - 1) if order is not None else None
            with t.cuda.stream(self.s_in):
                if before is not None:
                    self.s_in.wait_event(before)                     # uploads in chunk order (CopyOrder)
                if state["ev_comp"][slot] is not None:
                    self.s_in.wait_event(state["ev_comp"][slot])     # the kernels that read p_packed[slot] are done
                if packed_io:
                    self.p_packed[slot].copy_(src if src is not None else self.p_pin_packed_in[slot], non_blocking=True)
                else:
                    self.p_codes_in[slot][:, :self.s].copy_(src if src is not None else self.p_in[slot][:, :self.s], non_blocking=True)
                    _lib.check(self.L.gfx_pack2(self.p_codes_in[slot].data_ptr(), self.n, self.s, self.p_codes_in[slot].stride(0),
                                                self.p_packed[slot].data_ptr(), self.wpr, self._stream()))
                ev = t.cuda.Event()
                ev.record(self.s_in)
            if order is not None:
                order[0].publish(kglob, ev)
            return ev

        def finish(slot, s, test_p, n_nan, k, out, direct):
            state["ev_out"][slot].synchronize()
            _lib.check(self.L.gfx_check_device_status(self.schedule.handle, C.c_void_p(self.s_out.cuda_stream)))
            tall = self.p_tall[slot].numpy()
            stats = dict(test_p=test_p, n_nan=n_nan, corrected_per_animal=tall[:self.n].copy(),
                         excluded_per_animal=tall[self.n:].copy(), posteriors=None)
            if not direct:
                if out is None:
                    out = np.empty(io_shape(s), dtype=dt_io)
                if packed_io:
                    _parallel_copy(out, self.p_pin_packed_out[slot].numpy().view(np.uint32))
                else:
                    _parallel_copy(out, self.p_out[slot].numpy()[:, :s])
            return out, stats

        cur = next(it, None)
        if cur is None:
            return
        slot = 0
        k = 0
        ev_in = stage(cur, slot)
        prev = None
        while cur is not None:
            with t.cuda.stream(self.s_c):
                self.s_c.wait_event(ev_in)
                self.packed_orig = self.p_packed[slot]
                self.mendel_rank_device()
                self.thresholds_device(init_filter_p)           # host round trip: synchronises s_c
                self.correct_device(threshold_pair, threshold_single, tiethreshold, err_thresh)
                if state["ev_out"][slot] is not None:
                    self.s_c.wait_event(state["ev_out"][slot])   # the D2H that read p_codes_out[slot] is done
                if packed_io:
                    self.p_packed_out[slot].copy_(self.packed_live, non_blocking=True)
                else:
                    self.unpack_device(self.p_codes_out[slot])
                self.p_tall_dev[slot].copy_(self.tallies, non_blocking=True)
                state["ev_comp"][slot] = t.cuda.Event()
                state["ev_comp"][slot].record(self.s_c)
            out = outs[k] if outs is not None else None
            if out is not None and (out.shape != io_shape(self.s) or out.dtype != dt_io):
                raise ValueError("outs[%d] must be %s %s" % (k, np.dtype(dt_io).name, (io_shape(self.s),)))
            dst = pinned_view(out, self.s) if out is not None else None
            if packed_io and dst is None and self.p_pin_packed_out is None:
                self.p_pin_packed_out = [t.empty((self.n, self.wpr), dtype=t.int32, pin_memory=True) for _ in range(2)]
            with t.cuda.stream(self.s_out):
                self.s_out.wait_event(state["ev_comp"][slot])
                if packed_io:
                    (dst if dst is not None else self.p_pin_packed_out[slot]).copy_(self.p_packed_out[slot], non_blocking=True)
                else:
                    (dst if dst is not None else self.p_out[slot][:, :self.s]).copy_(self.p_codes_out[slot][:, :self.s], non_blocking=True)
                self.p_tall[slot].copy_(self.p_tall_dev[slot], non_blocking=True)
                state["ev_out"][slot] = t.cuda.Event()
                state["ev_out"][slot].record(self.s_out)
            mine = (slot, self.s, self.test_p, self.n_nan, k, out, dst is not None)
            k += 1
            # the GPU is busy with this chunk's correction: stage the next chunk, hand back the previous one
            nxt = next(it, None)
            if nxt is not None and width(nxt) != self.s and prev is not None:
                yield finish(*prev)       # a re-allocation is coming: nothing may stay in the old buffers
                prev = None
            if nxt is not None and width(nxt) != self.s:
                yield finish(*mine)
                mine = None
            ev_in = stage(nxt, 1 - slot) if nxt is not None else None
            if prev is not None:
                yield finish(*prev)
            prev = mine
            cur = nxt
            slot = 1 - slot
        if prev is not None:
            yield finish(*prev)

    # ---- config 5: trio scan ---------------------------------------------------------------------------
    def trio_scan(self, G=None):
        t = self.torch
        if G is not None:
            G = np.asarray(G)
            self._alloc(G.shape[1])
            hin = self.host_in.numpy()
            hin[:, :self.s] = G
            if self.ldc > self.s:
                hin[:, self.s:] = 9
            self.codes.copy_(self.host_in, non_blocking=True)
            self.load_codes_device(self.codes)
        inc = t.empty(self.n, dtype=t.int64, device=self.codes.device)
        tst = t.empty(self.n, dtype=t.int64, device=self.codes.device)
        _lib.check(self.L.gfx_trio_scan(self.schedule.handle, self.packed_orig.data_ptr(), self.s, self.wpr,
                                        inc.data_ptr(), tst.data_ptr(), self._stream()))
        return inc.cpu().numpy(), tst.cpu().numpy()


    def trio_scan_chunks(self, chunks):
        """The scan of a matrix held as column chunks (the pipeline's resident layout): `chunks` = iterable of
        (packed int32 cuda tensor [n, wpr_i], n_snps_i).  Returns device int64 tensors (inconsistent, tested)."""
        t = self.torch
        dev = t.device("cuda", t.cuda.current_device())
        inc = t.zeros(self.n, dtype=t.int64, device=dev)
        tst = t.zeros(self.n, dtype=t.int64, device=dev)
        for p, s_i in chunks:
            if p.shape[0] != self.n or p.dtype != t.int32 or not p.is_contiguous():
                raise ValueError("chunk must be a contiguous int32 tensor [%d, wpr]" % self.n)
            _lib.check(self.L.gfx_trio_scan_add(self.schedule.handle, p.data_ptr(), int(s_i), p.shape[1], inc.data_ptr(),
                                                tst.data_ptr(), self._stream()))
        return inc, tst

    def load_host(self, G):
        """Upload a uint8 [n, s] host matrix (0,1,2,9) and pack it (the first lines of `correct()`)."""
        G = np.asarray(G)
        if G.ndim != 2 or G.shape[0] != self.n:
            raise ValueError("genotype matrix must be [%d, n_snps]" % self.n)
        self._alloc(G.shape[1])
        hin = self.host_in.numpy()
        _parallel_copy(hin[:, :self.s], G)
        if self.ldc > self.s:
            hin[:, self.s:] = 9
        self.codes.copy_(self.host_in, non_blocking=True)
        self.load_codes_device(self.codes)

    def rank_rowsum_device(self, missing_value=1.0):
        """float64 sum over the SNPs of this matrix of E (the table of thresholds_device; missing cells count
        `missing_value`), per animal, on the device: the partial sums a SNP-sharded run all-reduces
        (`parallel.allreduce_tallies`).  Fixed summation order: repeatable bit for bit."""
        t = self.torch
        out = t.empty(self.n, dtype=t.float64, device=self.raw.device)
        _lib.check(self.L.gfx_rank_rowsum(self.raw.data_ptr(), self.ld_raw, self.packed_orig.data_ptr(), self.wpr, self.n,
                                          self.s, self.elut.data_ptr(), float(missing_value), out.data_ptr(), self._stream()))
        return out

    # ---- per-animal Mendel tally (error_checker.py:193-204, preindex.py:259-266) -----------------------------
    def mendel_tally(self, G, rows=None, missing_value=-1.0, want_matrix=True):
        """Transformed ranks and their per-animal sums in the reference's float16 arithmetic.

        G: uint8 [n, s] (0,1,2,9) or None to use the matrix already on the device; rows: row indices to tally
        (error_checker.py: the animals with both parents genotyped) or None for all.  Returns
        (E16 float16 [len(rows), s] or None, sums float16 [len(rows)], max16) -- the sums are NOT yet divided by their
        nanmax (the caller does, error_checker.py:202): sharded callers need the raw sums."""
        t = self.torch
        if G is not None:
            G = np.asarray(G)
            if G.ndim != 2 or G.shape[0] != self.n:
                raise ValueError("genotype matrix must be [%d, n_snps]" % self.n)
            self._alloc(G.shape[1])
            hin = self.host_in.numpy()
            _parallel_copy(hin[:, :self.s], G)
            if self.ldc > self.s:
                hin[:, self.s:] = 9
            self.codes.copy_(self.host_in, non_blocking=True)
            self.load_codes_device(self.codes)
        self.mendel_rank_device(fused=True)      # scores + histogram, 0xFFFF marker on missing cells
        self._hist_ready = False
        self.host_hist.copy_(self.hist, non_blocking=True)
        t.cuda.current_stream().synchronize()
        self.tally_hist = self.host_hist.numpy().view(np.uint64).copy()      # score patterns present (65536 = missing)
        e16, m16 = half_transform_table(self.tally_hist, missing_value)
        self.tally_table = e16
        dev = self.raw.device
        e16_dev = t.from_numpy(e16.view(np.int16)).to(dev)
        if rows is None:
            rows_np = np.arange(self.n, dtype=np.int32)
            rows_dev = None
        else:
            rows_np = np.ascontiguousarray(rows, dtype=np.int32)
            if rows_np.ndim != 1 or (len(rows_np) and (rows_np.min() < 0 or rows_np.max() >= self.n)):
                raise ValueError("rows out of range")
            rows_dev = t.from_numpy(rows_np).to(dev)
        out = t.empty(max(len(rows_np), 1), dtype=t.int16, device=dev)
        _lib.check(self.L.gfx_rank_rowsum_f16(self.raw.data_ptr(), self.ld_raw, self.n, self.s, e16_dev.data_ptr(),
                                              rows_dev.data_ptr() if rows_dev is not None else None, len(rows_np),
                                              out.data_ptr(), self._stream()))
        sums = out[:len(rows_np)].cpu().numpy().view(np.float16)
        E16 = None
        if want_matrix:
            src = self.raw if rows_dev is None else self.raw.index_select(0, rows_dev.long())
            E16 = e16[src[:, :self.s].cpu().numpy().view(np.uint16)]
        return E16, sums, m16


_ALL_HALF_E16 = None


def half_transform_table(hist, missing_value):
    """float16 E of every score pattern, as error_checker.py:193-197 / preindex.py:259-263 compute it on their
    float16 frame: log(log(x + 1) + 1) / nanmax(.), every step a float16 ufunc.  hist: uint64[65537] of the score
    patterns present (bin 65536 = missing cells, whose score is 1.0 until it is overwritten).  Entry 0xFFFF of the
    table is the value of a missing cell.  Returns (table float16[65536], max float16)."""
    global _ALL_HALF_E16
    hist = np.asarray(hist, dtype=np.uint64)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if _ALL_HALF_E16 is None:
            all16 = np.arange(65536, dtype=np.uint16).view(np.float16)
            _ALL_HALF_E16 = np.log(np.log(all16 + 1) + 1)
            assert _ALL_HALF_E16.dtype == np.float16
        present = hist[:65536] > 0
        if hist[65536] > 0:
            present = present.copy()
            present[0x3C00] = True
        m = np.nanmax(_ALL_HALF_E16[present]) if present.any() else np.float16(np.nan)
        table = _ALL_HALF_E16 / m
    table[0xFFFF] = np.float16(missing_value)
    return table, np.float16(m)


def pinned_array(shape, dtype=np.uint8):
    """numpy array in page-locked host memory: correct_chunks moves such arrays with the copy engines directly
    (no staging copy).  The array keeps its backing tensor alive."""
    t = _torch()
    dtype = np.dtype(dtype)
    if dtype == np.uint32:     # torch's unsigned types are limited: allocate int32, hand out a uint32 view
        return t.empty(tuple(shape), dtype=t.int32, pin_memory=True).numpy().view(np.uint32)
    return t.empty(tuple(shape), dtype=getattr(t, dtype.name), pin_memory=True).numpy()


class CopyOrder(object):
    """Host-to-device copies of MultiLane's chunks in chunk order.  Every lane stages its chunks on its own stream; without
    an order the first L uploads share the link and ALL arrive after L transfer times, with the GPU idle until then.  With
    it chunk k's upload waits (on the device) for chunk k - 1's, so chunk 0 is complete after one transfer time and its
    kernels run under the uploads of chunks 1 .. L - 1.  publish() / wait_for() are host-side; the ordering itself is a
    stream-wait on the previous upload's event."""

    def __init__(self, n_chunks):
        import threading
        self._cv = threading.Condition()
        self._events = [None] * n_chunks
        self._published = [False] * n_chunks
        self._failed = False

    def wait_for(self, k):
        """Event of chunk k's upload once its lane has enqueued it (None: no such chunk, or a lane failed)."""
        if k < 0 or k >= len(self._events):
            return None
        with self._cv:
            while not self._published[k] and not self._failed:
                self._cv.wait(timeout=0.05)
            return self._events[k]

    def publish(self, k, event):
        with self._cv:
            self._events[k] = event
            self._published[k] = True
            self._cv.notify_all()

    def fail(self):
        with self._cv:
            self._failed = True
            self._cv.notify_all()


class MultiLane(object):
    """L engines on one schedule, one host thread and one set of streams each: independent chunks are corrected
    concurrently, so that the latency-bound correction kernels of one chunk share the GPU with those of another and
    the host round trip of the thresholds is hidden.  Results do not depend on L (chunks are independent)."""

    def __init__(self, pedigree, ids, back=2, lanes=2):
        first = Engine(pedigree, ids, back)
        self.engines = [first] + [Engine(pedigree, ids, back, schedule=first.schedule) for _ in range(lanes - 1)]
        self.schedule = first.schedule
        self.n = first.n
        self.device = first.torch.cuda.current_device()

    def _run_lanes(self, fn):
        """fn(lane index, engine) in one thread per lane; re-raises the first exception."""
        import threading
        errs = []

        def wrap(l):
            try:
                self.engines[0].torch.cuda.set_device(self.device)
                fn(l, self.engines[l])
            except BaseException as e:  # noqa: B902 -- re-raised in the caller's thread
                errs.append(e)
        th = [threading.Thread(target=wrap, args=(l,)) for l in range(len(self.engines))]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if errs:
            raise errs[0]

    def correct_chunks(self, chunks, threshold_pair, threshold_single, init_filter_p=0.9, tiethreshold=0.05, err_thresh=1,
                       outs=None, packed_io=False):
        """Same contract as Engine.correct_chunks (results in input order); chunk k runs on lane k mod L."""
        import queue
        import threading
        L = len(self.engines)
        chunks = list(chunks)
        results = [None] * len(chunks)
        done = [threading.Event() for _ in chunks]
        errs = []
        _lib.check(self.engines[0].L.gfx_reset_device_status(self.schedule.handle))
        order = CopyOrder(len(chunks))
        for eng in self.engines:
            eng._lane = True

        def lane(l, eng):
            mine = list(range(l, len(chunks), L))
            eng._copy_order = (order, mine)
            try:
                gen = eng.correct_chunks((chunks[k] for k in mine), threshold_pair, threshold_single, init_filter_p=init_filter_p,
                                         tiethreshold=tiethreshold, err_thresh=err_thresh,
                                         outs=[outs[k] for k in mine] if outs is not None else None, packed_io=packed_io)
                for k, r in zip(mine, gen):
                    results[k] = r
                    done[k].set()
            except BaseException as e:  # noqa: B902
                errs.append(e)
                order.fail()
                for k in mine:
                    done[k].set()
            finally:
                eng._copy_order = None
        th = [threading.Thread(target=lambda l=l: (self.engines[0].torch.cuda.set_device(self.device), lane(l, self.engines[l])))
              for l in range(L)]
        for x in th:
            x.start()
        try:
            for k in range(len(chunks)):
                done[k].wait()
                if errs:
                    raise errs[0]
                yield results[k]
                results[k] = None
        finally:
            for x in th:
                x.join()


_ENGINES = {}


def engine_for(pedigree, ids, back):
    """Engines are cached per (pedigree object, id list, back): building the schedule is the
    per-pedigree set-up cost, the matrix chunks then stream through it."""
    ids = np.asarray(ids, dtype=np.int64)
    key = (id(pedigree), ids.tobytes(), int(back))
    e = _ENGINES.get(key)
    if e is None:
        if len(_ENGINES) > 4:
            _ENGINES.clear()
        e = Engine(pedigree, ids, back)
        _ENGINES[key] = e
        e._pedigree_ref = pedigree  # keep id() stable
    return e
