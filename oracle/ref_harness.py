"""Harness that runs the UNMODIFIED reference (/root/reference/src) in this container.

TEST INFRASTRUCTURE ONLY: used by oracle/make_golden.py to generate tests/golden/*.npz and by
local cross-checks.  /root/reference does not exist on the GPU box, so nothing that runs there
may import this module.  The product (genofix_b200/) never imports anything under oracle/.

What is patched, and why:
  * sys.path: oracle/shim first (pgmpy / matplotlib / seaborn / mgzip stand-ins), then the
    reference's src directory;
  * numpy aliases removed in numpy>=1.24 / 2.0 that the reference still uses: np.int, np.product
    (src/pedigree/pedigree_dag.py:222,234; src/correct_genotypes3.py:946);
  * D15: correctSingle stores beststate as [[...]] and the apply loop does set(maxstates)
    (src/correct_genotypes3.py:425,934) -> TypeError.  The harness flattens it to the 1-D tie set
    (what the pair branch :743-744 uses).  Applied by wrapping the static method, never by editing
    the reference;
  * D10: results are applied in `as_completed` order, which is nondeterministic.  The harness
    swaps the process pool for an in-process executor that computes every job at submit time
    (so workers still see the generation-start snapshot, src/correct_genotypes3.py:42-50) and
    yields results in ascending (sire, dam) / kid order.  This FIXES an under-specified order; it
    does not change any arithmetic.
"""
import concurrent.futures
import contextlib
import io
import multiprocessing
import os
import sys
import tempfile

import numpy as np

REF_SRC = "/root/reference/src"
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")

_loaded = {}


def available():
    return os.path.isdir(REF_SRC)


class _DoneFuture(object):
    def __init__(self, fn, args, kwargs):
        self.fn_name = getattr(fn, "__name__", str(fn))
        self.args = args
        self.kwargs = kwargs
        self._exc = None
        self._res = None
        try:
            self._res = fn(*args, **kwargs)
        except Exception as e:  # surfaced through .exception() like a real future
            self._exc = e

    def exception(self):
        return self._exc

    def result(self):
        if self._exc is not None:
            raise self._exc
        return self._res


class InProcessExecutor(object):
    """Drop-in for ProcessPoolExecutor(max_workers, initializer, initargs): runs in this process."""
    log = None  # optional list collecting (fn_name, args, result)

    def __init__(self, max_workers=None, initializer=None, initargs=()):
        if initializer is not None:
            initializer(*initargs)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def submit(self, fn, *args, **kwargs):
        f = _DoneFuture(fn, args, kwargs)
        if InProcessExecutor.log is not None:
            InProcessExecutor.log.append((f.fn_name, args, kwargs, f._res))
        return f


def _ordered_as_completed(futures):
    fs = list(futures)

    def key(f):
        return tuple(int(a) for a in f.args if isinstance(a, (int, np.integer)))
    return iter(sorted(fs, key=key))


def load():
    """Import the reference modules under the shim; returns a namespace dict."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference not present at %s" % REF_SRC)
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    if REF_SRC not in sys.path:
        sys.path.insert(1, REF_SRC)
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "product"):
        np.product = np.prod
    import correct_genotypes3 as cg3
    import bayesnet.model as model
    import pedigree.pedigree_dag as pdag
    import empirical.jalleledist3 as jad3

    orig_single = cg3.CorrectGenotypes.correctSingle

    def correctSingle_d15(*a, **k):
        r = orig_single(*a, **k)
        if r is not None:
            for snp in list(r.beststate.keys()):
                bs = r.beststate[snp]
                if len(bs) == 1 and isinstance(bs[0], list):
                    r.beststate[snp] = np.array(bs[0], dtype=int)
        return r
    correctSingle_d15.__name__ = "correctSingle"
    cg3.CorrectGenotypes.correctSingle = staticmethod(correctSingle_d15)
    _loaded.update(cg3=cg3, model=model, pdag=pdag, jad3=jad3)
    return _loaded


@contextlib.contextmanager
def deterministic_pools(log=None):
    old_pool = concurrent.futures.ProcessPoolExecutor
    old_ac = concurrent.futures.as_completed
    concurrent.futures.ProcessPoolExecutor = InProcessExecutor
    concurrent.futures.as_completed = _ordered_as_completed
    InProcessExecutor.log = log
    try:
        yield
    finally:
        concurrent.futures.ProcessPoolExecutor = old_pool
        concurrent.futures.as_completed = old_ac
        InProcessExecutor.log = None


def pedigree_from_rows(rows):
    ref = load()
    return ref["pdag"].PedigreeDAG.from_table(np.asarray(rows, dtype=np.int64))


def run_correct_matrix(genotypes_df, ped_rows, threshold_pair, threshold_single, back=2,
                       init_filter_p=0.9, filter_e=0.7, tiethreshold=0.05, err_thresh=1,
                       quiet=True, emp=None, truth_df=None):
    """Run reference CorrectGenotypes.correctMatrix deterministically.

    emp=None: correctMatrix(empC=None), the configuration that completes at HEAD.
    emp=dict(surround=, pseudocount=, weight_empirical=): the reference builds its own LD index on the same matrix
    (JointAllellicDistribution + countJointFrqAll, unmodified) and blends it into the ranks (:619-656).  The index
    OBJECT gets a `frequency` attribute first: initializerEmp (:34-40) copies `empC.frequency`, which jalleledist3
    objects do not have; no reference source is touched.  The matrix must not hold missing calls
    (getCountTable raises on any 9 in a window).  With an index the reference's debug statistics read the true
    genotypes (`debugreal`, :781-782) whenever DEBUGDIR is set, which Stats() needs: pass truth_df.

    Returns dict(corrected=ndarray uint8, raw=ndarray f16 [N,S], pair_results=list, single_results=list,
    stdout=str[, emp_results=list of (SNP_id, {(kid, observed): 3 floats}), emp_index=the reference's index])."""
    ref = load()
    cg3 = ref["cg3"]
    ped = pedigree_from_rows(ped_rows)
    log = []
    buf = io.StringIO()
    empC = None
    kw = {}
    with tempfile.TemporaryDirectory() as dbg, deterministic_pools(log):
        ctx = contextlib.redirect_stdout(buf) if quiet else contextlib.nullcontext()
        with ctx, np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                if emp is not None:
                    jad3 = ref["jad3"]
                    empC = jad3.JointAllellicDistribution("1", list(genotypes_df.columns), surround_size=emp["surround"],
                                                          pseudocount=emp.get("pseudocount", 0.0001))
                    empC.countJointFrqAll(genotypes_df, mask=None, threads=1)
                    empC.frequency = {}
                    kw["weight_empirical"] = emp.get("weight_empirical", 3)
                c = cg3.CorrectGenotypes(elimination_order="MinNeighbors")
                out = c.correctMatrix(genotypes_df, ped, empC, threshold_pair, threshold_single,
                                      back=back, init_filter_p=init_filter_p, filter_e=filter_e,
                                      tiethreshold=tiethreshold, threads=1, err_thresh=err_thresh,
                                      DEBUGDIR=dbg, debugreal=truth_df if emp is not None else None, **kw)
    ids = list(genotypes_df.index)
    row = {k: i for i, k in enumerate(ids)}
    raw = np.ones(genotypes_df.shape, dtype=np.float16)
    pair_results = []
    single_results = []
    emp_results = []
    for name, args, kwargs, res in log:
        if name == "mendelProbsSingle":
            raw[row[args[0]], :] = np.squeeze(res[1])
        elif name == "correctPair":
            pair_results.append((int(args[0]), int(args[1]), res))
        elif name == "correctSingle":
            single_results.append((int(args[0]), res))
        elif name == "getEmpProbs":
            emp_results.append((args[1], res))
    r = dict(corrected=out.to_numpy().astype(np.uint8), raw=raw, pair_results=pair_results,
             single_results=single_results, stdout=buf.getvalue())
    if emp is not None:
        r.update(emp_results=emp_results, emp_index=empC)
    return r


def run_mendel_probs(genotypes_df, ped_rows, back=2):
    """Reference mendelProbsSingle for every genotyped animal -> raw f16 [N,S] and anc*kids f16 [N,S,3]."""
    ref = load()
    cg3 = ref["cg3"]
    ped = pedigree_from_rows(ped_rows)
    cg3.initializer(genotypes_df, ped, None)
    raw = np.ones(genotypes_df.shape, dtype=np.float16)
    probs = np.zeros(genotypes_df.shape + (3,), dtype=np.float16)
    buf = io.StringIO()
    import warnings
    with contextlib.redirect_stdout(buf), np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, kid in enumerate(genotypes_df.index):
            p, e, _ = cg3.CorrectGenotypes.mendelProbsSingle(kid, back, elimination_order="MinNeighbors")
            raw[i] = np.squeeze(e)
            probs[i] = p
    return raw, probs


def run_error_checker(ped_rows, ids, obs, kids, chrom="chrA", quiet=True):
    """Run the reference's pedigree/error_checker.py main() (:57-204) on files written from the arrays.

    File conventions the reference needs: the header's last name keeps its newline (gensrandaccess.py:27), so a
    padding SNP on another chromosome closes the header; chromosome names must not parse as integers (:119 vs :109).
    Returns dict(with_parents=ids of probs_errors.csv rows, probs_errors=float16 [K, S], sums=float16 [K])."""
    load()
    import pandas as pd
    import pedigree.error_checker as ec
    S = obs.shape[1]
    snps = ["SNP%d" % i for i in range(S)]
    buf = io.StringIO()
    with tempfile.TemporaryDirectory() as d:
        with open(d + "/gens.txt", "w") as f:
            f.write("id " + " ".join(snps) + " pad\n")
            for k, row in zip(ids, obs):
                f.write(str(int(k)) + " " + " ".join(str(int(x)) for x in row) + " 0\n")
        with open(d + "/ped.txt", "w") as f:
            for r in ped_rows:
                f.write(" ".join(str(int(x)) for x in r[:4]) + "\n")
        with open(d + "/kids.txt", "w") as f:
            f.write("kid\n")
            for k in kids:
                f.write("%d\n" % int(k))
        with open(d + "/map.csv", "w") as f:
            f.write("snpid,chrom,pos,topAllele,B\n")
            for i, s in enumerate(snps):
                f.write("%s,%s,%d,A,B\n" % (s, chrom, 100 * i + 1))
            f.write("pad,chrB,1,A,B\n")
        argv0 = sys.argv
        main_mod = sys.modules["__main__"]
        doc0 = getattr(main_mod, "__doc__", None)
        if not doc0:
            main_mod.__doc__ = "harness\nerror_checker driver\n"   # :69 reads __main__.__doc__
        sys.argv = ["error_checker.py", "-g", d + "/gens.txt", "-p", d + "/ped.txt", "-k", d + "/kids.txt",
                    "-s", d + "/map.csv", "-T", "1", "-c", chrom, "-o", d + "/out"]
        try:
            import warnings
            ctx = contextlib.redirect_stdout(buf) if quiet else contextlib.nullcontext()
            with deterministic_pools(), ctx, contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"), \
                    warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ec.main()
        finally:
            sys.argv = argv0
            main_mod.__doc__ = doc0
        pe = pd.read_csv(d + "/out/probs_errors.csv", index_col=0)
        sums = np.loadtxt(d + "/out/individual_sum_probs_errors.csv", delimiter=",", ndmin=1)
    # to_csv prints the shortest repr of each float16, which identifies the float16 value uniquely
    pe16 = pe.to_numpy().astype(np.float16)
    assert np.allclose(pe16.astype(np.float64), pe.to_numpy(), rtol=2e-3, atol=1e-7, equal_nan=True)
    s16 = sums.astype(np.float16)
    assert np.array_equal(s16.astype(np.float64), sums, equal_nan=True)
    return dict(with_parents=np.array([int(x) for x in pe.index], dtype=np.int64), probs_errors=pe16, sums=s16,
                stdout=buf.getvalue())


def run_preindex(ped_rows, ids, obs, chrom_of_snp, surround=1, init_filter_p=0.9, seedskidschunk=10000, quiet=True):
    """Run the reference's empirical/preindex.py main() (:59-346) on files written from the arrays (same file
    conventions as run_error_checker; the map file has no header and columns chrom,snpid,cM,pos, :141).
    Returns {chromosome: dict(snps, windows, freq=list of float64 arrays in state-product order)} read back from the
    pickled indexes, plus the stdout (quantiles and mask counts are only printed)."""
    load()
    import gzip
    import itertools
    import cloudpickle
    import empirical.preindex as pi
    S = obs.shape[1]
    snps = ["SNP%d" % i for i in range(S)]
    buf = io.StringIO()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        with open(d + "/gens.txt", "w") as f:
            f.write("id " + " ".join(snps) + " pad\n")
            for k, row in zip(ids, obs):
                f.write(str(int(k)) + " " + " ".join(str(int(x)) for x in row) + " 0\n")
        with open(d + "/ped.txt", "w") as f:
            for r in ped_rows:
                f.write(" ".join(str(int(x)) for x in r[:4]) + "\n")
        with open(d + "/map.csv", "w") as f:
            for i, s in enumerate(snps):
                f.write("%s,%s,0,%d\n" % (chrom_of_snp[i], s, 100 * i + 1))
            f.write("0,pad,0,1\n")          # chromosome "0" is skipped (:198)
        argv0 = sys.argv
        main_mod = sys.modules["__main__"]
        doc0 = getattr(main_mod, "__doc__", None)
        if not doc0:
            main_mod.__doc__ = "harness\npreindex driver\n"
        sys.argv = ["preindex.py", "-g", d + "/gens.txt", "-p", d + "/ped.txt", "-s", d + "/map.csv", "-w", str(surround),
                    "-T", "1", "-q", str(init_filter_p), "-c", str(seedskidschunk), "-o", d + "/out"]
        try:
            import warnings
            ctx = contextlib.redirect_stdout(buf) if quiet else contextlib.nullcontext()
            with deterministic_pools(), ctx, contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"), \
                    warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pi.main()
        finally:
            sys.argv = argv0
            main_mod.__doc__ = doc0
        for chrom in sorted(set(chrom_of_snp)):
            path = d + "/out/%s/empiricalIndex.idx.gz" % chrom
            if not os.path.exists(path):
                continue
            with gzip.open(path, "rb") as f:
                e = cloudpickle.load(f)
            names = list(e.snp_ordered)
            freq = []
            for s in names:
                win = e.windows[s]
                tab = e.snp2frequency[s]
                freq.append(np.array([tab[tuple(zip(win, v))] for v in itertools.product([0, 1, 2], repeat=len(win))]
                                     if len(win) else [], dtype=np.float64))
            out[chrom] = dict(snps=names, windows=[list(e.windows[s]) for s in names], freq=freq)
    return out, buf.getvalue()
