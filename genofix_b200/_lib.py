"""ctypes binding of libgenofix_b200.so (the C ABI declared in include/genofix_b200.h).

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GFX_LIB_PATH") or os.path.join(_HERE, "libgenofix_b200.so")   # override: kernel experiments only

GFX_OK = 0
GFX_E_INVALID = -1
GFX_E_CUDA = -2
GFX_E_PEDIGREE = -3
GFX_E_TOO_WIDE = -4
GFX_E_NOMEM = -5

EXPORTS = [
    "gfx_last_error", "gfx_version", "gfx_launch_count", "gfx_schedule_create", "gfx_schedule_create_host",
    "gfx_schedule_destroy", "gfx_schedule_info", "gfx_schedule_copy", "gfx_pack2", "gfx_unpack2",
    "gfx_mendel_rank", "gfx_mendel_rank_hist", "gfx_rank_hist", "gfx_rank_rowsum", "gfx_rank_rowsum_f16", "gfx_correct_pairs", "gfx_correct_singles", "gfx_build_cut_table",
    "gfx_check_device_status", "gfx_reset_device_status", "gfx_debug_counters", "gfx_rank_form", "gfx_rank_timing", "gfx_rank_last_kernel_ms", "gfx_trio_scan", "gfx_trio_scan_add", "gfx_founders", "gfx_window_counts", "gfx_infer_host", "gfx_infer_pair_host",
    "gfx_kids_term_host", "gfx_emp_rank_blend", "gfx_emp_apply", "gfx_gene_drop", "gfx_gene_drop_linked", "gfx_haps_to_codes", "gfx_inject_errors", "gfx_confusion",
]


class PedigreeDesc(C.Structure):
    _fields_ = [("n_ped", C.c_int32), ("n_rows", C.c_int32), ("back", C.c_int32),
                ("ped_id", C.c_void_p), ("sire", C.c_void_p), ("dam", C.c_void_p),
                ("generation", C.c_void_p), ("row", C.c_void_p), ("kid_off", C.c_void_p), ("kid", C.c_void_p)]


class ScheduleInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_rows", "n_generations", "n_pairs", "n_pair_jobs", "n_singles", "n_loopy", "n_blankets",
        "max_blanket_nodes", "max_width", "max_pairs_in_generation", "n_trios", "n_lone")]


class CorrectParams(C.Structure):
    _fields_ = [("test_p", C.c_double), ("suspect_t", C.c_double), ("tiethreshold", C.c_double),
                ("threshold", C.c_double), ("err_thresh", C.c_double), ("cutraw_dev", C.c_void_p),
                ("tp_raw", C.c_uint32), ("reserved", C.c_uint32)]


class EmpIndex(C.Structure):
    _fields_ = [("wstart_dev", C.c_void_p), ("wlen_dev", C.c_void_p), ("foff_dev", C.c_void_p), ("freq_dev", C.c_void_p),
                ("weight", C.c_double), ("max_window", C.c_int32), ("reserved", C.c_int32)]


class GfxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("genofix_b200 [%d]: %s" % (code, msg))
        self.code = code
        self.msg = msg


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built: run __graft_entry__.build()."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libgenofix_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`"
                          % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.gfx_last_error.restype = C.c_char_p
    L.gfx_launch_count.restype = C.c_int64
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.gfx_schedule_create.argtypes = [C.POINTER(PedigreeDesc), C.POINTER(vp)]
    L.gfx_schedule_create_host.argtypes = [C.POINTER(PedigreeDesc), C.POINTER(vp)]
    L.gfx_schedule_destroy.argtypes = [vp]
    L.gfx_schedule_destroy.restype = None
    L.gfx_schedule_info.argtypes = [vp, C.POINTER(ScheduleInfo)]
    L.gfx_schedule_copy.argtypes = [vp, C.c_int, vp, i64]
    L.gfx_pack2.argtypes = [vp, i64, i64, i64, vp, i64, vp]
    L.gfx_unpack2.argtypes = [vp, i64, i64, i64, vp, i64, vp]
    L.gfx_mendel_rank.argtypes = [vp, vp, i64, i64, vp, i64, vp]
    L.gfx_mendel_rank_hist.argtypes = [vp, vp, i64, i64, vp, i64, vp, vp]
    L.gfx_rank_hist.argtypes = [vp, i64, vp, i64, i64, i64, vp, vp]
    L.gfx_rank_rowsum.argtypes = [vp, i64, vp, i64, i64, i64, vp, C.c_double, vp, vp]
    L.gfx_rank_rowsum_f16.argtypes = [vp, i64, i64, i64, vp, vp, i32, vp, vp]
    L.gfx_correct_pairs.argtypes = [vp, i32, vp, i64, i64, vp, i64, vp, C.POINTER(CorrectParams), vp, i64, vp, vp, vp]
    L.gfx_correct_singles.argtypes = [vp, vp, i64, i64, vp, i64, vp, C.POINTER(CorrectParams), vp, i64, vp, vp, vp]
    L.gfx_gene_drop.argtypes = [i32, vp, vp, vp, vp, i32, i64, i64, C.c_uint64, vp, vp]
    L.gfx_gene_drop_linked.argtypes = [i32, vp, vp, vp, vp, i32, i64, i64, C.c_uint64, i32, vp, vp, vp, vp, vp]
    L.gfx_haps_to_codes.argtypes = [vp, i64, i64, i64, vp, vp]
    L.gfx_inject_errors.argtypes = [vp, i64, i64, i64, C.c_double, C.c_uint64, vp]
    L.gfx_confusion.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp]
    L.gfx_build_cut_table.argtypes = [vp, C.c_double, C.c_double, vp, C.POINTER(C.c_uint32)]
    L.gfx_check_device_status.argtypes = [vp, vp]
    L.gfx_reset_device_status.argtypes = [vp]
    L.gfx_debug_counters.argtypes = [vp, C.c_int, vp]
    L.gfx_emp_rank_blend.argtypes = [vp, i64, i64, i64, vp, i64, vp, C.POINTER(EmpIndex), vp, i64, vp]
    L.gfx_emp_apply.argtypes = [vp, i32, vp, vp, i64, i64, vp, i64, vp, C.POINTER(EmpIndex), C.POINTER(CorrectParams), vp, C.c_double, vp, vp]
    L.gfx_rank_form.argtypes = [vp, C.c_int]
    L.gfx_rank_timing.argtypes = [vp, C.c_int]
    L.gfx_rank_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.gfx_trio_scan.argtypes = [vp, vp, i64, i64, vp, vp, vp]
    L.gfx_trio_scan_add.argtypes = [vp, vp, i64, i64, vp, vp, vp]
    L.gfx_founders.argtypes = [C.POINTER(PedigreeDesc), vp, vp, C.POINTER(i32)]
    L.gfx_window_counts.argtypes = [vp, vp, i64, i64, i64, i64, vp, vp, i32, i32, vp, vp]
    L.gfx_infer_host.argtypes = [i32, vp, vp, vp, i32, vp]
    L.gfx_infer_pair_host.argtypes = [i32, vp, vp, vp, i32, i32, vp]
    L.gfx_kids_term_host.argtypes = [i32, vp, vp, vp]
    _lib = L
    return L


def check(rc):
    """Raise on a negative status; KeyError for pedigrees outside the reference's domain (the
    reference raises KeyError there, pedigree/pedigree_dag.py:126-128)."""
    if rc >= 0:
        return rc
    msg = lib().gfx_last_error().decode("utf-8", "replace")
    if rc == GFX_E_PEDIGREE:
        raise KeyError(msg)
    raise GfxError(rc, msg)
