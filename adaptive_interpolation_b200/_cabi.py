"""ctypes binding of libai_b200.so (include/ai_b200.h).

This is the same mechanism the reference uses to call its generated evaluator
(run_test.py:33-43 `cptr_from_numpy`, :250-251 `ctypes.cdll.LoadLibrary(...).eval`), pointed at
ONE prebuilt CUDA library instead of a per-approximation ISPC build.

There is no CPU fallback: if the library is missing or cannot be loaded, `lib()` raises.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# AI_B200_LIB selects another build of the same library (tools/tune.py variants); never a fallback
LIB_PATH = os.environ.get("AI_B200_LIB") or os.path.join(_PKG, "_lib", "libai_b200.so")

AI_OK = 0
BASIS_IDS = {"chebyshev": 0, "legendre": 1, "monomial": 2}
DTYPE_F32, DTYPE_F64 = 0, 1
MAX_ORDER = 16
MAX_MULTI = 4
TREE_LAYOUT_STANDARD, TREE_LAYOUT_TRIM = 0, 1

# every symbol include/ai_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "ai_last_error", "ai_version", "ai_device_count", "ai_device_sm_arch",
    "ai_table_create", "ai_table_create_from_tree", "ai_table_plan", "ai_table_destroy", "ai_table_get_info",
    "ai_table_get_flat",
    "ai_eval_f32", "ai_eval_f64", "ai_eval_index_f32", "ai_eval_index_f64",
    "ai_eval_sum_f32", "ai_eval_sum_f64", "ai_eval_multi_f32", "ai_eval_multi_f64",
    "ai_eval_host_f32", "ai_eval_host_f64",
    "ai_host_alloc", "ai_host_free", "ai_host_copy", "ai_measure_fma_peak", "ai_copy_f32", "ai_measure_host_roundtrip",
]


class TableInfo(ctypes.Structure):
    _fields_ = [
        ("basis", ctypes.c_int32), ("order", ctypes.c_int32), ("dtype", ctypes.c_int32),
        ("device", ctypes.c_int32), ("n_leaves", ctypes.c_int64), ("grid_cells", ctypes.c_int32),
        ("grid_log2_scale", ctypes.c_int32), ("search_steps", ctypes.c_int32),
        ("record_stride", ctypes.c_int32), ("image_bytes", ctypes.c_int64),
        ("smem_bytes", ctypes.c_int64), ("residency", ctypes.c_int32), ("bank_copies", ctypes.c_int32),
        ("smem_resident", ctypes.c_int32), ("lookup_mode", ctypes.c_int32), ("block_threads", ctypes.c_int32),
        ("blocks_per_sm", ctypes.c_int32), ("sm_count", ctypes.c_int32),
        ("lower_bound", ctypes.c_double), ("upper_bound", ctypes.c_double),
        ("header_packed", ctypes.c_int32), ("cell_records", ctypes.c_int32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class AiError(Exception):
    """A failing C-ABI call; carries the status code and ai_last_error()."""

    def __init__(self, code, message):
        Exception.__init__(self, "ai_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib():
    """Load the CUDA library (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libai_b200.so not found at %s. Build it with `python -m adaptive_interpolation_b200.build` "
            "(needs nvcc); there is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.ai_last_error.restype = ctypes.c_char_p
    L.ai_version.restype = ctypes.c_char_p
    L.ai_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.ai_device_sm_arch.argtypes = [i32] + [ctypes.POINTER(ctypes.c_int)] * 3
    L.ai_table_create.argtypes = [i32, i32, i32, i64, vp, vp, i32, ctypes.POINTER(vp)]
    L.ai_table_create_from_tree.argtypes = [i32, i32, i32, i32, i32, vp, i64, i32, ctypes.POINTER(vp)]
    L.ai_table_plan.argtypes = [i32, i32, i32, i64, vp, vp, i64, ctypes.POINTER(TableInfo)]
    L.ai_table_destroy.argtypes = [vp]
    L.ai_table_get_info.argtypes = [vp, ctypes.POINTER(TableInfo)]
    L.ai_table_get_flat.argtypes = [vp, vp, vp]
    for name in ("ai_eval_f32", "ai_eval_f64", "ai_eval_index_f32", "ai_eval_index_f64",
                 "ai_eval_sum_f32", "ai_eval_sum_f64"):
        getattr(L, name).argtypes = [vp, vp, vp, i64, vp]
    for name in ("ai_eval_multi_f32", "ai_eval_multi_f64"):
        getattr(L, name).argtypes = [ctypes.POINTER(vp), i32, vp, ctypes.POINTER(vp), i64, vp,
                                     ctypes.POINTER(ctypes.c_int)]
    for name in ("ai_eval_host_f32", "ai_eval_host_f64"):
        getattr(L, name).argtypes = [vp, vp, vp, i64, ctypes.POINTER(ctypes.c_double)]
    L.ai_host_alloc.argtypes = [ctypes.POINTER(vp), i64]
    L.ai_host_free.argtypes = [vp]
    L.ai_host_copy.argtypes = [vp, vp, i64, i32]
    L.ai_measure_fma_peak.argtypes = [i32, i32, ctypes.POINTER(ctypes.c_double)]
    L.ai_copy_f32.argtypes = [vp, vp, i64, i32, vp]
    L.ai_measure_host_roundtrip.argtypes = [vp, vp, vp, i64]
    for name in EXPORTS:
        if name not in ("ai_last_error", "ai_version"):
            getattr(L, name).restype = ctypes.c_int
    _lib = L
    return L


def check(code):
    if code != AI_OK:
        raise AiError(code, lib().ai_last_error().decode("utf-8", "replace"))


def version():
    return lib().ai_version().decode()


def device_count():
    n = ctypes.c_int(0)
    check(lib().ai_device_count(ctypes.byref(n)))
    return n.value
