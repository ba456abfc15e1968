"""ctypes binding of libhsolve_cuda (include/hsolve_cuda.h).  There is no CPU fallback: if the shared library
is missing this module raises on import, and every compute entry point fails without a B200."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# LIBHSOLVE_CUDA selects another build of the same library (kernel-variant experiments); default: the in-tree build
LIB_PATH = os.environ.get("LIBHSOLVE_CUDA") or os.path.join(_HERE, "libhsolve_cuda.so")

HS_OK, HS_EARG, HS_EDIM, HS_ETREE, HS_ESINGULAR, HS_ECUDA, HS_ENOMEM, HS_ENOTIMPL, HS_ESIZE = range(9)
HS_F64, HS_C64 = 0, 1
HS_ON_DEVICE, HS_CSC_ZERO_BASED, HS_CSC_INT32 = 1, 2, 4
HS_GET_D, HS_GET_S, HS_GET_L, HS_GET_R, HS_GET_FRONT, HS_GET_PIV = range(6)
HS_HSS_D, HS_HSS_U, HS_HSS_V, HS_HSS_B12, HS_HSS_B21, HS_HSS_R, HS_HSS_W = range(7)

i64p = C.POINTER(C.c_int64)


class hs_opts(C.Structure):
    _fields_ = [("swlevel", C.c_int64), ("swsize", C.c_int64), ("atol", C.c_double), ("rtol", C.c_double),
                ("c_tol", C.c_double), ("leafsize", C.c_int64), ("kest", C.c_int64), ("stepsize", C.c_int64),
                ("verbose", C.c_int32), ("subtree", C.c_int32),
                # extension: HSS storage of the Schur complements + the host-supplied sketch matrices (hsolve_cuda.h)
                ("hss", C.c_int32), ("pad0", C.c_int32), ("sketch_omega", C.c_void_p), ("sketch_psi", C.c_void_p),
                ("sketch_rows", C.c_int64), ("sketch_cols", C.c_int64), ("sketch_seed", C.c_uint64)]


class hs_elimtree(C.Structure):
    _fields_ = [("nnodes", C.c_int64), ("fathers", i64p), ("lsons", i64p), ("rsons", i64p), ("inter_ptr", i64p),
                ("inter_idx", i64p), ("bound_ptr", i64p), ("bound_idx", i64p), ("index_base", C.c_int32)]


class hs_tree(C.Structure):
    _fields_ = [("nnodes", C.c_int64), ("left", i64p), ("right", i64p), ("int_ptr", i64p), ("int_idx", i64p),
                ("bnd_ptr", i64p), ("bnd_idx", i64p), ("iloc_ptr", i64p), ("iloc_idx", i64p), ("bloc_ptr", i64p),
                ("bloc_idx", i64p), ("index_base", C.c_int32)]


class hs_stats_t(C.Structure):
    _fields_ = [("nnodes", C.c_int64), ("nlevels", C.c_int64), ("n", C.c_int64), ("max_ni", C.c_int64),
                ("max_nb", C.c_int64), ("factor_flops", C.c_double), ("solve_bytes", C.c_double),
                ("extadd_bytes", C.c_double), ("front_bytes", C.c_double), ("ms_analyze", C.c_double),
                ("ms_h2d", C.c_double), ("ms_assemble", C.c_double), ("ms_panel", C.c_double), ("ms_trsm", C.c_double),
                ("ms_gemm", C.c_double), ("ms_factor_total", C.c_double), ("ms_solve_fwd", C.c_double),
                ("ms_solve_bwd", C.c_double), ("ms_solve_total", C.c_double), ("launches_factor", C.c_int64),
                ("launches_solve", C.c_int64), ("singular_front", C.c_int64), ("singular_col", C.c_int64),
                ("maxrank", C.c_int64), ("gemm_flops", C.c_double), ("gemm_launches", C.c_int64),
                ("panel_launches", C.c_int64), ("ms_extend_add", C.c_double), ("ms_small", C.c_double), ("ms_solve_prep", C.c_double),
                ("ms_compress", C.c_double), ("lowrank_bytes", C.c_double),
                ("ms_hss", C.c_double), ("hss_bytes", C.c_double), ("hss_maxrank", C.c_int64), ("hss_rounds", C.c_int64),
                ("hss_nodes", C.c_int64), ("sketch_flops", C.c_double), ("gemm_flops_big", C.c_double), ("ms_gemm_big", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/hsolve_cuda.h declares: (name, restype, argtypes)
PROTOTYPES = [
    ("hs_version", C.c_int32, []),
    ("hs_last_error", C.c_char_p, []),
    ("hs_create", C.c_int32, [C.POINTER(C.c_void_p), C.c_int32]),
    ("hs_set_stream", C.c_int32, [C.c_void_p, C.c_void_p]),
    ("hs_destroy", C.c_int32, [C.c_void_p]),
    ("hs_device_count", C.c_int32, []),
    ("hs_set_profile", C.c_int32, [C.c_void_p, C.c_int32]),
    ("hs_launch_count", C.c_int32, [C.c_void_p, i64p]),
    ("hs_nd_create", C.c_int32, [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_void_p)]),
    ("hs_nd_elimtree", C.c_int32, [C.c_void_p, C.POINTER(hs_elimtree)]),
    ("hs_nd_free", C.c_int32, [C.c_void_p]),
    ("hs_symfact", C.c_int32, [C.POINTER(hs_elimtree), C.c_int32, C.POINTER(C.c_void_p)]),
    ("hs_symbolic_tree", C.c_int32, [C.c_void_p, C.POINTER(hs_tree)]),
    ("hs_symbolic_perm", C.c_int32, [C.c_void_p, C.POINTER(i64p), i64p]),
    ("hs_symbolic_depth", C.c_int32, [C.c_void_p, i64p]),
    ("hs_symbolic_free", C.c_int32, [C.c_void_p]),
    ("hs_factor", C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(hs_tree),
                              C.POINTER(hs_opts), C.c_int32, C.POINTER(C.c_void_p)]),
    ("hs_refactor", C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32]),
    ("hs_analyze", C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(hs_tree),
                               C.POINTER(hs_opts), C.c_int32, C.POINTER(C.c_void_p)]),
    ("hs_schur_export", C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    ("hs_schur_import", C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    ("hs_solve_sweep", C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32]),
    ("hs_factor_free", C.c_int32, [C.c_void_p]),
    ("hs_solve", C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32]),
    ("hs_node_get", C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, i64p]),
    ("hs_maxrank", C.c_int32, [C.c_void_p, i64p]),
    ("hs_node_rank", C.c_int32, [C.c_void_p, C.c_int64, i64p, i64p]),
    ("hs_stats", C.c_int32, [C.c_void_p, C.POINTER(hs_stats_t)]),
    ("hs_resolved_swlevel", C.c_int32, [C.c_void_p, i64p]),
    ("hs_spmv", C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("hs_matrix_device", C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), i64p]),
    ("hs_matrix_checksum", C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("hs_hss_info", C.c_int32, [C.c_void_p, C.c_int64, i64p, i64p]),
    ("hs_hss_get", C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, i64p]),
    ("hs_gmres", C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_int64, C.POINTER(C.c_double), i64p,
                             C.POINTER(C.c_int32), C.c_int32]),
]

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a).  This package has no CPU fallback.")

lib = C.CDLL(LIB_PATH)
for _name, _res, _args in PROTOTYPES:
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class HSolveError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class SingularException(HSolveError, np.linalg.LinAlgError):
    pass


class DimensionMismatch(HSolveError, ValueError):
    pass


class ArgumentError(HSolveError, ValueError):
    pass


def check(code: int) -> None:
    """Map a status code to the exception the reference would raise."""
    if code == HS_OK:
        return
    msg = (lib.hs_last_error() or b"").decode("utf-8", "replace")
    if code == HS_ESINGULAR:
        raise SingularException(code, msg)
    if code == HS_EDIM:
        raise DimensionMismatch(code, msg)
    if code == HS_EARG:
        raise ArgumentError(code, msg)
    if code == HS_ENOMEM:
        raise MemoryError(msg)
    if code == HS_ENOTIMPL:
        raise NotImplementedError(msg)
    raise HSolveError(code, msg)


def as_i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def ptr(a: np.ndarray, typ=i64p):
    return a.ctypes.data_as(typ)


_ctx_cache = {}


def default_context(device: int = 0) -> C.c_void_p:
    """One hs_ctx per device per process."""
    if device not in _ctx_cache:
        h = C.c_void_p()
        check(lib.hs_create(C.byref(h), device))
        _ctx_cache[device] = h
    return _ctx_cache[device]
