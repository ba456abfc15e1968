"""``factor`` — host mirror of reference src/factorization.jl:5-11 over ``hs_factor``."""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from .factornode import FactorNode, _Handle
from .nesteddissection import NDLoc, NestedDissection
from .options import SolverOptions, chkopts, to_c

__all__ = ["factor"]


def factor(A, nd: NestedDissection, nd_loc: NDLoc, opts: SolverOptions = None, device: int = 0, **kw) -> FactorNode:
    """``factor(A::SparseMatrixCSC{T}, nd, nd_loc, opts=SolverOptions(); kw...) → FactorNode{T}``
    (factorization.jl:5-11).  ``A`` must already be permuted like the tree (test/rungmres.jl:17-19).
    Raises what the reference raises: ``ArgumentError`` for bad options, ``DimensionMismatch``, ``SingularException``."""
    opts = (opts or SolverOptions()).copy(**kw)   # :6
    chkopts(opts)                                 # :7
    if not sp.issparse(A):
        raise TypeError("factor expects a sparse matrix (SparseMatrixCSC)")
    if not (sp.issparse(A) and A.format == "csc"):
        A = sp.csc_matrix(A)      # a CSC input is used as it is (its cached format flags stay valid between calls)
    if not A.has_canonical_format:
        # duplicate entries would overwrite each other in the device gather (SparseMatrixCSC cannot hold duplicates)
        A = A.copy()
        A.sum_duplicates()
    if A.shape[0] != A.shape[1]:
        raise _lib.DimensionMismatch(_lib.HS_EDIM, "A must be square")
    cx = np.iscomplexobj(A.data)
    dtype = np.complex128 if cx else np.float64
    n = A.shape[0]
    # SciPy's 0-based (usually int32) index arrays go to the library as they are; it widens / shifts on the device
    flags = _lib.HS_CSC_ZERO_BASED
    if A.indices.dtype == np.int32 and A.indptr.dtype == np.int32:
        colptr, rowval = np.ascontiguousarray(A.indptr), np.ascontiguousarray(A.indices)
        flags |= _lib.HS_CSC_INT32
    else:
        colptr, rowval = _lib.as_i64(A.indptr), _lib.as_i64(A.indices)
    nzval = np.ascontiguousarray(A.data, dtype=dtype)
    nd._need_analyzed()
    from .parallel import _tree_struct
    tree, keep = _tree_struct(nd, nd_loc)
    ctx = _lib.default_context(device)
    copts, _keep_sk = to_c(opts, dtype=dtype)
    h = C.c_void_p()
    rc = _lib.lib.hs_factor(ctx, _lib.HS_C64 if cx else _lib.HS_F64, n, colptr.ctypes.data_as(C.c_void_p),
                            rowval.ctypes.data_as(C.c_void_p), nzval.ctypes.data_as(C.c_void_p), C.byref(tree),
                            C.byref(copts), flags, C.byref(h))
    if rc != _lib.HS_OK:
        if h:
            _lib.lib.hs_factor_free(h)
        _lib.check(rc)
    hd = _Handle(h, ctx, dtype, n, nd, nd_loc)
    # strong references to the arrays that were uploaded: gmres() reuses the device-resident copy only for these very
    # objects (identity, not addresses - a freed temporary's address can be recycled) with unchanged contents
    hd.A_ref = (A.indptr, A.indices, A.data)
    return FactorNode(hd, nd.root)
