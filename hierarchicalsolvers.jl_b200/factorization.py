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
    A = sp.csc_matrix(A)
    if A.shape[0] != A.shape[1]:
        raise _lib.DimensionMismatch(_lib.HS_EDIM, "A must be square")
    A.sort_indices()
    A.sum_duplicates()
    cx = np.iscomplexobj(A.data)
    dtype = np.complex128 if cx else np.float64
    n = A.shape[0]
    colptr = _lib.as_i64(A.indptr)
    rowval = _lib.as_i64(A.indices)
    nzval = np.ascontiguousarray(A.data, dtype=dtype)
    nd._need_analyzed()
    arrs = dict(left=nd.left, right=nd.right, int_ptr=nd.int_ptr, int_idx=nd.int_idx, bnd_ptr=nd.bnd_ptr,
                bnd_idx=nd.bnd_idx, iloc_ptr=nd_loc.iloc_ptr, iloc_idx=nd_loc.iloc_idx, bloc_ptr=nd_loc.bloc_ptr,
                bloc_idx=nd_loc.bloc_idx)
    arrs = {k: _lib.as_i64(v) for k, v in arrs.items()}
    # tree ids are 1-based; colptr/rowval come from SciPy 0-based → shift them to the tree's base
    tree = _lib.hs_tree(nd.nnodes, *[_lib.ptr(arrs[k]) for k in ("left", "right", "int_ptr", "int_idx", "bnd_ptr", "bnd_idx",
                                                                 "iloc_ptr", "iloc_idx", "bloc_ptr", "bloc_idx")], 1)
    # left/right are node numbers (0-based) in the flat tree: lift them to the base as well
    l1 = np.where(arrs["left"] >= 0, arrs["left"] + 1, -1).astype(np.int64)
    r1 = np.where(arrs["right"] >= 0, arrs["right"] + 1, -1).astype(np.int64)
    tree.left, tree.right = _lib.ptr(l1), _lib.ptr(r1)
    colptr1, rowval1 = colptr + 1, rowval + 1
    ctx = _lib.default_context(device)
    copts = to_c(opts)
    h = C.c_void_p()
    rc = _lib.lib.hs_factor(ctx, _lib.HS_C64 if cx else _lib.HS_F64, n, colptr1.ctypes.data_as(C.c_void_p),
                            rowval1.ctypes.data_as(C.c_void_p), nzval.ctypes.data_as(C.c_void_p), C.byref(tree),
                            C.byref(copts), 0, C.byref(h))
    if rc != _lib.HS_OK:
        if h:
            _lib.lib.hs_factor_free(h)
        _lib.check(rc)
    return FactorNode(_Handle(h, ctx, dtype, n, nd, nd_loc), nd.root)
