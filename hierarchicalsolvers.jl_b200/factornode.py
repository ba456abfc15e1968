"""``FactorNode`` and ``ldiv!`` — host mirror of reference src/factornode.jl over the device-resident factor tree."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib

__all__ = ["FactorNode", "ldiv", "maxrank", "isleaf", "isbranch", "eltype"]


class _Handle:
    """Owns the ``hs_fac``; freed when the last FactorNode view goes away (the Julia shim uses a finalizer)."""

    def __init__(self, h, ctx, dtype, n, nd, nd_loc):
        self.h, self.ctx, self.dtype, self.n, self.nd, self.nd_loc = h, ctx, dtype, n, nd, nd_loc

    def __del__(self):
        try:
            if self.h:
                _lib.lib.hs_factor_free(self.h)
                self.h = None
        except Exception:
            pass


class FactorNode:
    """One eliminated front of the factor tree (factornode.jl:7-39).  ``D, S, L, R`` are copied from the device on
    access and are exactly the reference's quantities: ``D = A_ii``, ``L = A_bi·A_ii⁻¹``, ``R = A_ii⁻¹·A_ib`` and the
    Schur complement ``S`` permuted by ``[int_loc; bnd_loc]``."""

    def __init__(self, handle: _Handle, node: int):
        self._hd, self._k = handle, int(node)

    # ---- fields -------------------------------------------------------------------------------
    def _get(self, which):
        dims = (C.c_int64 * 2)()
        _lib.check(_lib.lib.hs_node_get(self._hd.h, self._k, which, None, dims))
        r, c = int(dims[0]), int(dims[1])
        if which == _lib.HS_GET_PIV:
            out = np.zeros(r, dtype=np.int64)
        else:
            out = np.zeros((r, c), dtype=self._hd.dtype, order="F")
        _lib.check(_lib.lib.hs_node_get(self._hd.h, self._k, which, out.ctypes.data_as(C.c_void_p), dims))
        return out

    D = property(lambda self: self._get(_lib.HS_GET_D))
    S = property(lambda self: self._get(_lib.HS_GET_S))
    L = property(lambda self: self._get(_lib.HS_GET_L))
    R = property(lambda self: self._get(_lib.HS_GET_R))
    front = property(lambda self: self._get(_lib.HS_GET_FRONT))
    piv = property(lambda self: self._get(_lib.HS_GET_PIV))

    @property
    def int(self):
        return self._hd.nd.node(self._k).int

    @property
    def bnd(self):
        return self._hd.nd.node(self._k).bnd

    @property
    def int_loc(self):
        return self._hd.nd_loc.node(self._k).int

    @property
    def bnd_loc(self):
        return self._hd.nd_loc.node(self._k).bnd

    @property
    def left(self) -> Optional["FactorNode"]:
        c = int(self._hd.nd.left[self._k])
        return None if c < 0 else FactorNode(self._hd, c)

    @property
    def right(self) -> Optional["FactorNode"]:
        c = int(self._hd.nd.right[self._k])
        return None if c < 0 else FactorNode(self._hd, c)

    @property
    def dtype(self):
        return self._hd.dtype

    @property
    def shape(self):
        return (self._hd.n, self._hd.n)

    def node(self, k: int) -> "FactorNode":
        """Node ``k`` of the tree in post-order numbering (root = nnodes-1)."""
        return FactorNode(self._hd, k)

    def stats(self) -> dict:
        s = _lib.hs_stats_t()
        _lib.check(_lib.lib.hs_stats(self._hd.h, C.byref(s)))
        return s.asdict()

    def ranks(self):
        """``(rank(F.L), rank(F.R))`` of this node — ``LowRankMatrix`` ranks of a compressed node
        (factorization.jl:173,179), ``(0, 0)`` when ``L`` and ``R`` are dense."""
        rl, rr = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib.hs_node_rank(self._hd.h, self._k, C.byref(rl), C.byref(rr)))
        return int(rl.value), int(rr.value)

    def hss(self):
        """``F.S`` as the reference stores it for a compressed node — an ``HssMatrix`` (factorization.jl:110-111) — or
        ``None`` when ``S`` is dense.  Returns the HSS tree in pre-order as a list of dicts: ``lo, hi`` (rows of
        ``S[perm,perm]``), ``left, right, parent`` (list positions, -1 = none), ``leaf`` and the generators of
        HssMatrices.jl's type: leaves ``D, U, V``; branches ``B12, B21`` and (below the root) ``R1, R2, W1, W2``."""
        n = C.c_int64()
        _lib.check(_lib.lib.hs_hss_info(self._hd.h, self._k, C.byref(n), None))
        if n.value == 0:
            return None
        info = np.zeros((n.value, 8), dtype=np.int64)
        _lib.check(_lib.lib.hs_hss_info(self._hd.h, self._k, C.byref(n), info.ctypes.data_as(_lib.i64p)))

        def get(t, which):
            dims = (C.c_int64 * 2)()
            _lib.check(_lib.lib.hs_hss_get(self._hd.h, self._k, t, which, None, dims))
            out = np.zeros((int(dims[0]), int(dims[1])), dtype=self._hd.dtype, order="F")
            if out.size:
                _lib.check(_lib.lib.hs_hss_get(self._hd.h, self._k, t, which, out.ctypes.data_as(C.c_void_p), dims))
            return out
        nodes = []
        for t in range(n.value):
            lo, hi, left, right, r0, r1, parent, leaf = (int(v) for v in info[t])
            d = dict(lo=lo, hi=hi, left=left, right=right, parent=parent, leaf=bool(leaf), rank_u=r0, rank_v=r1)
            if leaf:
                d.update(D=get(t, _lib.HS_HSS_D), U=get(t, _lib.HS_HSS_U), V=get(t, _lib.HS_HSS_V))
            else:
                d.update(B12=get(t, _lib.HS_HSS_B12), B21=get(t, _lib.HS_HSS_B21))
                if parent >= 0:
                    R, W = get(t, _lib.HS_HSS_R), get(t, _lib.HS_HSS_W)
                    ra0, ra1 = int(info[left][4]), int(info[left][5])
                    d.update(R1=R[:ra0], R2=R[ra0:], W1=W[:ra1], W2=W[ra1:])
            nodes.append(d)
        return nodes

    def hssrank(self) -> int:
        """``hssrank(F.S)`` (factornode.jl:53): largest off-diagonal generator rank; 0 when ``S`` is dense."""
        nodes = self.hss()
        if not nodes:
            return 0
        return max([max(*d["B12"].shape, *d["B21"].shape) for d in nodes if not d["leaf"]] + [0])

    def resolved_swlevel(self) -> int:
        v = C.c_int64()
        _lib.check(_lib.lib.hs_resolved_swlevel(self._hd.h, C.byref(v)))
        return int(v.value)

    def refactor(self, A) -> "FactorNode":
        """Numeric re-factorization with new values on the same sparsity and tree."""
        import scipy.sparse as sp
        if not (sp.issparse(A) and A.format == "csc"):
            A = sp.csc_matrix(A)
        if not A.has_canonical_format:
            A = A.copy()
            A.sum_duplicates()
        if np.iscomplexobj(A.data) and self._hd.dtype == np.float64:
            raise TypeError("refactor: complex values for a real factorization")
        nz = np.ascontiguousarray(A.data, dtype=self._hd.dtype)
        _lib.check(_lib.lib.hs_refactor(self._hd.h, nz.ctypes.data_as(C.c_void_p), 0))
        self._hd.A_ref = (A.indptr, A.indices, A.data)     # what the device now holds
        return self

    def __repr__(self):
        return f"FactorNode{{{np.dtype(self._hd.dtype).name}}}"  # Base.show factornode.jl:42

    # `F \ B`
    def solve(self, B):
        return ldiv(self, B)


def eltype(F: FactorNode):
    return F.dtype


def isleaf(F) -> bool:
    return F.left is None and F.right is None  # factornode.jl:45


def isbranch(F) -> bool:
    return F.left is not None and F.right is not None  # factornode.jl:46


def maxrank(F: FactorNode) -> int:
    """factornode.jl:49-57."""
    v = C.c_int64()
    _lib.check(_lib.lib.hs_maxrank(F._hd.h, C.byref(v)))
    return int(v.value)


def ldiv(*args):
    """``ldiv!`` (factornode.jl:62-74).

    ``ldiv(F, B)``     → new array, ``B`` untouched (the reference's 2-argument method is NOT in place, :62)
    ``ldiv(C, F, B)``  → writes the result into ``C`` and returns it (vectors :63-67, matrices :68-74)
    """
    if len(args) == 2:
        F, B = args
        Cout = None
    elif len(args) == 3:
        Cout, F, B = args
    else:
        raise TypeError("ldiv(F, B) or ldiv(C, F, B)")
    if F._k != F._hd.nd.root:
        raise ValueError("ldiv! is defined on the root FactorNode")
    hd = F._hd
    B = np.asarray(B)
    if B.shape[0] != hd.n:
        raise _lib.DimensionMismatch(_lib.HS_EDIM, f"B has {B.shape[0]} rows, expected {hd.n}")
    if np.iscomplexobj(B) and hd.dtype == np.float64:
        # the reference has no ldiv! method for mismatched element types (MethodError); never drop the imaginary part
        raise TypeError("ldiv: complex right-hand side for a real (Float64) factorization")
    vec = B.ndim == 1
    nrhs = 1 if vec else B.shape[1]
    Bf = np.asfortranarray(B.reshape(hd.n, nrhs), dtype=hd.dtype)
    if Bf is B or np.shares_memory(Bf, B):
        Bf = Bf.copy(order="F")
    _lib.check(_lib.lib.hs_solve(hd.h, nrhs, Bf.ctypes.data_as(C.c_void_p), hd.n, Bf.ctypes.data_as(C.c_void_p), hd.n, 0))
    res = Bf.reshape(-1) if vec else Bf
    if Cout is None:
        return res
    Cout[...] = res
    return Cout
