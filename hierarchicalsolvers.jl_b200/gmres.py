"""GMRES with the factorization as right preconditioner — the reference driver's solve step
(test/rungmres.jl:47-48: ``gmres(A, b; Pr=F, reltol=1e-9, restart=30, log=true, maxiter=30)``)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import scipy.sparse as sp

from . import _lib
from .factornode import FactorNode, ldiv

__all__ = ["gmres", "ConvergenceHistory"]


@dataclass
class ConvergenceHistory:
    """Subset of IterativeSolvers' ``ConvergenceHistory`` the driver reads: ``ch[:resnorm]``, iterations, convergence."""
    resnorm: List[float] = field(default_factory=list)
    isconverged: bool = False

    @property
    def iters(self) -> int:
        return len(self.resnorm)

    def __getitem__(self, key):
        if key in ("resnorm", ":resnorm"):
            return self.resnorm
        raise KeyError(key)


def gmres(A, b, Pr: Optional[FactorNode] = None, reltol: float = 1e-9, restart: int = 30, maxiter: int = 30,
          log: bool = False, device_resident: bool = True, device: int = 0, A_is_factored: Optional[bool] = None):
    """Restarted GMRES(restart), x0 = 0, right preconditioner ``Pr`` (a ``FactorNode``), modified Gram-Schmidt.

    ``device_resident=True`` runs the whole iteration in HBM (``hs_gmres``); ``False`` drives the Arnoldi loop on
    the host and calls ``ldiv!`` once per step the way IterativeSolvers does.

    ``A_is_factored``: ``True`` - ``A`` is the matrix ``Pr`` factored (the copy already resident in HBM is used, nothing
    is uploaded); ``False`` - upload ``A``; ``None`` (default) - reuse the resident copy only if ``A``'s arrays are the
    very objects handed to ``factor``/``refactor`` AND their contents still match the device copy (bitwise checksum),
    otherwise upload."""
    if not (sp.issparse(A) and A.format == "csc"):
        A = sp.csc_matrix(A)
    n = A.shape[0]
    cx = np.iscomplexobj(A.data) or np.iscomplexobj(b) or (Pr is not None and Pr.dtype == np.complex128)
    dtype = np.complex128 if cx else np.float64
    b = np.ascontiguousarray(b, dtype=dtype)
    if device_resident:
        x = np.zeros(n, dtype=dtype)
        res = np.zeros(max(maxiter, 1), dtype=np.float64)
        nit, conv = C.c_int64(), C.c_int32()
        ctx = Pr._hd.ctx if Pr is not None else _lib.default_context(device)
        same = bool(A_is_factored) and Pr is not None
        if A_is_factored is None and Pr is not None:
            ref = getattr(Pr._hd, "A_ref", None)
            if ref is not None and A.indptr is ref[0] and A.indices is ref[1] and A.data is ref[2] and A.data.dtype == dtype:
                w = np.ascontiguousarray(A.data).view(np.uint64)
                hsum, hxor = C.c_uint64(), C.c_uint64()
                _lib.check(_lib.lib.hs_matrix_checksum(Pr._hd.h, C.byref(hsum), C.byref(hxor)))
                same = (int(np.add.reduce(w, dtype=np.uint64)) == hsum.value
                        and int(np.bitwise_xor.reduce(w)) == hxor.value)
        if same:   # the matrix the factorization already holds in HBM
            cp = rv = nzp = None
        else:
            colptr, rowval = _lib.as_i64(A.indptr), _lib.as_i64(A.indices)
            nz = np.ascontiguousarray(A.data, dtype=dtype)
            cp, rv, nzp = (a.ctypes.data_as(C.c_void_p) for a in (colptr, rowval, nz))
        _lib.check(_lib.lib.hs_gmres(ctx, _lib.HS_C64 if cx else _lib.HS_F64, n, cp, rv, nzp, 0,
                                     Pr._hd.h if Pr is not None else None, b.ctypes.data_as(C.c_void_p),
                                     x.ctypes.data_as(C.c_void_p), reltol, restart, maxiter,
                                     res.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nit), C.byref(conv), 0))
        hist = ConvergenceHistory(list(res[: nit.value]), bool(conv.value))
        return (x, hist) if log else x
    # host-driven loop: one ldiv! per Arnoldi step
    P = (lambda v: v) if Pr is None else (lambda v: ldiv(Pr, v))
    x = np.zeros(n, dtype=dtype)
    r = b.copy()
    beta = np.linalg.norm(r)
    tol = reltol * beta
    hist = ConvergenceHistory()
    it, resid = 0, beta
    while it < maxiter and resid > tol:
        V = np.zeros((n, restart + 1), dtype=dtype)
        H = np.zeros((restart + 1, restart), dtype=dtype)
        cs = np.zeros(restart, dtype=dtype)
        sn = np.zeros(restart, dtype=dtype)
        g = np.zeros(restart + 1, dtype=dtype)
        V[:, 0] = r / beta
        g[0] = beta
        k = 0
        while k < restart and it < maxiter and resid > tol:
            w = A @ P(V[:, k])
            for j in range(k + 1):
                H[j, k] = np.vdot(V[:, j], w)
                w = w - H[j, k] * V[:, j]
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] != 0:
                V[:, k + 1] = w / H[k + 1, k]
            for j in range(k):
                t = cs[j] * H[j, k] + sn[j] * H[j + 1, k]
                H[j + 1, k] = -np.conj(sn[j]) * H[j, k] + cs[j] * H[j + 1, k]
                H[j, k] = t
            a, c = H[k, k], H[k + 1, k]
            den = np.sqrt(abs(a) ** 2 + abs(c) ** 2)
            if den == 0:
                cs[k], sn[k] = 1.0, 0.0
            elif a == 0:
                cs[k], sn[k] = 0.0, 1.0
            else:
                cs[k], sn[k] = abs(a) / den, (a / abs(a)) * np.conj(c) / den
            H[k, k] = cs[k] * a + sn[k] * c
            H[k + 1, k] = 0.0
            g[k + 1] = -np.conj(sn[k]) * g[k]
            g[k] = cs[k] * g[k]
            resid = abs(g[k + 1])
            hist.resnorm.append(float(resid))
            k += 1
            it += 1
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k]) if k else np.zeros(0, dtype=dtype)
        x = x + P(V[:, :k] @ y)
        if it < maxiter and resid > tol:
            r = b - A @ x
            beta = np.linalg.norm(r)
            resid = beta
    hist.isconverged = bool(resid <= tol)
    return (x, hist) if log else x
