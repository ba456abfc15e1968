"""Problem sources: the reference's ``.mat`` problem format and synthetic grid problems.

The reference ships no ordering code: every problem arrives as a MATLAB file holding a sparse
``A``, a right-hand side ``b`` and a serialized elimination tree ``elim_tree``
(reference ``util/read_problem.jl:5-25``; schema in ``src/nesteddissection.jl:105-148``).
The four fixtures the reference's driver uses are not distributed with it, so this module also
generates stand-ins with the same schema: 5-point (2D) / 7-point (3D) Poisson and complex
Helmholtz operators on structured grids with a geometric-bisection elimination tree.

Everything here is host-side integer/sparse bookkeeping.  It is shared by the product, the tests and
the benchmark; it contains no solver arithmetic.

Elimination-tree container (``ElimTree``) — ragged version of the ``.mat`` schema
    fathers, lsons, rsons : int64[nnodes]   1-based node ids, -1 = none   (nesteddissection.jl:110,122)
    inter_ptr/inter_idx    : CSR-style ragged lists of 1-based DOF ids    (``inter[1:ninter[i], i]``)
    bound_ptr/bound_idx    : likewise for ``bound``
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

__all__ = [
    "ElimTree",
    "Problem",
    "grid_problem",
    "grid_elimtree",
    "grid_operator",
    "read_problem",
    "write_problem",
    "nested_dissection",
]


@dataclass
class ElimTree:
    fathers: np.ndarray
    lsons: np.ndarray
    rsons: np.ndarray
    inter_ptr: np.ndarray
    inter_idx: np.ndarray
    bound_ptr: np.ndarray
    bound_idx: np.ndarray

    @property
    def nnodes(self) -> int:
        return int(self.fathers.shape[0])

    def ninter(self) -> np.ndarray:
        return np.diff(self.inter_ptr)

    def nbound(self) -> np.ndarray:
        return np.diff(self.bound_ptr)

    def inter(self, i: int) -> np.ndarray:
        """1-based DOF ids of node ``i`` (0-based node position)."""
        return self.inter_idx[self.inter_ptr[i]:self.inter_ptr[i + 1]]

    def bound(self, i: int) -> np.ndarray:
        return self.bound_idx[self.bound_ptr[i]:self.bound_ptr[i + 1]]

    # -- the padded matrices of the .mat schema (util/read_problem.jl:14-20) -------------------
    def to_padded(self) -> dict:
        nn = self.nnodes
        ninter = self.ninter()
        nbound = self.nbound()
        inter = np.zeros((max(int(ninter.max(initial=0)), 1), nn), dtype=np.float64)
        bound = np.zeros((max(int(nbound.max(initial=0)), 1), nn), dtype=np.float64)
        for i in range(nn):
            inter[: ninter[i], i] = self.inter(i)
            bound[: nbound[i], i] = self.bound(i)
        row = lambda v: np.asarray(v, dtype=np.float64).reshape(1, nn)
        return {
            "fathers": row(self.fathers),
            "lsons": row(self.lsons),
            "rsons": row(self.rsons),
            "ninter": row(ninter),
            "nbound": row(nbound),
            "inter": inter,
            "bound": bound,
        }

    @staticmethod
    def from_padded(d: dict) -> "ElimTree":
        vec = lambda k: np.asarray(d[k]).reshape(-1).astype(np.int64)
        fathers, lsons, rsons = vec("fathers"), vec("lsons"), vec("rsons")
        ninter, nbound = vec("ninter"), vec("nbound")
        inter = np.asarray(d["inter"]).astype(np.int64)
        bound = np.asarray(d["bound"]).astype(np.int64)
        nn = fathers.shape[0]
        if inter.ndim == 1:
            inter = inter.reshape(-1, nn)
        if bound.ndim == 1:
            bound = bound.reshape(-1, nn)
        if not (lsons.shape[0] == rsons.shape[0] == ninter.shape[0] == nbound.shape[0]
                == inter.shape[1] == bound.shape[1] == nn):
            raise ValueError("dimensions inconsistent among inputs")  # nesteddissection.jl:107
        iptr = np.zeros(nn + 1, dtype=np.int64)
        bptr = np.zeros(nn + 1, dtype=np.int64)
        np.cumsum(ninter, out=iptr[1:])
        np.cumsum(nbound, out=bptr[1:])
        iidx = np.concatenate([inter[: ninter[i], i] for i in range(nn)]) if nn else np.zeros(0, np.int64)
        bidx = np.concatenate([bound[: nbound[i], i] for i in range(nn)]) if nn else np.zeros(0, np.int64)
        return ElimTree(fathers, lsons, rsons, iptr, iidx.astype(np.int64), bptr, bidx.astype(np.int64))


@dataclass
class Problem:
    A: sp.csc_matrix
    b: np.ndarray
    elim_tree: ElimTree
    name: str = ""


# ------------------------------------------------------------------------------------------------
# .mat I/O  (util/read_problem.jl:5-25)
# ------------------------------------------------------------------------------------------------
def read_problem(path: str) -> Problem:
    """Read ``A``, ``b``, ``elim_tree`` from a MATLAB v5/v7 file (reference util/read_problem.jl:5-25)."""
    from scipy.io import loadmat

    m = loadmat(path, squeeze_me=False, struct_as_record=True)
    for k in ("A", "b", "elim_tree"):
        if k not in m:
            raise KeyError(f"{path}: variable '{k}' missing")
    et = m["elim_tree"]
    fields = {name: et[name][0, 0] for name in et.dtype.names}
    tree = ElimTree.from_padded(fields)
    A = sp.csc_matrix(m["A"])
    b = np.asarray(m["b"]).reshape(-1)
    return Problem(A, b, tree, name=path)


def write_problem(path: str, prob: Problem) -> None:
    from scipy.io import savemat

    savemat(path, {"A": sp.csc_matrix(prob.A), "b": prob.b.reshape(-1, 1),
                   "elim_tree": prob.elim_tree.to_padded()}, do_compression=True)


# ------------------------------------------------------------------------------------------------
# synthetic grids
# ------------------------------------------------------------------------------------------------
def _strides(shape):
    # x fastest: id = x + nx*(y + ny*z)
    st = [1]
    for s in shape[:-1]:
        st.append(st[-1] * s)
    return st


def _slab_ids(lo, hi, st):
    """Flattened 0-based ids of all cells of the box [lo, hi)."""
    ids = np.zeros(1, dtype=np.int64)
    for k in range(len(lo)):
        r = np.arange(lo[k], hi[k], dtype=np.int64) * st[k]
        ids = (ids[None, :] + r[:, None]).reshape(-1) if k else r
    return ids


def _shell_ids(lo, hi, shape, st):
    """Cells of box [lo,hi) having a stencil neighbour outside the box but inside the domain."""
    parts = []
    d = len(lo)
    for k in range(d):
        if hi[k] - lo[k] <= 0:
            continue
        if lo[k] > 0:
            l2, h2 = list(lo), list(hi)
            h2[k] = lo[k] + 1
            parts.append(_slab_ids(l2, h2, st))
        if hi[k] < shape[k]:
            l2, h2 = list(lo), list(hi)
            l2[k] = hi[k] - 1
            parts.append(_slab_ids(l2, h2, st))
    if not parts:
        return np.zeros(0, dtype=np.int64)
    return np.unique(np.concatenate(parts))


def grid_elimtree(shape, nmax: int = 100) -> ElimTree:
    """Geometric-bisection elimination tree for a nearest-neighbour stencil on a box grid.

    Invariants produced are the ones ``_symfact!``/``_factor_*`` rely on
    (nesteddissection.jl:42-65, factorization.jl:30-42,62-75): leaves' ``inter ∪ bound`` partition the
    DOFs, ``A[leaf.inter, outside leaf] = 0``; for a branch ``inter ∪ bound`` is the disjoint union of
    the children's ``bound``; the root's ``bound`` is empty.
    """
    shape = tuple(int(s) for s in shape)
    st = _strides(shape)
    d = len(shape)
    # breadth-first construction; node ids are 1-based positions in creation order
    boxes = [(tuple([0] * d), shape)]
    fathers = [-1]
    lsons = [-1]
    rsons = [-1]
    q = 0
    while q < len(boxes):
        lo, hi = boxes[q]
        ext = [hi[k] - lo[k] for k in range(d)]
        vol = int(np.prod(ext))
        if vol > nmax and max(ext) > 1:
            k = int(np.argmax(ext))  # split the longest side (first on ties)
            mid = lo[k] + ext[k] // 2
            hl = list(hi); hl[k] = mid
            lr = list(lo); lr[k] = mid
            for (l2, h2) in ((lo, tuple(hl)), (tuple(lr), hi)):
                boxes.append((tuple(l2), tuple(h2)))
                fathers.append(q + 1)
                lsons.append(-1)
                rsons.append(-1)
            lsons[q] = len(boxes) - 1
            rsons[q] = len(boxes)
        q += 1
    nn = len(boxes)
    inter_l, bound_l = [None] * nn, [None] * nn
    shells = [None] * nn
    for i in range(nn - 1, -1, -1):
        lo, hi = boxes[i]
        sh = _shell_ids(lo, hi, shape, st)
        shells[i] = sh
        if lsons[i] == -1:
            cells = _slab_ids(lo, hi, st)
            cells.sort()
            mask = np.isin(cells, sh, assume_unique=True)
            inter_l[i] = cells[~mask]
            bound_l[i] = cells[mask]
        else:
            cand = np.concatenate([shells[lsons[i] - 1], shells[rsons[i] - 1]])
            cand.sort()
            mask = np.isin(cand, sh, assume_unique=True)
            inter_l[i] = cand[~mask]
            bound_l[i] = cand[mask]
            shells[lsons[i] - 1] = None
            shells[rsons[i] - 1] = None
    iptr = np.zeros(nn + 1, dtype=np.int64)
    bptr = np.zeros(nn + 1, dtype=np.int64)
    np.cumsum([len(x) for x in inter_l], out=iptr[1:])
    np.cumsum([len(x) for x in bound_l], out=bptr[1:])
    iidx = np.concatenate(inter_l) + 1
    bidx = np.concatenate(bound_l) + 1 if bptr[-1] else np.zeros(0, dtype=np.int64)
    return ElimTree(np.asarray(fathers, np.int64), np.asarray(lsons, np.int64), np.asarray(rsons, np.int64),
                    iptr, iidx.astype(np.int64), bptr, bidx.astype(np.int64))


def grid_operator(shape, kind: str = "poisson", ppw: float = 10.0) -> sp.csc_matrix:
    """5-point / 7-point operator on a box grid (Dirichlet).

    ``poisson``   : L = 2d·I − Σ shifts                         (float64, SPD)
    ``helmholtz`` : L − κ²·I − iκ·diag(domain-boundary rows),  κ = k·h = 2π/ppw   (complex128,
                    complex symmetric, indefinite; first-order absorbing term on the boundary rows)
    """
    shape = tuple(int(s) for s in shape)
    d = len(shape)
    eyes = [sp.identity(s, dtype=np.float64, format="csr") for s in shape]
    L = None
    for k in range(d):
        T = sp.diags([-np.ones(shape[k] - 1), 2 * np.ones(shape[k]), -np.ones(shape[k] - 1)], [-1, 0, 1],
                     format="csr")
        term = None
        # x fastest => kron order is reversed: A = kron(I_z, kron(I_y, T_x)) + ...
        for j in range(d - 1, -1, -1):
            f = T if j == k else eyes[j]
            term = f if term is None else sp.kron(term, f, format="csr")
        L = term if L is None else L + term
    if kind == "poisson":
        return sp.csc_matrix(L)
    if kind == "helmholtz":
        kap = 2.0 * np.pi / ppw
        n = int(np.prod(shape))
        bmask = np.zeros(shape[::-1], dtype=bool)  # C-order array indexed [z][y][x]
        for k in range(d):
            ax = d - 1 - k
            sl = [slice(None)] * d
            sl[ax] = 0
            bmask[tuple(sl)] = True
            sl[ax] = shape[k] - 1
            bmask[tuple(sl)] = True
        diag = -(kap ** 2) * np.ones(n) - 1j * kap * bmask.reshape(-1)
        return sp.csc_matrix(L.astype(np.complex128) + sp.diags(diag, 0))
    raise ValueError(f"unknown kind {kind!r}")


def grid_problem(shape, kind: str = "poisson", nmax: int = 100, ppw: float = 10.0, seed: int = 123) -> Problem:
    A = grid_operator(shape, kind, ppw)
    n = A.shape[0]
    b = np.random.default_rng(seed).standard_normal(n)
    if kind == "helmholtz":
        b = b.astype(np.complex128)
    tree = grid_elimtree(shape, nmax)
    return Problem(A, b, tree, name=f"{kind}{len(shape)}d_{'x'.join(map(str, shape))}_nmax{nmax}")


def nested_dissection(A, nmax: int = 100) -> ElimTree:
    """Elimination tree (the ``elim_tree`` schema of util/read_problem.jl:14-20) for an arbitrary sparse matrix, by
    recursive graph bisection of the pattern of ``A + Aᵀ`` with METIS until a part holds at most ``nmax`` DOFs
    (``hs_nd_create``, csrc/hs_ordering.cpp).  The reference has no ordering code — its trees come with the problem
    files — so this is what makes matrices without such a file usable: ``from_elimtree(nested_dissection(A))`` →
    ``symfact`` → ``postorder`` / ``permute`` → ``factor`` as in test/rungmres.jl:15-19."""
    import ctypes as C

    from . import _lib
    A = sp.csc_matrix(A)
    if A.shape[0] != A.shape[1]:
        raise _lib.DimensionMismatch(_lib.HS_EDIM, "A must be square")
    n = A.shape[0]
    flags = _lib.HS_CSC_ZERO_BASED
    if A.indices.dtype == np.int32 and A.indptr.dtype == np.int32:
        colptr, rowval = np.ascontiguousarray(A.indptr), np.ascontiguousarray(A.indices)
        flags |= _lib.HS_CSC_INT32
    else:
        colptr, rowval = _lib.as_i64(A.indptr), _lib.as_i64(A.indices)
    h = C.c_void_p()
    _lib.check(_lib.lib.hs_nd_create(n, colptr.ctypes.data_as(C.c_void_p), rowval.ctypes.data_as(C.c_void_p), flags, 1,
                                     int(nmax), C.byref(h)))
    try:
        et = _lib.hs_elimtree()
        _lib.check(_lib.lib.hs_nd_elimtree(h, C.byref(et)))
        nn = int(et.nnodes)

        def arr(p, m):
            return np.ctypeslib.as_array(p, shape=(m,)).astype(np.int64, copy=True) if m else np.zeros(0, np.int64)
        iptr, bptr = arr(et.inter_ptr, nn + 1), arr(et.bound_ptr, nn + 1)
        out = ElimTree(arr(et.fathers, nn), arr(et.lsons, nn), arr(et.rsons, nn), iptr, arr(et.inter_idx, int(iptr[-1])),
                       bptr, arr(et.bound_idx, int(bptr[-1])))
    finally:
        _lib.lib.hs_nd_free(h)
    return out
