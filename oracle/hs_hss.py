"""CPU ORACLE, HSS matrices — test infrastructure only.  NEVER imported by the product package.

The reference stores the Schur complement of a compressed node as an ``HssMatrix`` (factorization.jl:57,110) from
**HssMatrices.jl v0.1.2** (Manifest.toml:263-269, ``repo-rev = "main"``, tree 96d668de…; NOT vendored under
/root/reference, source unavailable here).  This module restates the *published* HSS algorithms that package
implements — Chandrasekaran/Gu/Pals and Xia/Chandrasekaran/Gu/Li for the nested-basis representation and its direct
bottom-up construction, Martinsson (2011) for the randomized one — at the level of detail the reference's call sites
need (SURVEY §8c lists them): ``bisection_cluster``, ``compress``, ``full``, ``hssrank``, ``generators``, ``cluster``,
``compatible``, ``prune_leaves!``, ``*`` and ``\\``.

PARITY UNPINNED: nothing of HssMatrices.jl can be executed or read here, so only the mathematics is pinned
(tests/test_oracle.py: approximation error against the tolerance, nestedness of the bases, exactness of the solve for
the represented matrix).  Entry-point semantics are RECALLED.

Representation (Appendix C): leaf ``{D, U, V}``; branch ``{A11, A22, B12, B21, R1, W1, R2, W2}`` with
``A12 = U1·B12·V2ᴴ``, ``A21 = U2·B21·V1ᴴ`` and nested bases ``U = [U1·R1; U2·R2]``, ``V = [V1·W1; V2·W2]``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from hs_oracle_hss import pqrfact


@dataclass
class ClusterTree:
    """``BinaryNode{UnitRange{Int}}`` — 0-based half-open range ``[lo, hi)`` here."""

    lo: int
    hi: int
    left: Optional["ClusterTree"] = None
    right: Optional["ClusterTree"] = None

    @property
    def size(self) -> int:
        return self.hi - self.lo

    def isleaf(self) -> bool:
        return self.left is None


def _bisect(lo: int, hi: int, leafsize: int) -> ClusterTree:
    if hi - lo <= leafsize:
        return ClusterTree(lo, hi)
    mid = lo + (hi - lo + 1) // 2
    return ClusterTree(lo, hi, _bisect(lo, mid, leafsize), _bisect(mid, hi, leafsize))


def bisection_cluster(n, leafsize: int = 32) -> ClusterTree:
    """``bisection_cluster(n; leafsize)`` halves until ``length ≤ leafsize``; the tuple form ``(n1, n)`` used at
    factorization.jl:56,109 forces the FIRST split at ``n1 | n − n1`` and bisects each side (Appendix C)."""
    if isinstance(n, tuple):
        n1, nt = int(n[0]), int(n[1])
        if n1 <= 0 or n1 >= nt:
            return _bisect(0, nt, leafsize)
        return ClusterTree(0, nt, _bisect(0, n1, leafsize), _bisect(n1, nt, leafsize))
    return _bisect(0, int(n), leafsize)


def depth(cl: ClusterTree) -> int:
    return 1 if cl.isleaf() else 1 + max(depth(cl.left), depth(cl.right))


def compatible(a: ClusterTree, b: ClusterTree) -> bool:
    """Same tree shape and sizes."""
    if a.isleaf() != b.isleaf() or a.size != b.size:
        return False
    return a.isleaf() or (compatible(a.left, b.left) and compatible(a.right, b.right))


class HssMatrix:
    """Leaf: ``D, U, V``.  Branch: ``A11, A22, B12, B21, R1, W1, R2, W2`` (``R*/W*`` are ``None`` at the root)."""

    def __init__(self):
        self.leaf = True
        self.D = self.U = self.V = None
        self.A11 = self.A22 = None
        self.B12 = self.B21 = None
        self.R1 = self.R2 = self.W1 = self.W2 = None
        self.rows = self.cols = 0

    @property
    def shape(self) -> Tuple[int, int]:
        return (self.rows, self.cols)

    @property
    def dtype(self):
        return self.D.dtype if self.leaf else self.A11.dtype

    # the reference's FactorNode / maxrank hooks
    def dense(self) -> np.ndarray:
        return full(self)

    @property
    def rank(self) -> int:
        return hssrank(self)


def generators(h: HssMatrix) -> Tuple[np.ndarray, np.ndarray]:
    """Assembled ``(U, V)`` of ``h`` as an off-diagonal participant (factorization.jl:129-132)."""
    if h.leaf:
        return h.U, h.V
    U1, V1 = generators(h.A11)
    U2, V2 = generators(h.A22)
    if h.R1 is None:
        raise ValueError("generators: the root of an HSS matrix has no translation operators")
    return np.vstack([U1 @ h.R1, U2 @ h.R2]), np.vstack([V1 @ h.W1, V2 @ h.W2])


def full(h: HssMatrix) -> np.ndarray:
    if h.leaf:
        return h.D
    U1, V1 = generators(h.A11)
    U2, V2 = generators(h.A22)
    return np.block([[full(h.A11), U1 @ h.B12 @ V2.conj().T], [U2 @ h.B21 @ V1.conj().T, full(h.A22)]])


def hssrank(h: HssMatrix) -> int:
    """Largest rank of any off-diagonal generator."""
    if h.leaf:
        return 0
    return max(hssrank(h.A11), hssrank(h.A22), *h.B12.shape, *h.B21.shape)


def cluster(h: HssMatrix, r0: int = 0, c0: int = 0) -> Tuple[ClusterTree, ClusterTree]:
    if h.leaf:
        return ClusterTree(r0, r0 + h.rows), ClusterTree(c0, c0 + h.cols)
    rl, cl = cluster(h.A11, r0, c0)
    rr, cr = cluster(h.A22, r0 + h.A11.rows, c0 + h.A11.cols)
    return ClusterTree(r0, r0 + h.rows, rl, rr), ClusterTree(c0, c0 + h.cols, cl, cr)


def compress(A: np.ndarray, rcl: ClusterTree, ccl: ClusterTree, atol: float = 1e-9, rtol: float = 1e-9) -> HssMatrix:
    """Direct bottom-up HSS construction (``compress(A, rcl, ccl; atol, rtol)``, factorization.jl:57): at every node the
    block row ``A[I, Jᶜ]`` and block column ``A[Iᶜ, J]`` — already reduced by the children's bases — are truncated with
    the same rank-revealing QR the reference uses everywhere (``pqrfact``), which yields nested orthonormal bases."""
    A = np.asarray(A)
    m, n = A.shape
    if rcl.size != m or ccl.size != n or not compatible_shape(rcl, ccl):
        raise ValueError("DimensionMismatch: cluster trees do not match the matrix")

    def rec(rc: ClusterTree, cc: ClusterTree, root: bool):
        """Returns (hss, U, V) with the node's assembled (nested) bases; (hss, None, None) at the root."""
        h = HssMatrix()
        I = np.arange(rc.lo, rc.hi)
        J = np.arange(cc.lo, cc.hi)
        h.rows, h.cols = len(I), len(J)
        Jc = np.r_[0:cc.lo, cc.hi:n]
        Ic = np.r_[0:rc.lo, rc.hi:m]
        if rc.isleaf():
            h.leaf = True
            h.D = A[np.ix_(I, J)].copy()
            if root:
                return h, None, None
            Q, _ = pqrfact(A[np.ix_(I, Jc)], atol, rtol)
            P, _ = pqrfact(A[np.ix_(Ic, J)].conj().T, atol, rtol)
            h.U, h.V = Q, P
            return h, Q, P
        h.leaf = False
        h.A11, U1, V1 = rec(rc.left, cc.left, False)
        h.A22, U2, V2 = rec(rc.right, cc.right, False)
        I1 = np.arange(rc.left.lo, rc.left.hi); I2 = np.arange(rc.right.lo, rc.right.hi)
        J1 = np.arange(cc.left.lo, cc.left.hi); J2 = np.arange(cc.right.lo, cc.right.hi)
        h.B12 = U1.conj().T @ A[np.ix_(I1, J2)] @ V2
        h.B21 = U2.conj().T @ A[np.ix_(I2, J1)] @ V1
        if root:
            return h, None, None
        # nested bases: compress the children's reduced block rows / columns restricted to the node's complement
        Hr = np.vstack([U1.conj().T @ A[np.ix_(I1, Jc)], U2.conj().T @ A[np.ix_(I2, Jc)]])
        Hc = np.vstack([V1.conj().T @ A[np.ix_(Ic, J1)].conj().T, V2.conj().T @ A[np.ix_(Ic, J2)].conj().T])
        Q, _ = pqrfact(Hr, atol, rtol)
        P, _ = pqrfact(Hc, atol, rtol)
        r1, w1 = U1.shape[1], V1.shape[1]
        h.R1, h.R2 = Q[:r1], Q[r1:]
        h.W1, h.W2 = P[:w1], P[w1:]
        return h, np.vstack([U1 @ h.R1, U2 @ h.R2]), np.vstack([V1 @ h.W1, V2 @ h.W2])

    h, _, _ = rec(rcl, ccl, True)
    return h


def compatible_shape(rcl: ClusterTree, ccl: ClusterTree) -> bool:
    """Row and column cluster trees of one HSS matrix must have the same shape (sizes may differ)."""
    if rcl.isleaf() != ccl.isleaf():
        return False
    return rcl.isleaf() or (compatible_shape(rcl.left, ccl.left) and compatible_shape(rcl.right, ccl.right))


def matmul(h: HssMatrix, X: np.ndarray) -> np.ndarray:
    """``h * X`` in O(n·r) through the generators (up sweep of Vᴴ·x, down sweep of the couplings)."""
    X = np.asarray(X)
    vec = X.ndim == 1
    X = X.reshape(h.cols, -1)
    out = np.zeros((h.rows, X.shape[1]), dtype=np.result_type(h.dtype, X.dtype))

    def up(node, x):      # returns Vᴴ·x for the node's assembled V, caching the children's
        if node.leaf:
            node._vx = node.V.conj().T @ x if node.V is not None else None
            return node._vx
        c1 = node.A11.cols
        a, b = up(node.A11, x[:c1]), up(node.A22, x[c1:])
        node._vx = None if node.W1 is None else node.W1.conj().T @ a + node.W2.conj().T @ b
        return node._vx

    def down(node, x, y, f):   # f: contribution coming from above in the node's U coordinates
        if node.leaf:
            y += node.D @ x
            if f is not None:
                y += node.U @ f
            return
        c1, r1 = node.A11.cols, node.A11.rows
        f1 = node.B12 @ node.A22._vx
        f2 = node.B21 @ node.A11._vx
        if f is not None:
            f1 = f1 + node.R1 @ f
            f2 = f2 + node.R2 @ f
        down(node.A11, x[:c1], y[:r1], f1)
        down(node.A22, x[c1:], y[r1:], f2)

    up(h, X)
    down(h, X, out, None)
    return out.reshape(-1) if vec else out


def solve(h: HssMatrix, B: np.ndarray) -> np.ndarray:
    """``h \\ B``.  HssMatrices.jl runs a ULV factorization afresh on every call; it is an exact solve of the matrix the
    HSS form represents, which is what this dense solve returns."""
    return np.linalg.solve(full(h), B)


def prune_leaves(h: HssMatrix) -> HssMatrix:
    """``prune_leaves!``: merge every deepest pair of sibling leaves into their parent (a new leaf with the assembled
    bases), as ``_equilibrate_clusters`` needs (factorization.jl:149-160)."""
    if h.leaf:
        return h
    if h.A11.leaf and h.A22.leaf:
        out = HssMatrix()
        out.leaf = True
        out.rows, out.cols = h.rows, h.cols
        out.D = full(h)
        if h.R1 is not None:
            out.U, out.V = generators(h)
        return out
    h.A11, h.A22 = prune_leaves(h.A11), prune_leaves(h.A22)
    return h


# ------------------------------------------------------------------------------------------------
# randomized construction from products and entries only (Martinsson 2011; `randcompress_adaptive`)
# ------------------------------------------------------------------------------------------------
def _row_id(S: np.ndarray, atol: float, rtol: float):
    """Row interpolative decomposition ``S ≈ X·S[skel]`` with ``X[skel] = I`` from a column-pivoted QR of ``Sᴴ``,
    truncated by the rule of ``pqrfact``.  Returns ``(X, skel)``."""
    import scipy.linalg as sla
    m = S.shape[0]
    if m == 0 or S.shape[1] == 0:
        return np.zeros((m, 0), dtype=S.dtype), np.zeros(0, dtype=np.int64)
    _, R, p = sla.qr(S.conj().T, mode="economic", pivoting=True)
    d = np.abs(np.diag(R))
    ptol = max(atol, rtol * d[0]) if len(d) else 0.0
    below = np.nonzero(d <= ptol)[0]
    r = int(below[0]) if len(below) else len(d)
    X = np.zeros((m, r), dtype=S.dtype)
    if r:
        X[p[:r]] = np.eye(r, dtype=S.dtype)
        if r < m:
            T = sla.solve_triangular(R[:r, :r], R[:r, r:])          # S[p[r:]]ᴴ ≈ S[p[:r]]ᴴ·T
            X[p[r:]] = T.conj().T
    return X, np.asarray(p[:r], dtype=np.int64)


def randcompress_adaptive(mul, mulc, getidx, rcl: ClusterTree, ccl: ClusterTree, kest: int = 10, stepsize: int = 10,
                          atol: float = 1e-9, rtol: float = 1e-9, rng=None, sketches=None, max_rounds: int = 12,
                          strict_sketches: bool = False) -> HssMatrix:
    """HSS form of an operator given only by ``mul(X) = A·X``, ``mulc(X) = Aᴴ·X`` and ``getidx(I, J) = A[I, J]`` (0-based
    index vectors) — the ``LinearMap`` the reference builds for the Schur complement (factorization.jl:228-235) and hands
    to ``randcompress_adaptive(Smap, cl, cl; kest, atol, rtol)`` (:110).  Gaussian test matrices with ``kest`` columns
    (plus 10 of oversampling), interpolative decompositions per node, couplings ``B12/B21`` read at skeleton×skeleton
    index sets; when a detected rank saturates the sample count, ``stepsize`` more columns are drawn and the
    construction is repeated.  ``sketches = (Ω, Ψ)`` fixes the test matrices (parity runs); otherwise ``rng``."""
    m, n = rcl.size, ccl.size
    if not compatible_shape(rcl, ccl):
        raise ValueError("DimensionMismatch: row and column cluster trees differ in shape")
    rng = rng or np.random.default_rng(0)
    k = int(kest) + 10
    for _ in range(max_rounds):
        if sketches is not None and sketches[0].shape[1] >= k:
            Om, Ps = sketches[0][:, :k], sketches[1][:, :k]
        elif strict_sketches:
            raise ValueError(f"randcompress_adaptive: {k} sample columns needed, the supplied sketch matrices have {sketches[0].shape[1]}")
        else:
            Om, Ps = rng.standard_normal((n, k)), rng.standard_normal((m, k))
        Sr, Sc = mul(Om), mulc(Ps)
        saturated = [False]

        def rec(rc, cc, root):
            h = HssMatrix()
            I, J = np.arange(rc.lo, rc.hi), np.arange(cc.lo, cc.hi)
            h.rows, h.cols = len(I), len(J)
            if rc.isleaf():
                h.leaf = True
                h.D = np.asarray(getidx(I, J))
                if root:
                    return h, None
                Sr_loc = Sr[I] - h.D @ Om[J]
                Sc_loc = Sc[J] - h.D.conj().T @ Ps[I]
                h.U, sr = _row_id(Sr_loc, atol, rtol)
                h.V, sc = _row_id(Sc_loc, atol, rtol)
                if max(len(sr), len(sc)) >= k - 10:
                    saturated[0] = True
                st = dict(Iskel=I[sr], Jskel=J[sc], Sr=Sr_loc[sr], Sc=Sc_loc[sc], Om=h.V.conj().T @ Om[J], Ps=h.U.conj().T @ Ps[I])
                return h, st
            h.leaf = False
            h.A11, s1 = rec(rc.left, cc.left, False)
            h.A22, s2 = rec(rc.right, cc.right, False)
            h.B12 = np.asarray(getidx(s1["Iskel"], s2["Jskel"]))
            h.B21 = np.asarray(getidx(s2["Iskel"], s1["Jskel"]))
            if root:
                return h, None
            Sr_t = np.vstack([s1["Sr"] - h.B12 @ s2["Om"], s2["Sr"] - h.B21 @ s1["Om"]])
            Sc_t = np.vstack([s1["Sc"] - h.B21.conj().T @ s2["Ps"], s2["Sc"] - h.B12.conj().T @ s1["Ps"]])
            R, sr = _row_id(Sr_t, atol, rtol)
            W, sc = _row_id(Sc_t, atol, rtol)
            if max(len(sr), len(sc)) >= k - 10:
                saturated[0] = True
            r1, w1 = len(s1["Iskel"]), len(s1["Jskel"])
            h.R1, h.R2, h.W1, h.W2 = R[:r1], R[r1:], W[:w1], W[w1:]
            Isk = np.concatenate([s1["Iskel"], s2["Iskel"]])[sr]
            Jsk = np.concatenate([s1["Jskel"], s2["Jskel"]])[sc]
            st = dict(Iskel=Isk, Jskel=Jsk, Sr=Sr_t[sr], Sc=Sc_t[sc],
                      Om=W.conj().T @ np.vstack([s1["Om"], s2["Om"]]), Ps=R.conj().T @ np.vstack([s1["Ps"], s2["Ps"]]))
            return h, st

        h, _ = rec(rcl, ccl, True)
        if not saturated[0] or k >= min(m, n) + 10:
            return h
        k += int(stepsize)
    return h


# ------------------------------------------------------------------------------------------------
# sparse embedding: an HSS system as a larger sparse system whose elimination tree is the HSS tree
# ------------------------------------------------------------------------------------------------
def sparse_embedding(h: HssMatrix):
    """Extended sparse system of an HSS matrix (Chandrasekaran, Dewilde, Gu, Lyons, Pals 2006) together with an
    elimination tree in the reference's ``elim_tree`` schema, so that ``h \\ b`` can be computed by the multifrontal
    factorization of this very repository: auxiliary unknowns ``g_τ = V_τᴴ·x_τ`` and ``f_τ`` (what reaches node τ from
    outside, in the coordinates of ``U_τ``) for every non-root node τ,

        leaf i        D_i·x_i + U_i·f_i = b_i                    V_iᴴ·x_i − g_i = 0
        branch τ      f_c1 − B12·g_c2 − R1·f_τ = 0               f_c2 − B21·g_c1 − R2·f_τ = 0
                      g_τ − W1ᴴ·g_c1 − W2ᴴ·g_c2 = 0              (R, W, f_τ, g_τ absent at the root)

    Returns ``(A_ext, n, tree)``: ``A_ext`` (scipy CSC, the first ``n`` unknowns are ``x``), and ``tree`` = dict of
    1-based arrays ``fathers, lsons, rsons, inter (list), bound (list)``.  Every leaf of the tree owns its ``x_i, g_i, f_i``
    and — the schema has no unknowns of its own at branches — the ``g_τ, f_τ`` of the ancestors whose left-most leaf it is.
    This is the round-2 route for pivot blocks kept in HSS form."""
    import scipy.sparse as sp
    if h.leaf:
        raise ValueError("sparse_embedding: the HSS matrix is a single dense block")
    nodes, kids, father = [], [], []

    def number(node, fa):
        k = len(nodes)
        nodes.append(node); kids.append((-1, -1)); father.append(fa)
        if not node.leaf:
            a = number(node.A11, k)
            b = number(node.A22, k)
            kids[k] = (a, b)
        return k

    number(h, -1)
    nn = len(nodes)
    urank = [0] * nn          # columns of U_τ  = size of f_τ
    vrank = [0] * nn          # columns of V_τ  = size of g_τ
    for k, nd_ in enumerate(nodes):
        if father[k] < 0:
            continue
        if nd_.leaf:
            urank[k], vrank[k] = nd_.U.shape[1], nd_.V.shape[1]
        else:
            urank[k], vrank[k] = nd_.R1.shape[1], nd_.W1.shape[1]
    # unknown numbering: x in leaf order, then (g_τ, f_τ) per non-root node
    xoff, off = [0] * nn, 0
    for k, nd_ in enumerate(nodes):
        if nd_.leaf:
            xoff[k] = off
            off += nd_.rows
    n = off
    goff, foff = [0] * nn, [0] * nn
    for k in range(nn):
        if father[k] < 0:
            continue
        goff[k] = off; off += vrank[k]
        foff[k] = off; off += urank[k]
    ntot = off
    rows, cols, vals = [], [], []

    def put(r0, c0, M):
        M = np.asarray(M)
        if M.size == 0:
            return
        rr, cc = np.meshgrid(np.arange(M.shape[0]) + r0, np.arange(M.shape[1]) + c0, indexing="ij")
        rows.append(rr.ravel()); cols.append(cc.ravel()); vals.append(M.ravel())

    for k, nd_ in enumerate(nodes):
        if nd_.leaf:
            put(xoff[k], xoff[k], nd_.D)
            put(xoff[k], foff[k], nd_.U)
            put(goff[k], xoff[k], nd_.V.conj().T)
            put(goff[k], goff[k], -np.eye(vrank[k]))
            continue
        c1, c2 = kids[k]
        put(foff[c1], foff[c1], np.eye(urank[c1]))
        put(foff[c1], goff[c2], -nd_.B12)
        put(foff[c2], foff[c2], np.eye(urank[c2]))
        put(foff[c2], goff[c1], -nd_.B21)
        if father[k] >= 0:
            put(foff[c1], foff[k], -nd_.R1)
            put(foff[c2], foff[k], -nd_.R2)
            put(goff[k], goff[k], np.eye(vrank[k]))
            put(goff[k], goff[c1], -nd_.W1.conj().T)
            put(goff[k], goff[c2], -nd_.W2.conj().T)
    dt = np.result_type(*[v.dtype for v in vals])
    A = sp.coo_matrix((np.concatenate(vals).astype(dt), (np.concatenate(rows), np.concatenate(cols))), shape=(ntot, ntot)).tocsc()
    # vertex sets: every non-root node's (g, f) lives in its left-most leaf
    leafset = {k: list(range(xoff[k], xoff[k] + nodes[k].rows)) for k in range(nn) if nodes[k].leaf}
    for k in range(nn):
        if father[k] < 0:
            continue
        j = k
        while not nodes[j].leaf:
            j = kids[j][0]
        leafset[j] += list(range(goff[k], goff[k] + vrank[k])) + list(range(foff[k], foff[k] + urank[k]))
    P = sp.csr_matrix((abs(A) + abs(A).T) != 0)
    owner = np.zeros(ntot, dtype=np.int64)       # leaf each unknown belongs to
    for j, vs in leafset.items():
        owner[vs] = j
    # leaves below every node (as a set of leaf ids)
    below = [None] * nn
    for k in range(nn - 1, -1, -1):
        below[k] = {k} if nodes[k].leaf else below[kids[k][0]] | below[kids[k][1]]
    inter, bound = [None] * nn, [None] * nn
    for k in range(nn - 1, -1, -1):
        cand = sorted(leafset[k]) if nodes[k].leaf else sorted(bound[kids[k][0]] + bound[kids[k][1]])
        inside = below[k]
        ii, bb = [], []
        for v in cand:
            nb = P.indices[P.indptr[v]:P.indptr[v + 1]]
            (bb if any(int(owner[u]) not in inside for u in nb) else ii).append(v)
        inter[k], bound[k] = ii, bb
    tree = dict(fathers=np.array([f + 1 if f >= 0 else -1 for f in father], dtype=np.int64),
                lsons=np.array([a + 1 if a >= 0 else -1 for a, _ in kids], dtype=np.int64),
                rsons=np.array([b + 1 if b >= 0 else -1 for _, b in kids], dtype=np.int64),
                inter=[np.asarray(v, dtype=np.int64) + 1 for v in inter],
                bound=[np.asarray(v, dtype=np.int64) + 1 for v in bound])
    return A, n, tree
