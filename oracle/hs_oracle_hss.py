"""CPU ORACLE, compressed path — test infrastructure only.  NEVER imported by the product package.

Restatement of the reference's ``compression_flag == true`` branch (src/factorization.jl:78-112) in the form the
CUDA library implements in this round:

* the Gauss transforms ``L ≈ Abi·Aii⁻¹`` and ``R ≈ Aii⁻¹·Aib`` are low-rank, from a truncated column-pivoted QR of
  the dense coupling blocks (dense-children methods, factorization.jl:171-182) — followed line by line;
* the Schur complement operator ``S = Abb − (Abi·R.U)·R.Vᴴ`` (factorization.jl:228-249) is *evaluated densely and
  kept dense* by default (``hss=False``, the form the CUDA library implements; ``hss=True`` / ``hss="rand"`` store the
  HSS approximation through oracle/hs_hss.py — the second by the reference's matrix-free randomized route).  The reference hands that operator to ``HssMatrices.randcompress_adaptive`` (factorization.jl:110,
  third party, source absent) and stores the HSS approximation; here the HSS tolerance is taken to zero.  Because
  every ``S`` stays dense, the HSS-children methods (``_assemble_blocks`` :126-140, ``_equilibrate_clusters``
  :143-168, all-HSS ``blockfactor`` blockmatrix.jl:121-130, HSS Gauss transforms :184-209) are never reached.
* compressed leaves (factorization.jl:45-59) keep dense ``L``, ``R`` exactly like the reference; only their ``S`` would
  be HSS there — with a dense ``S`` they coincide with the uncompressed leaf.

PARITY UNPINNED (see hs_oracle.py).  ``pqrfact`` is LowRankApprox.jl 0.4.3 (Manifest.toml:427-431, not vendored):
its truncation rule is RECALLED: partial column-pivoted QR ``A[:,p] ≈ Q·R`` cut at the first
``k`` with ``|R[k,k]| ≤ max(atol, rtol·|R[1,1]|)``.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

import hs_oracle as base
from hs_oracle import (BlockMatrix, FactorNode, _assemble_blocks, _factor_leaf, blockfactor, blockldiv_inplace,
                        blockrdiv_inplace, isbranch, isleaf)


class LowRankMatrix:
    """``LowRankApprox.LowRankMatrix(U, V)`` ≡ ``U·Vᴴ``."""

    def __init__(self, U, V):
        self.U = np.asarray(U)
        self.V = np.asarray(V)

    @property
    def rank(self) -> int:
        return self.U.shape[1]

    @property
    def shape(self):
        return (self.U.shape[0], self.V.shape[0])

    def matmul(self, X):
        return self.U @ (self.V.conj().T @ X)

    def dense(self):
        return self.U @ self.V.conj().T


def pqrfact(M: np.ndarray, atol: float, rtol: float):
    """``pqrfact(M; sketch=:none, atol, rtol)`` → ``(Q, R[:, invperm(p)])`` with ``M ≈ Q·R[:, invperm(p)]``."""
    m, n = M.shape
    if m == 0 or n == 0:
        return np.zeros((m, 0), dtype=M.dtype), np.zeros((0, n), dtype=M.dtype)
    Q, R, p = sla.qr(M, mode="economic", pivoting=True)
    d = np.abs(np.diag(R))
    ptol = max(atol, rtol * d[0]) if len(d) else 0.0
    below = np.nonzero(d <= ptol)[0]
    k = int(below[0]) if len(below) else len(d)
    Rk = np.zeros((k, n), dtype=R.dtype)
    Rk[:, p] = R[:k]          # R[:, invperm(p)]
    return Q[:, :k], Rk


def _lgauss_transform(D, Abi: BlockMatrix, atol, rtol) -> LowRankMatrix:
    """factorization.jl:171-176."""
    Q, Rp = pqrfact(Abi.dense(), atol, rtol)
    L = LowRankMatrix(Q, Rp.conj().T)
    L.V = blockrdiv_inplace(L.V.conj().T, D).conj().T if L.rank else L.V
    return L


def _rgauss_transform(D, Aib: BlockMatrix, atol, rtol) -> LowRankMatrix:
    """factorization.jl:177-182."""
    Q, Rp = pqrfact(Aib.dense(), atol, rtol)
    R = LowRankMatrix(Q, Rp.conj().T)
    R.U = blockldiv_inplace(D, R.U) if R.rank else R.U
    return R


def _S(F):
    """Dense view of a child's Schur complement (an ``HssMatrix`` in ``hss`` mode)."""
    return F.S.dense() if hasattr(F.S, "dense") else F.S


def _to_hss(S, n1, leafsize, atol, rtol):
    """``compress(S[perm,perm], cl, cl)`` / ``randcompress_adaptive(Smap, cl, cl)`` with
    ``cl = bisection_cluster((n1, n); leafsize)`` (factorization.jl:56-57,109-110): the first split separates what the
    parent eliminates from what it passes up.  The randomized construction of the reference and this direct one
    approximate the same operator to the same tolerance; they do not produce identical generators."""
    import hs_hss
    cl = hs_hss.bisection_cluster((n1, S.shape[0]), leafsize)
    return hs_hss.compress(S, cl, cl, atol, rtol)


def _schur_complement(Abb, Abi, R: "LowRankMatrix", perm):
    """factorization.jl:228-249: the Schur complement ``S[perm,perm]`` as products and entry evaluation only.
    Returns ``(Smul, Smulc, Sidx)``; index arguments of ``Sidx`` are 0-based vectors."""
    iperm = np.argsort(perm)
    UU, UV = Abi @ R.U, R.V                                  # :230  U = Abi*R  (LowRankMatrix)

    def Smul(x):                                             # :239-244  _sample_schur!
        z = x[iperm]
        y = np.empty((len(perm), x.shape[1]), dtype=np.result_type(Abb, x))
        y[iperm] = Abb @ z - UU @ (UV.conj().T @ z)
        return y

    def Smulc(x):                                            # :233  the same with Abb', U'
        z = x[iperm]
        y = np.empty((len(perm), x.shape[1]), dtype=np.result_type(Abb, x))
        y[iperm] = Abb.conj().T @ z - UV @ (UU.conj().T @ z)
        return y

    def Sidx(i, j):                                          # :246-249  _getindex_schur
        ii, jj = perm[i], perm[j]
        return Abb[np.ix_(ii, jj)] - UU[ii] @ UV[jj].conj().T

    return Smul, Smulc, Sidx


def _equilibrate_clusters(S1, S2):
    """factorization.jl:143-168: prune HSS leaf levels of ``S1.A11`` / ``S2.A11`` until their cluster trees are compatible
    (the all-HSS ``blockfactor`` needs matching block structure).  Returns the (possibly pruned) pair."""
    import hs_hss as H
    r1, r2 = H.cluster(S1.A11)[0], H.cluster(S2.A11)[0]
    guard = 0
    while not H.compatible(r1, r2) and guard < 64:
        d1, d2 = H.depth(r1), H.depth(r2)
        if d1 >= d2 and not S1.A11.leaf:
            S1.A11 = H.prune_leaves(S1.A11)
        if d2 >= d1 and not S2.A11.leaf:
            S2.A11 = H.prune_leaves(S2.A11)
        r1, r2 = H.cluster(S1.A11)[0], H.cluster(S2.A11)[0]
        guard += 1
        if S1.A11.leaf and S2.A11.leaf:
            break     # sizes differ: nothing left to prune (the reference's `compatible` also compares sizes)
    return S1, S2


def _hss_children_blocks(A, S1, S2, int1, int2, bnd1, bnd2):
    """``_assemble_blocks`` for HSS children (factorization.jl:126-140): diagonal blocks from ``S.A11`` / ``S.A22``, the
    children's off-diagonal blocks as low-rank pairs read off the generators, sparse couplings from ``A``."""
    import hs_hss as H
    Ui1, Vi1 = H.generators(S1.A11); Ui1 = Ui1 @ S1.B12          # :129
    Ui2, Vi2 = H.generators(S2.A11); Ui2 = Ui2 @ S2.B12          # :130
    Ub1, Vb1 = H.generators(S1.A22); Ub1 = Ub1 @ S1.B21          # :131
    Ub2, Vb2 = H.generators(S2.A22); Ub2 = Ub2 @ S2.B21          # :132
    sub = base._sub
    Aii = BlockMatrix(H.full(S1.A11), sub(A, int1, int2), sub(A, int2, int1), H.full(S2.A11))             # :135
    lr = dict(ib1=LowRankMatrix(Ui1, Vb1), ib2=LowRankMatrix(Ui2, Vb2), bi1=LowRankMatrix(Ub1, Vi1), bi2=LowRankMatrix(Ub2, Vi2))
    Aib = BlockMatrix(lr["ib1"].dense(), sub(A, int1, bnd2), sub(A, int2, bnd1), lr["ib2"].dense())      # :136
    Abi = BlockMatrix(lr["bi1"].dense(), sub(A, bnd1, int2), sub(A, bnd2, int1), lr["bi2"].dense())      # :137
    Abb = BlockMatrix(H.full(S1.A22), sub(A, bnd1, bnd2), sub(A, bnd2, bnd1), H.full(S2.A22))             # :138
    return Aii, Aib, Abi, Abb, lr


def _blkdiag(X, Y):
    out = np.zeros((X.shape[0] + Y.shape[0], X.shape[1] + Y.shape[1]), dtype=np.result_type(X, Y))
    out[:X.shape[0], :X.shape[1]] = X
    out[X.shape[0]:, X.shape[1]:] = Y
    return out


def _gauss_transforms_hss_children(D, Aib, Abi, lr, atol, rtol):
    """factorization.jl:184-209: the children's low-rank blocks are concatenated without recompression (``_recompress!``
    is dead code, :251-259), the sparse anti-diagonal couplings are compressed by ``pqrfact(…; sketch=:randn)`` — emulated
    by the unsketched ``pqrfact`` — and appended; note the extra ``0.5×`` of the right transform (:202)."""
    L = LowRankMatrix(_blkdiag(lr["bi1"].U, lr["bi2"].U), _blkdiag(lr["bi1"].V, lr["bi2"].V))          # :185
    X = np.block([[np.zeros_like(Abi.A11), Abi.A12], [Abi.A21, np.zeros_like(Abi.A22)]])
    if np.count_nonzero(Abi.A12) + np.count_nonzero(Abi.A21) > 0:                                      # :186
        Q, Rp = pqrfact(X, atol, rtol)                                                                  # :189
        L.U, L.V = np.hstack([L.U, Q]), np.hstack([L.V, Rp.conj().T])
    L.V = blockrdiv_inplace(L.V.conj().T, D).conj().T if L.rank else L.V                                # :193
    R = LowRankMatrix(_blkdiag(lr["ib1"].U, lr["ib2"].U), _blkdiag(lr["ib1"].V, lr["ib2"].V))          # :198
    X = np.block([[np.zeros_like(Aib.A11), Aib.A12], [Aib.A21, np.zeros_like(Aib.A22)]])
    if np.count_nonzero(Aib.A12) + np.count_nonzero(Aib.A21) > 0:                                      # :199
        Q, Rp = pqrfact(X, 0.5 * atol, 0.5 * rtol)                                                      # :202
        R.U, R.V = np.hstack([R.U, Q]), np.hstack([R.V, Rp.conj().T])
    R.U = blockldiv_inplace(D, R.U) if R.rank else R.U                                                  # :206
    return L, R


def _factor_branch_compressed(A, Fl, Fr, nd, nd_loc, atol, rtol, hss=False, leafsize=32, kest=-1, stepsize=10, sketches=None) -> FactorNode:
    """factorization.jl:78-112 with the Schur operator of :228-249 evaluated densely (module docstring).  ``hss=True``
    stores its HSS approximation (oracle/hs_hss.py) like the reference; parents then assemble from the approximated
    blocks, which is what the reference's HSS-children methods (:126-140, blockmatrix.jl:121-130) do in HSS arithmetic."""
    int1 = nd.left.bnd[nd_loc.left.int - 1]
    bnd1 = nd.left.bnd[nd_loc.left.bnd - 1]
    int2 = nd.right.bnd[nd_loc.right.int - 1]
    bnd2 = nd.right.bnd[nd_loc.right.bnd - 1]
    def _split(S):   # an HSS Schur complement whose first split separates the parent's int from its bnd
        return hasattr(S, "dense") and not S.leaf
    if hss and _split(Fl.S) and _split(Fr.S) and Fl.S.A11.rows == len(int1) and Fr.S.A11.rows == len(int2):
        # both children are HSS: the HSS methods by dispatch (:86-91, :126-140, :184-209)
        S1, S2 = _equilibrate_clusters(Fl.S, Fr.S)                                                  # :86-90
        Aii, Aib, Abi, Abb, lr = _hss_children_blocks(A, S1, S2, int1, int2, bnd1, bnd2)
        D = blockfactor(Aii)          # :94 (all-HSS method blockmatrix.jl:121-130, its recompress! tolerance → 0)
        L, R = _gauss_transforms_hss_children(D, Aib, Abi, lr, 0.5 * atol, 0.5 * rtol)             # :99-100
    else:
        Aii, Aib, Abi, Abb = _assemble_blocks(A, _S(Fl), _S(Fr), int1, int2, bnd1, bnd2)
        D = blockfactor(Aii)                                   # :94
        L = _lgauss_transform(D, Abi, 0.5 * atol, 0.5 * rtol)  # :99
        R = _rgauss_transform(D, Aib, 0.5 * atol, 0.5 * rtol)  # :100
    U = Abi.dense() @ R.U                                  # :230  U = Abi*R
    S = Abb.dense() - U @ R.V.conj().T                     # :242,:248
    perm = np.concatenate([nd_loc.int, nd_loc.bnd]) - 1    # :107
    Sp = S[np.ix_(perm, perm)]
    if hss == "rand" and len(perm):
        # the reference's own route (:102-110): matrix-free operator + randomized adaptive HSS construction
        import hs_hss
        if kest < 0:
            kest = int(np.ceil(0.5 * L.rank))                                             # :102-104
        Smul, Smulc, Sidx = _schur_complement(Abb.dense(), Abi.dense(), R, perm)         # :108
        cl = hs_hss.bisection_cluster((len(nd_loc.int), len(perm)), leafsize)             # :109
        # `sketches = (Ω, Ψ)`: host-supplied Gaussian test matrices shared with the CUDA library (the parity anchor of
        # SURVEY §8c): a node with m boundary rows and k samples uses Ω[:m, :k], Ψ[:m, :k]
        sk = None if sketches is None else (sketches[0][:len(perm)], sketches[1][:len(perm)])
        if cl.isleaf():
            Sp = hs_hss.HssMatrix(); Sp.leaf = True; Sp.rows = Sp.cols = len(perm)
            Sp.D = np.asarray(Sidx(np.arange(len(perm)), np.arange(len(perm))))            # a single dense block
        else:
            Sp = hs_hss.randcompress_adaptive(Smul, Smulc, Sidx, cl, cl, kest=kest, stepsize=stepsize, atol=atol, rtol=rtol,
                                              rng=np.random.default_rng(len(perm)), sketches=sk, strict_sketches=sk is not None)  # :110
    elif hss and len(perm):
        Sp = _to_hss(Sp, len(nd_loc.int), leafsize, atol, rtol)   # :109-110 by the direct construction
    return FactorNode(D, Sp, L, R, nd.int, nd.bnd, nd_loc.int, nd_loc.bnd, Fl, Fr)


def factor(A, nd, nd_loc, swlevel=5, swsize=1, atol=1e-6, rtol=1e-6, leafsize=32, kest=-1, stepsize=10, hss=False,
           sketches=None, **_unused) -> FactorNode:
    """factorization.jl:5-11 — options as ``SolverOptions`` (HierarchicalSolvers.jl:30-40 defaults).  ``hss=False`` (what
    the CUDA library implements this round) keeps every Schur complement dense; ``hss=True`` stores the HSS
    approximation of compressed nodes' Schur complements (direct construction), ``hss="rand"`` builds it the reference's
    way: the matrix-free operator of :228-249 handed to the randomized adaptive construction with ``kest``/``stepsize``."""
    A = sp.csr_matrix(A)
    sw = max(base.depth(nd) + swlevel, 0) if swlevel < 0 else swlevel   # :8
    return _factor(A, nd, nd_loc, 1, sw, swsize, atol, rtol, hss, leafsize, kest, stepsize, sketches)


def _factor(A, nd, nd_loc, level, swlevel, swsize, atol, rtol, hss=False, leafsize=32, kest=-1, stepsize=10, sketches=None) -> FactorNode:
    """factorization.jl:14-27."""
    compression_flag = level <= swlevel and len(nd.bnd) >= swsize   # :15
    if isleaf(nd):
        F = _factor_leaf(A, nd, nd_loc)        # :30-42; the compressed leaf (:45-59) differs only in the storage of S
        if compression_flag and hss and F.S.shape[0]:
            F.S = _to_hss(F.S, len(nd_loc.int), leafsize, atol, rtol)   # :56-57
        return F
    elif isbranch(nd):
        Fl = _factor(A, nd.left, nd_loc.left, level + 1, swlevel, swsize, atol, rtol, hss, leafsize, kest, stepsize, sketches)
        Fr = _factor(A, nd.right, nd_loc.right, level + 1, swlevel, swsize, atol, rtol, hss, leafsize, kest, stepsize, sketches)
        if compression_flag:
            return _factor_branch_compressed(A, Fl, Fr, nd, nd_loc, atol, rtol, hss, leafsize, kest, stepsize, sketches)
        if hss and (hasattr(Fl.S, "dense") or hasattr(Fr.S, "dense")):
            # an uncompressed node above compressed children (always the root, factorization.jl:15 with |bnd| = 0)
            class _V:      # children viewed through their dense Schur blocks
                def __init__(self, F): self.S = _S(F)
            F = base._factor_branch(A, _V(Fl), _V(Fr), nd, nd_loc)
            F.left, F.right = Fl, Fr
            return F
        return base._factor_branch(A, Fl, Fr, nd, nd_loc)
    raise RuntimeError("Expected nested dissection to be a binary tree. Found a node with only one child.")


def node_ranks(F: FactorNode):
    """(rank(L), rank(R)) per node in post-order; (0, 0) for uncompressed nodes."""
    out = []
    for X in base.nodes_postorder(F):
        out.append((X.L.rank if isinstance(X.L, LowRankMatrix) else 0, X.R.rank if isinstance(X.R, LowRankMatrix) else 0))
    return out
