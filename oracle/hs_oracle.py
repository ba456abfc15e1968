"""CPU ORACLE — test infrastructure only.  NEVER imported by the product package.

A NumPy/SciPy restatement of the hot path of bonevbs/HierarchicalSolvers.jl (the multifrontal
nested-dissection factorization and its tree solve) that the CUDA implementation is checked against.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.

PARITY UNPINNED.  The reference holds no tests, no golden vectors and no fixtures, and it cannot be
executed in this environment (no Julia; HssMatrices.jl / LowRankApprox.jl are not vendored).  This
restatement therefore follows the reference source line by line (citations on every function) and is
pinned only by independent mathematics: SuperLU (``scipy.sparse.linalg.splu``) solutions of the same
systems and algebraic identities of the factors (tests/test_oracle.py).

Conventions: index sets (``int``, ``bnd``, ``int_loc``, ``bnd_loc``) hold **1-based** values exactly as
the Julia code does; conversion to 0-based happens only where an array is subscripted.  Dense blocks are
column-major-agnostic ``numpy.ndarray``; every ``\\`` / ``/`` is a fresh LAPACK ``gesv`` as in the
reference (which never stores an LU, see factorization.jl:36-37, blockmatrix.jl:118,139-142).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

# ------------------------------------------------------------------------------------------------
# nesteddissection.jl
# ------------------------------------------------------------------------------------------------


class NDNode:
    """``BinaryNode{Tuple{int,bnd}}`` (nesteddissection.jl:8,19-21)."""

    __slots__ = ("int", "bnd", "left", "right")

    def __init__(self, int_, bnd, left=None, right=None):
        self.int = np.asarray(int_, dtype=np.int64).reshape(-1)
        self.bnd = np.asarray(bnd, dtype=np.int64).reshape(-1)
        self.left = left
        self.right = right


def isleaf(nd) -> bool:
    return nd.left is None and nd.right is None


def isbranch(nd) -> bool:
    return nd.left is not None and nd.right is not None


def depth(nd) -> int:
    """``HssMatrices.depth`` — height of the tree.  RECALLED semantics: a lone leaf has
    depth 1 here; only used to resolve negative ``swlevel`` (factorization.jl:8)."""
    if nd is None:
        return 0
    return 1 + max(depth(nd.left), depth(nd.right))


def parse_elimtree(fathers, lsons, rsons, ninter, inter, nbound, bound) -> NDNode:
    """nesteddissection.jl:105-148.  ``inter``/``bound`` are callables ``i -> 1-based ids`` or padded
    matrices (max×nnodes); node ids are 1-based, -1 = none."""
    fathers = np.asarray(fathers).reshape(-1)
    lsons = np.asarray(lsons).reshape(-1)
    rsons = np.asarray(rsons).reshape(-1)
    nnodes = len(fathers)
    if not (nnodes == len(lsons) == len(rsons) == len(ninter) == len(nbound)):
        raise ValueError("dimensions inconsistent among inputs")  # :107 DimensionMismatch
    get_i = inter if callable(inter) else (lambda i: np.asarray(inter)[: ninter[i - 1], i - 1])
    get_b = bound if callable(bound) else (lambda i: np.asarray(bound)[: nbound[i - 1], i - 1])
    roots = np.nonzero(fathers == -1)[0] + 1
    if len(roots) != 1:
        raise ValueError("found either less than or more than one root.")  # :111 ArgumentError
    sind = [int(roots[0])]
    ilast = -2
    snodes: List[NDNode] = []
    while sind:
        i = sind[-1]
        l, r = int(lsons[i - 1]), int(rsons[i - 1])
        if r == -1 and l == -1:  # :122 leaf
            snodes.append(NDNode(get_i(i), get_b(i)))
            ilast = sind.pop()
        elif ilast == r:  # :125 moving up from the right
            right = snodes.pop()
            left = snodes.pop() if l != -1 else None
            snodes.append(NDNode(get_i(i), get_b(i), left, right))
            ilast = sind.pop()
        elif ilast == l and r == -1:  # :134
            left = snodes.pop()
            snodes.append(NDNode(get_i(i), get_b(i), left, None))
            ilast = sind.pop()
        elif (ilast == l and r != -1) or (l == -1):  # :139 (the reference's `rsons != -1` is always true)
            ilast = i
            sind.append(r)
        else:
            ilast = i
            sind.append(l)
    return snodes.pop()


def _findall_in(values, container):
    """``findall(in(container), values)`` → 1-based positions (nesteddissection.jl:42-43)."""
    return np.nonzero(np.isin(values, container))[0].astype(np.int64) + 1


def symfact(nd: NDNode) -> Tuple[NDNode, NDNode]:
    """``symfact!`` nesteddissection.jl:29-34 (mutates ``nd``; returns ``(nd, nd_loc)``)."""
    nd_loc = _symfact(nd, 1)
    nd_loc.int = np.arange(1, len(nd.bnd) + 1, dtype=np.int64)
    nd_loc.bnd = np.zeros(0, dtype=np.int64)
    return nd, nd_loc


def _symfact(nd: NDNode, level: int) -> NDNode:
    """nesteddissection.jl:35-69."""
    empty = np.zeros(0, dtype=np.int64)
    if isleaf(nd):
        return NDNode(empty, empty)
    if nd.left is not None:
        left_loc = _symfact(nd.left, level + 1)
        left_loc.int = _findall_in(nd.left.bnd, nd.int)
        left_loc.bnd = _findall_in(nd.left.bnd, nd.bnd)
        intl = nd.left.bnd[left_loc.int - 1]
        bndl = nd.left.bnd[left_loc.bnd - 1]
    else:
        intl, bndl, left_loc = empty, empty, None
    if nd.right is not None:
        right_loc = _symfact(nd.right, level + 1)
        right_loc.int = _findall_in(nd.right.bnd, nd.int)
        right_loc.bnd = _findall_in(nd.right.bnd, nd.bnd)
        intr = nd.right.bnd[right_loc.int - 1]
        bndr = nd.right.bnd[right_loc.bnd - 1]
    else:
        intr, bndr, right_loc = empty, empty, None
    nd.int = np.concatenate([intl, intr])  # :64
    nd.bnd = np.concatenate([bndl, bndr])  # :65
    return NDNode(empty, empty, left_loc, right_loc)


def _postorder_nodes(nd):
    out = []

    def rec(x):
        if x.left is not None:
            rec(x.left)
        if x.right is not None:
            rec(x.right)
        out.append(x)

    rec(nd)
    return out


def postorder(nd: NDNode) -> np.ndarray:
    """nesteddissection.jl:73-79 — all ``int`` in post-order, then the root's ``bnd`` (1-based perm)."""
    parts = [x.int for x in _postorder_nodes(nd)] + [nd.bnd]
    return np.concatenate(parts).astype(np.int64)


def invperm(p: np.ndarray) -> np.ndarray:
    ip = np.empty_like(p)
    ip[p - 1] = np.arange(1, len(p) + 1, dtype=p.dtype)
    return ip


def permuted(nd: NDNode, perm: np.ndarray) -> NDNode:
    """``permuted!`` nesteddissection.jl:82-88 (``perm`` 1-based)."""
    if nd.left is not None:
        nd.left = permuted(nd.left, perm)
    if nd.right is not None:
        nd.right = permuted(nd.right, perm)
    nd.int = perm[nd.int - 1]
    nd.bnd = perm[nd.bnd - 1]
    return nd


def permute_matrix(A, perm: np.ndarray):
    """``permute(A, perm, perm)`` = ``A[perm, perm]`` (test/rungmres.jl:18)."""
    p0 = perm - 1
    return sp.csc_matrix(sp.csc_matrix(A)[p0][:, p0])


# ------------------------------------------------------------------------------------------------
# blockmatrix.jl
# ------------------------------------------------------------------------------------------------


@dataclass
class BlockMatrix:
    """blockmatrix.jl:5-19."""

    A11: np.ndarray
    A12: np.ndarray
    A21: np.ndarray
    A22: np.ndarray

    def __post_init__(self):
        if self.A11.shape[0] != self.A12.shape[0] or self.A11.shape[1] != self.A21.shape[1] \
                or self.A22.shape[0] != self.A21.shape[0] or self.A22.shape[1] != self.A12.shape[1]:
            raise ValueError("DimensionMismatch in BlockMatrix")  # :13-16

    @property
    def shape(self):
        return (self.A11.shape[0] + self.A22.shape[0], self.A11.shape[1] + self.A22.shape[1])

    def dense(self) -> np.ndarray:  # Matrix(B) :67-75
        return np.block([[self.A11, self.A12], [self.A21, self.A22]])


def _solve(A, B):
    """Julia ``A \\ B`` for square dense A: LU with partial pivoting (LAPACK gesv)."""
    if A.shape[0] == 0:
        return np.zeros((0,) + B.shape[1:], dtype=np.result_type(A, B))
    if B.size == 0:
        return np.zeros(B.shape, dtype=np.result_type(A, B))
    return sla.solve(A, B, check_finite=False)


def _rsolve(B, A):
    """Julia ``B / A``  = ``(A.' \\ B.').'``."""
    if A.shape[0] == 0 or B.size == 0:
        return np.zeros(B.shape, dtype=np.result_type(A, B))
    return sla.solve(A.T, B.T, check_finite=False).T


def block_mul(A: BlockMatrix, B: BlockMatrix) -> BlockMatrix:
    """blockmatrix.jl:94-98."""
    return BlockMatrix(A.A11 @ B.A11 + A.A12 @ B.A21, A.A11 @ B.A12 + A.A12 @ B.A22,
                       A.A21 @ B.A11 + A.A22 @ B.A21, A.A21 @ B.A12 + A.A22 @ B.A22)


@dataclass
class BlockFactorization:
    """blockmatrix.jl:106-108 — holds ``A11, A12, A21, S22`` *unfactored*."""

    B: BlockMatrix


def blockfactor(A: BlockMatrix) -> BlockFactorization:
    """blockmatrix.jl:115-120."""
    if A.A11.shape[0] != A.A11.shape[1] or A.A22.shape[0] != A.A22.shape[1]:
        raise ValueError("DimensionMismatch: diagonal block not square")
    S22 = A.A22 - A.A21 @ _solve(A.A11, A.A12)
    return BlockFactorization(BlockMatrix(A.A11, A.A12, A.A21, S22))


def blockldiv_inplace(F: BlockFactorization, B: np.ndarray) -> np.ndarray:
    """``blockldiv!`` blockmatrix.jl:134-144 (returns a new array, as the reference does)."""
    A = F.B
    n1 = A.A11.shape[1]
    Y = np.empty(B.shape, dtype=np.result_type(A.A11, B))
    Y[:n1] = _solve(A.A11, B[:n1])
    Y[n1:] = B[n1:] - A.A21 @ Y[:n1]
    Y[n1:] = _solve(A.A22, Y[n1:])
    Y[:n1] = Y[:n1] - _solve(A.A11, A.A12 @ Y[n1:])
    return Y


def blockrdiv_inplace(Ain: np.ndarray, F: BlockFactorization) -> np.ndarray:
    """``blockrdiv!`` blockmatrix.jl:146-156."""
    B = F.B
    m1 = B.A11.shape[0]
    Y = np.empty(Ain.shape, dtype=np.result_type(B.A11, Ain))
    Y[:, :m1] = _rsolve(Ain[:, :m1], B.A11)
    Y[:, m1:] = Ain[:, m1:] - Y[:, :m1] @ B.A12
    Y[:, m1:] = _rsolve(Y[:, m1:], B.A22)
    Y[:, :m1] = Y[:, :m1] - _rsolve(Y[:, m1:] @ B.A21, B.A11)
    return Y


def blockldiv(F: BlockFactorization, B: BlockMatrix) -> BlockMatrix:
    """blockmatrix.jl:159-172."""
    A = F.B
    B11 = _solve(A.A11, B.A11)
    B21 = B.A21 - A.A21 @ B11
    B21 = _solve(A.A22, B21)
    B11 = B11 - _solve(A.A11, A.A12 @ B21)
    B12 = _solve(A.A11, B.A12)
    B22 = B.A22 - A.A21 @ B12
    B22 = _solve(A.A22, B22)
    B12 = B12 - _solve(A.A11, A.A12 @ B22)
    return BlockMatrix(B11, B12, B21, B22)


def blockrdiv(B: BlockMatrix, F: BlockFactorization) -> BlockMatrix:
    """blockmatrix.jl:174-187."""
    A = F.B
    B11 = _rsolve(B.A11, A.A11)
    B12 = B.A12 - B11 @ A.A12
    B12 = _rsolve(B12, A.A22)
    B11 = B11 - _rsolve(B12 @ A.A21, A.A11)
    B21 = _rsolve(B.A21, A.A11)
    B22 = B.A22 - B21 @ A.A12
    B22 = _rsolve(B22, A.A22)
    B21 = B21 - _rsolve(B22 @ A.A21, A.A11)
    return BlockMatrix(B11, B12, B21, B22)


# ------------------------------------------------------------------------------------------------
# factornode.jl
# ------------------------------------------------------------------------------------------------


@dataclass
class FactorNode:
    """factornode.jl:7-39."""

    D: object
    S: np.ndarray
    L: object
    R: object
    int: np.ndarray
    bnd: np.ndarray
    int_loc: np.ndarray
    bnd_loc: np.ndarray
    left: Optional["FactorNode"] = None
    right: Optional["FactorNode"] = None

    def D_dense(self):
        return _blockfact_dense(self.D) if isinstance(self.D, BlockFactorization) else self.D

    def L_dense(self):
        return self.L.dense() if hasattr(self.L, "dense") else self.L

    def R_dense(self):
        return self.R.dense() if hasattr(self.R, "dense") else self.R


def _blockfact_dense(F: BlockFactorization) -> np.ndarray:
    """The matrix a ``BlockFactorization`` represents: ``[A11 A12; A21 S22 + A21·A11⁻¹·A12]``."""
    A = F.B
    A22 = A.A22 + A.A21 @ _solve(A.A11, A.A12)
    return np.block([[A.A11, A.A12], [A.A21, A22]])


def maxrank(F: FactorNode) -> int:
    """factornode.jl:49-57 — 0 when nothing is compressed."""
    rkl = maxrank(F.left) if F.left is not None else 0
    rkr = maxrank(F.right) if F.right is not None else 0
    rk = 0
    for X in (F.S, F.L, F.R):
        r = getattr(X, "rank", None)
        if r is not None:
            rk = max(rk, int(r() if callable(r) else r))
    return max(rkl, rkr, rk)


def nodes_postorder(F: FactorNode) -> List[FactorNode]:
    return _postorder_nodes(F)


# ------------------------------------------------------------------------------------------------
# factorization.jl — uncompressed path
# ------------------------------------------------------------------------------------------------


def _sub(A, rows, cols) -> np.ndarray:
    """``Matrix(view(A, rows, cols))`` with 1-based index vectors (factorization.jl:33-40,118-121)."""
    if len(rows) == 0 or len(cols) == 0:
        return np.zeros((len(rows), len(cols)), dtype=A.dtype)
    return np.asarray(A[rows - 1][:, cols - 1].todense())


def factor(A, nd: NDNode, nd_loc: NDNode, swlevel: int = 0, **opts) -> FactorNode:
    """factorization.jl:5-11.  ``swlevel = 0`` (nothing compressed) is handled in this file; any other value goes
    to the compressed path in ``oracle/hs_oracle_hss.py``."""
    if swlevel != 0:
        import hs_oracle_hss
        return hs_oracle_hss.factor(A, nd, nd_loc, swlevel=swlevel, **opts)
    A = sp.csr_matrix(A)
    return _factor(A, nd, nd_loc, 1)


def _factor(A, nd, nd_loc, level) -> FactorNode:
    """factorization.jl:14-27."""
    if isleaf(nd):
        return _factor_leaf(A, nd, nd_loc)
    elif isbranch(nd):
        Fl = _factor(A, nd.left, nd_loc.left, level + 1)
        Fr = _factor(A, nd.right, nd_loc.right, level + 1)
        return _factor_branch(A, Fl, Fr, nd, nd_loc)
    raise RuntimeError("Expected nested dissection to be a binary tree. Found a node with only one child.")


def _factor_leaf(A, nd, nd_loc) -> FactorNode:
    """factorization.jl:30-42."""
    int_, bnd = nd.int, nd.bnd
    D = _sub(A, int_, int_)
    Abi = _sub(A, bnd, int_)
    L = _rsolve(Abi, D)
    R = _solve(D, _sub(A, int_, bnd))
    perm = np.concatenate([nd_loc.int, nd_loc.bnd]) - 1
    S = _sub(A, bnd, bnd) - Abi @ R
    return FactorNode(D, S[np.ix_(perm, perm)], L, R, int_, bnd, nd_loc.int, nd_loc.bnd)


def _assemble_blocks(A, S1, S2, int1, int2, bnd1, bnd2):
    """factorization.jl:115-123."""
    ni1, nb1, ni2, nb2 = len(int1), len(bnd1), len(int2), len(bnd2)
    Aii = BlockMatrix(S1[:ni1, :ni1], _sub(A, int1, int2), _sub(A, int2, int1), S2[:ni2, :ni2])
    Aib = BlockMatrix(S1[:ni1, ni1:ni1 + nb1], _sub(A, int1, bnd2), _sub(A, int2, bnd1), S2[:ni2, ni2:ni2 + nb2])
    Abi = BlockMatrix(S1[ni1:ni1 + nb1, :ni1], _sub(A, bnd1, int2), _sub(A, bnd2, int1), S2[ni2:ni2 + nb2, :ni2])
    Abb = BlockMatrix(S1[ni1:ni1 + nb1, ni1:ni1 + nb1], _sub(A, bnd1, bnd2), _sub(A, bnd2, bnd1),
                      S2[ni2:ni2 + nb2, ni2:ni2 + nb2])
    return Aii, Aib, Abi, Abb


def _factor_branch(A, Fl, Fr, nd, nd_loc) -> FactorNode:
    """factorization.jl:62-75."""
    int1 = nd.left.bnd[nd_loc.left.int - 1]
    bnd1 = nd.left.bnd[nd_loc.left.bnd - 1]
    int2 = nd.right.bnd[nd_loc.right.int - 1]
    bnd2 = nd.right.bnd[nd_loc.right.bnd - 1]
    Aii, Aib, Abi, Abb = _assemble_blocks(A, Fl.S, Fr.S, int1, int2, bnd1, bnd2)
    D = blockfactor(Aii)
    L = blockrdiv(Abi, D)
    R = blockldiv(D, Aib)
    S = Abb.dense() - block_mul(Abi, R).dense()
    perm = np.concatenate([nd_loc.int, nd_loc.bnd]) - 1
    return FactorNode(D, S[np.ix_(perm, perm)], L, R, nd.int, nd.bnd, nd_loc.int, nd_loc.bnd, Fl, Fr)


# ------------------------------------------------------------------------------------------------
# factornode.jl — ldiv!
# ------------------------------------------------------------------------------------------------


def ldiv(F: FactorNode, B: np.ndarray) -> np.ndarray:
    """``ldiv!(C, F, B)`` factornode.jl:62-74.  Returns ``C`` (same shape as ``B``)."""
    vec = B.ndim == 1
    C = np.array(B.reshape(len(B), -1), dtype=np.result_type(F.S, B), copy=True)
    _lsolve(F, C)
    _dsolve(F, C)
    if len(F.bnd):
        C[F.bnd - 1] = _solve(_S_dense(F.S), C[F.bnd - 1])  # :72
    _rsolve_tree(F, C)
    return C.reshape(-1) if vec else C


def _S_dense(S):
    return S.dense() if hasattr(S, "dense") else S


def _apply(M, X):
    if hasattr(M, "matmul"):  # low-rank operators of the compressed path
        return M.matmul(X)
    return (M.dense() if isinstance(M, BlockMatrix) else M) @ X


def _lsolve(F, rhs):
    """factornode.jl:77-82."""
    if F.left is not None:
        _lsolve(F.left, rhs)
    if F.right is not None:
        _lsolve(F.right, rhs)
    rhs[F.bnd - 1] = rhs[F.bnd - 1] - _apply(F.L, rhs[F.int - 1])


def _rsolve_tree(F, rhs):
    """factornode.jl:83-88."""
    rhs[F.int - 1] = rhs[F.int - 1] - _apply(F.R, rhs[F.bnd - 1])
    if F.left is not None:
        _rsolve_tree(F.left, rhs)
    if F.right is not None:
        _rsolve_tree(F.right, rhs)


def _dsolve(F, rhs):
    """factornode.jl:89-99."""
    if F.left is not None:
        _dsolve(F.left, rhs)
    if F.right is not None:
        _dsolve(F.right, rhs)
    if isinstance(F.D, BlockFactorization):
        rhs[F.int - 1] = blockldiv_inplace(F.D, rhs[F.int - 1])
    elif hasattr(F.D, "solve"):
        rhs[F.int - 1] = F.D.solve(rhs[F.int - 1])
    else:
        rhs[F.int - 1] = _solve(F.D, rhs[F.int - 1])


# ------------------------------------------------------------------------------------------------
# GMRES as the reference's driver uses it (test/rungmres.jl:47-48; IterativeSolvers.jl 0.9.0 semantics,
# RECALLED: x0 = 0, right preconditioner, modified Gram-Schmidt, restart cycles, stop on the running
# residual estimate ≤ reltol·‖b‖, at most ``maxiter`` Arnoldi steps in total).
# ------------------------------------------------------------------------------------------------


def gmres(A, b, Pr=None, reltol=1e-9, restart=30, maxiter=30):
    """Returns ``(x, resnorms, converged)``; ``Pr`` is a callable ``v -> Pr⁻¹ v`` (the role ``ldiv!`` plays)."""
    n = len(b)
    dtype = np.result_type(A.dtype, b.dtype, np.float64)
    x = np.zeros(n, dtype=dtype)
    P = (lambda v: v) if Pr is None else Pr
    r = b.astype(dtype)  # x0 = 0
    beta = np.linalg.norm(r)
    tol = reltol * beta
    res = []
    it = 0
    resid = beta
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
            for j in range(k + 1):  # modified Gram-Schmidt
                H[j, k] = np.vdot(V[:, j], w)
                w = w - H[j, k] * V[:, j]
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] != 0:
                V[:, k + 1] = w / H[k + 1, k]
            for j in range(k):  # previous Givens rotations
                t = cs[j] * H[j, k] + sn[j] * H[j + 1, k]
                H[j + 1, k] = -np.conj(sn[j]) * H[j, k] + cs[j] * H[j + 1, k]
                H[j, k] = t
            a, c = H[k, k], H[k + 1, k]
            den = np.sqrt(abs(a) ** 2 + abs(c) ** 2)
            if den == 0:
                cs[k], sn[k] = 1.0, 0.0
            else:
                cs[k] = abs(a) / den if a != 0 else 0.0
                sn[k] = (a / abs(a)) * np.conj(c) / den if a != 0 else 1.0
            H[k, k] = cs[k] * a + sn[k] * c
            H[k + 1, k] = 0.0
            g[k + 1] = -np.conj(sn[k]) * g[k]
            g[k] = cs[k] * g[k]
            resid = abs(g[k + 1])
            res.append(float(resid))
            k += 1
            it += 1
        y = sla.solve_triangular(H[:k, :k], g[:k]) if k else np.zeros(0, dtype=dtype)
        x = x + P(V[:, :k] @ y)
        if it < maxiter and resid > tol:
            r = b - A @ x
            beta = np.linalg.norm(r)
            resid = beta
    return x, res, bool(resid <= tol)


# ------------------------------------------------------------------------------------------------
# convenience: the whole reference driver (test/rungmres.jl:15-19,32,47) on an in-memory problem
# ------------------------------------------------------------------------------------------------


def tree_from_elimtree(et) -> NDNode:
    """``read_problem`` tail (util/read_problem.jl:14-24) for a ragged ``ElimTree``-like object with fields
    fathers/lsons/rsons/inter_ptr/inter_idx/bound_ptr/bound_idx."""
    ninter = np.diff(et.inter_ptr)
    nbound = np.diff(et.bound_ptr)
    gi = lambda i: et.inter_idx[et.inter_ptr[i - 1]:et.inter_ptr[i]]
    gb = lambda i: et.bound_idx[et.bound_ptr[i - 1]:et.bound_ptr[i]]
    return parse_elimtree(et.fathers, et.lsons, et.rsons, ninter, gi, nbound, gb)


def prepare(A, et):
    """rungmres.jl:15-19: parse, ``symfact!``, post-order permutation of ``A`` and of the tree."""
    nd = tree_from_elimtree(et)
    nd, nd_loc = symfact(nd)
    perm = postorder(nd)
    Ap = permute_matrix(A, perm)
    nd = permuted(nd, invperm(perm))
    return Ap, nd, nd_loc, perm
