"""Nested-dissection tree and symbolic factorization — host side of reference src/nesteddissection.jl.

The reference keeps a pointer tree of ``BinaryNode{Tuple{int,bnd}}``; here the tree is flat (CSR-style ragged
arrays, nodes numbered in post-order) because that is what the level-batched device plan consumes, and the
heavy lifting (``symfact!`` and friends) runs in C++ inside libhsolve_cuda (``hs_symfact``).  Index *values*
(DOF ids, positions) are 1-based exactly as in the ``.mat`` files and in Julia; only array subscripts are Python's.

    nd            = parse_elimtree(fathers, lsons, rsons, ninter, inter, nbound, bound)   # :105-148
    nd, nd_loc    = symfact(nd)                                                          # :29-69  (symfact!)
    perm          = postorder(nd)                                                        # :73-79
    A             = permute(A, perm, perm)                                               # SparseArrays.permute
    nd            = permuted(nd, invperm(perm))                                          # :82-88  (permuted!)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import scipy.sparse as sp

from . import _lib
from .problems import ElimTree

__all__ = ["NestedDissection", "NDNodeView", "parse_elimtree", "symfact", "postorder", "permuted", "invperm", "permute",
           "contigious", "getinterior", "getboundary", "depth", "isleaf", "isbranch"]


class NestedDissection:
    """Flat nested-dissection tree.  Before ``symfact`` it only wraps the serialized elimination tree; after it,
    nodes are numbered 0..nnodes-1 in post-order (root last) and carry ``int``/``bnd`` (1-based global ids)."""

    def __init__(self, elim: Optional[ElimTree] = None):
        self.elim = elim
        self.analyzed = False
        self.nnodes = 0 if elim is None else elim.nnodes
        self.left = self.right = None
        self.int_ptr = self.int_idx = self.bnd_ptr = self.bnd_idx = None
        self.depth = 0
        self.root = -1
        self._perm = None

    # -- reference-style access ------------------------------------------------------------------
    def node(self, k: Optional[int] = None) -> "NDNodeView":
        self._need_analyzed()
        return NDNodeView(self, self.root if k is None else k)

    @property
    def int(self):  # root properties, like `nd.int` on the reference root node
        return self.node().int

    @property
    def bnd(self):
        return self.node().bnd

    def _need_analyzed(self):
        if not self.analyzed:
            raise RuntimeError("call symfact(nd) first: the flat tree is built by the symbolic phase")

    def copy(self) -> "NestedDissection":
        out = NestedDissection(self.elim)
        out.__dict__.update({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()})
        return out


class NDLoc:
    """``nd_loc`` of ``symfact!``: per node, the positions inside its own ``bnd`` that become the parent's interior
    (``int``) and boundary (``bnd``) — nesteddissection.jl:42-43,31-32."""

    def __init__(self, nd: NestedDissection, iloc_ptr, iloc_idx, bloc_ptr, bloc_idx):
        self.nd = nd
        self.iloc_ptr, self.iloc_idx, self.bloc_ptr, self.bloc_idx = iloc_ptr, iloc_idx, bloc_ptr, bloc_idx

    def node(self, k: Optional[int] = None) -> "NDNodeView":
        return NDNodeView(self.nd, self.nd.root if k is None else k, loc=self)

    @property
    def int(self):
        return self.node().int

    @property
    def bnd(self):
        return self.node().bnd


class NDNodeView:
    """A node of either tree with ``.int .bnd .left .right`` like the reference's ``NestedDissection`` nodes."""

    def __init__(self, nd: NestedDissection, k: int, loc: Optional[NDLoc] = None):
        self.nd, self.k, self.loc = nd, int(k), loc

    @property
    def int(self):
        if self.loc is not None:
            return self.loc.iloc_idx[self.loc.iloc_ptr[self.k]:self.loc.iloc_ptr[self.k + 1]]
        return self.nd.int_idx[self.nd.int_ptr[self.k]:self.nd.int_ptr[self.k + 1]]

    @property
    def bnd(self):
        if self.loc is not None:
            return self.loc.bloc_idx[self.loc.bloc_ptr[self.k]:self.loc.bloc_ptr[self.k + 1]]
        return self.nd.bnd_idx[self.nd.bnd_ptr[self.k]:self.nd.bnd_ptr[self.k + 1]]

    @property
    def left(self):
        c = int(self.nd.left[self.k])
        return None if c < 0 else NDNodeView(self.nd, c, self.loc)

    @property
    def right(self):
        c = int(self.nd.right[self.k])
        return None if c < 0 else NDNodeView(self.nd, c, self.loc)


def isleaf(x) -> bool:
    return x.left is None and x.right is None


def isbranch(x) -> bool:
    return x.left is not None and x.right is not None


def depth(nd) -> int:
    """``HssMatrices.depth`` as used by factorization.jl:8 (leaf = 1)."""
    nd = nd.nd if isinstance(nd, NDNodeView) else nd
    nd._need_analyzed()
    return int(nd.depth)


def parse_elimtree(fathers, lsons, rsons, ninter, inter, nbound, bound) -> NestedDissection:
    """nesteddissection.jl:105-148.  Accepts the padded matrices of the ``.mat`` schema."""
    et = ElimTree.from_padded({"fathers": fathers, "lsons": lsons, "rsons": rsons, "ninter": ninter, "inter": inter,
                               "nbound": nbound, "bound": bound})
    return from_elimtree(et)


def from_elimtree(et: ElimTree) -> NestedDissection:
    """Same as ``parse_elimtree`` for the ragged container ``problems.ElimTree`` (no padding)."""
    if int(np.count_nonzero(np.asarray(et.fathers) == -1)) != 1:
        raise _lib.ArgumentError(_lib.HS_EARG, "found either less than or more than one root.")  # :111
    return NestedDissection(et)


def symfact(nd: NestedDissection) -> Tuple[NestedDissection, NDLoc]:
    """``symfact!`` nesteddissection.jl:29-69 (runs ``hs_symfact`` in C++)."""
    et = nd.elim
    if et is None:
        raise ValueError("symfact: NestedDissection holds no elimination tree")
    arrs = [_lib.as_i64(a) for a in (et.fathers, et.lsons, et.rsons, et.inter_ptr, et.inter_idx, et.bound_ptr, et.bound_idx)]
    cet = _lib.hs_elimtree(et.nnodes, *[_lib.ptr(a) for a in arrs], 1)
    h = C.c_void_p()
    _lib.check(_lib.lib.hs_symfact(C.byref(cet), 0, C.byref(h)))
    try:
        t = _lib.hs_tree()
        _lib.check(_lib.lib.hs_symbolic_tree(h, C.byref(t)))
        nn = int(t.nnodes)
        cp = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.int64)
        out = NestedDissection(et)
        out.nnodes = nn
        out.left, out.right = cp(t.left, nn), cp(t.right, nn)
        out.int_ptr = cp(t.int_ptr, nn + 1)
        out.bnd_ptr = cp(t.bnd_ptr, nn + 1)
        iloc_ptr, bloc_ptr = cp(t.iloc_ptr, nn + 1), cp(t.bloc_ptr, nn + 1)
        out.int_idx = cp(t.int_idx, int(out.int_ptr[-1]))
        out.bnd_idx = cp(t.bnd_idx, int(out.bnd_ptr[-1]))
        iloc_idx, bloc_idx = cp(t.iloc_idx, int(iloc_ptr[-1])), cp(t.bloc_idx, int(bloc_ptr[-1]))
        pp, pn = _lib.i64p(), C.c_int64()
        _lib.check(_lib.lib.hs_symbolic_perm(h, C.byref(pp), C.byref(pn)))
        out._perm = cp(pp, int(pn.value))
        d = C.c_int64()
        _lib.check(_lib.lib.hs_symbolic_depth(h, C.byref(d)))
        out.depth = int(d.value)
        out.root = nn - 1
        out.analyzed = True
    finally:
        _lib.lib.hs_symbolic_free(h)
    # positions are reported with the tree's index base (1)
    return out, NDLoc(out, iloc_ptr, iloc_idx, bloc_ptr, bloc_idx)


def postorder(nd: NestedDissection) -> np.ndarray:
    """nesteddissection.jl:73-79 — every node's ``int`` in post-order, then the root's ``bnd`` (1-based)."""
    nd._need_analyzed()
    return np.concatenate([nd.int_idx, nd.node().bnd]).astype(np.int64)


def invperm(p: np.ndarray) -> np.ndarray:
    p = np.asarray(p, dtype=np.int64)
    ip = np.empty_like(p)
    ip[p - 1] = np.arange(1, len(p) + 1, dtype=np.int64)
    return ip


def permuted(nd: NestedDissection, perm: np.ndarray) -> NestedDissection:
    """``permuted!`` nesteddissection.jl:82-88: ``nd.int = perm[nd.int]`` on every node (in place)."""
    nd._need_analyzed()
    perm = np.asarray(perm, dtype=np.int64)
    nd.int_idx = perm[nd.int_idx - 1]
    nd.bnd_idx = perm[nd.bnd_idx - 1]
    return nd


def permute(A, p: np.ndarray, q: np.ndarray):
    """``SparseArrays.permute(A, p, q)`` = ``A[p, q]`` with 1-based permutations (test/rungmres.jl:18)."""
    A = sp.csc_matrix(A)
    return sp.csc_matrix(A[np.asarray(p) - 1][:, np.asarray(q) - 1])


def contigious(idx: np.ndarray):
    """nesteddissection.jl:91 — a ``range`` when the index vector is one."""
    idx = np.asarray(idx)
    if len(idx) and np.array_equal(idx, np.arange(idx[0], idx[-1] + 1)):
        return range(int(idx[0]), int(idx[-1]) + 1)
    return idx


def getinterior(nd: NestedDissection):
    """nesteddissection.jl:100 (the definition that wins): ``1:nd.int[end]``, valid after the post-order permutation."""
    return range(1, int(nd.node().int[-1]) + 1)


def getboundary(nd: NestedDissection):
    """nesteddissection.jl:101."""
    return nd.node().bnd
