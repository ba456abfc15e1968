"""hierarchicalsolvers.jl_b200 — B200-native multifrontal nested-dissection factor + tree solve behind the
API of bonevbs/HierarchicalSolvers.jl (export list: reference src/HierarchicalSolvers.jl:20-28).

Host mirror (Python, because Julia is absent from this image) over the C ABI of ``libhsolve_cuda``
(include/hsolve_cuda.h).  All numerics run in hand-written sm_100a CUDA kernels; there is no CPU fallback.
"""
from . import _lib
from .options import SolverOptions, chkopts
from .problems import (ElimTree, Problem, grid_problem, grid_elimtree, grid_operator, read_problem, write_problem,
                       nested_dissection)
from .nesteddissection import (NestedDissection, parse_elimtree, from_elimtree, symfact, postorder, permuted, invperm,
                               permute, contigious, getinterior, getboundary, depth)
from .factornode import FactorNode, ldiv, maxrank, isleaf, isbranch, eltype
from .factorization import factor
from .gmres import gmres, ConvergenceHistory
from ._lib import SingularException, DimensionMismatch, ArgumentError, HSolveError

__all__ = [
    "SolverOptions", "chkopts",
    "NestedDissection", "parse_elimtree", "from_elimtree", "postorder", "getinterior", "getboundary", "symfact",
    "permuted", "invperm", "permute", "contigious", "depth",
    "FactorNode", "ldiv", "maxrank", "isleaf", "isbranch", "eltype",
    "factor", "gmres", "ConvergenceHistory",
    "ElimTree", "Problem", "grid_problem", "grid_elimtree", "grid_operator", "read_problem", "write_problem",
    "SingularException", "DimensionMismatch", "ArgumentError", "HSolveError", "nested_dissection",
]


def prepare(A, elim_tree):
    """The driver's preamble (test/rungmres.jl:15-19) in one call: parse, ``symfact!``, post-order permutation of
    ``A`` and of the tree.  Returns ``(A_permuted, nd, nd_loc, perm)``."""
    nd = from_elimtree(elim_tree)
    nd, nd_loc = symfact(nd)
    perm = postorder(nd)
    Ap = permute(A, perm, perm)
    nd = permuted(nd, invperm(perm))
    return Ap, nd, nd_loc, perm
