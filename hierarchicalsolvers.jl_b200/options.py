"""``SolverOptions`` — reference src/HierarchicalSolvers.jl:30-79 (same fields, defaults and checks)."""
from __future__ import annotations

from dataclasses import dataclass, fields, replace

from . import _lib


@dataclass
class SolverOptions:
    swlevel: int = 5        # switching level at which to start compression
    swsize: int = 1         # minimum boundary size for compression
    atol: float = 1e-6      # absolute compression tolerance
    rtol: float = 1e-6      # relative compression tolerance
    c_tol: float = 0.5      # low-rank vs HSS tolerance factor (validated, unused — as in the reference)
    leafsize: int = 32      # HSS leaf size
    kest: int = -1          # rank estimate
    stepsize: int = 10      # rank increment of the adaptive sampler
    verbose: bool = False

    def copy(self, **kw) -> "SolverOptions":
        """``copy(opts; kw...)`` HierarchicalSolvers.jl:62-71."""
        names = {f.name for f in fields(self)}
        for k in kw:
            if k not in names:
                raise TypeError(f"type SolverOptions has no field {k}")
        return replace(self, **kw)


def chkopts(opts: SolverOptions) -> None:
    """``chkopts!`` HierarchicalSolvers.jl:73-79."""
    def bad(name):
        raise _lib.ArgumentError(_lib.HS_EARG, name)
    if not opts.swsize >= 1: bad("swsize")
    if not opts.atol >= 0.0: bad("atol")
    if not opts.rtol >= 0.0: bad("rtol")
    if not (0.0 < opts.c_tol <= 1.0): bad("c_tol")
    if not opts.leafsize >= 1: bad("leafsize")


def to_c(opts: SolverOptions, subtree: bool = False) -> _lib.hs_opts:
    return _lib.hs_opts(int(opts.swlevel), int(opts.swsize), float(opts.atol), float(opts.rtol), float(opts.c_tol),
                        int(opts.leafsize), int(opts.kest), int(opts.stepsize), int(bool(opts.verbose)),
                        int(bool(subtree)))
