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
    # --- extensions of this library (not fields of the reference struct) ---
    hss: bool = True        # compressed nodes store S as an HSS matrix (`randcompress_adaptive`, factorization.jl:102-110) like
                            # the reference; False: S evaluated exactly and kept dense (HSS tolerance -> 0, the round-1 form)
    sketches: object = None # (Omega, Psi): host-supplied Gaussian test matrices (rows >= largest compressed boundary, cols >= the
                            # largest sample count); a node with m boundary rows and k samples uses Omega[:m, :k], Psi[:m, :k].
                            # None: drawn on the device from a counter-based generator seeded with `sketch_seed`
    sketch_seed: int = 123  # test/rungmres.jl:7 seeds Julia's global RNG with 123

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
    if opts.verbose:
        # c_tol is validated and then ignored by the reference itself (factorization.jl:97: tolerance hard-coded 0.5x)
        print("hsolve: option c_tol has no effect (as in the reference, factorization.jl:97-100)")
        if not opts.hss:
            print("hsolve: hss=False - Schur complements stay dense: leafsize, kest, stepsize have no effect")


def to_c(opts: SolverOptions, subtree: bool = False, dtype=None):
    """``(hs_opts, keepalive)``: the C struct and the host arrays it points to (sketch matrices in the factorization's
    dtype, column-major)."""
    import ctypes as C

    import numpy as np
    keep = []
    om = ps = None
    rows = cols = 0
    if opts.sketches is not None:
        Om, Ps = opts.sketches
        dt = np.float64 if dtype is None else dtype
        Om = np.asfortranarray(Om, dtype=dt)
        Ps = np.asfortranarray(Ps, dtype=dt)
        if Om.ndim != 2 or Om.shape != Ps.shape:
            raise _lib.DimensionMismatch(_lib.HS_EDIM, "sketches must be two matrices of the same shape")
        keep += [Om, Ps]
        om, ps = Om.ctypes.data_as(C.c_void_p), Ps.ctypes.data_as(C.c_void_p)
        rows, cols = Om.shape
    o = _lib.hs_opts(int(opts.swlevel), int(opts.swsize), float(opts.atol), float(opts.rtol), float(opts.c_tol),
                     int(opts.leafsize), int(opts.kest), int(opts.stepsize), int(bool(opts.verbose)),
                     int(bool(subtree)), int(bool(opts.hss)), 0, om, ps, int(rows), int(cols), int(opts.sketch_seed))
    return o, keep
