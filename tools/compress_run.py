"""Compressed vs uncompressed factorization of one grid problem: factor / solve time, ranks, GMRES iterations.
    python tools/compress_run.py [grid] [kind] [swlevel] [swsize] [tol]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import _pkg  # noqa: E402

hs = _pkg.load()
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kind = sys.argv[2] if len(sys.argv) > 2 else "poisson"
swlevel = int(sys.argv[3]) if len(sys.argv) > 3 else -2
swsize = int(sys.argv[4]) if len(sys.argv) > 4 else 480
tol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-2
prob = hs.grid_problem((grid, grid), kind)
Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
if os.environ.get("HS_PROFILE"):
    import ctypes as C
    hs._lib.lib.hs_set_profile(hs._lib.default_context(0), 1)
only = os.environ.get("HS_ONLY")
for label, opts in (("dense", dict(swlevel=0)), ("compressed", dict(swlevel=swlevel, swsize=swsize, atol=tol, rtol=tol))):
    if only and label != only:
        continue
    for rep in range(1 if only else 2):
        t0 = time.perf_counter()
        F = hs.factor(Ap, nd, nd_loc, **opts) if rep == 0 else F.refactor(Ap) or F
        t1 = time.perf_counter()
    x = hs.ldiv(F, prob.b)
    x = hs.ldiv(F, prob.b)
    st = F.stats()
    res = np.linalg.norm(Ap @ x - prob.b) / np.linalg.norm(prob.b)
    t2 = time.perf_counter()
    xs, ch = hs.gmres(Ap, prob.b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True)
    t3 = time.perf_counter()
    print(f"{label:10s} grid {grid} {kind} {opts}: factor {st['ms_factor_total']:.2f} ms (compress {st['ms_compress']:.2f}), "
          f"solve {st['ms_solve_total']:.2f} ms, maxrank {hs.maxrank(F)}, apply resid {res:.2e}, gmres iters {ch.iters} "
          f"conv {ch.isconverged} ({1e3 * (t3 - t2):.0f} ms) final resid {np.linalg.norm(Ap @ xs - prob.b) / np.linalg.norm(prob.b):.2e}, "
          f"launches {st['launches_factor']}/{st['launches_solve']}")
    if os.environ.get("HS_PROFILE"):
        print("  phases(ms):", {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")})
    del F
