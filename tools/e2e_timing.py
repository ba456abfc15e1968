"""Where the end-to-end time of hs.factor / hs.gmres goes (host side).  python tools/e2e_timing.py [grid]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import _pkg
hs = _pkg.load()
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
prob = hs.grid_problem((grid, grid), "poisson")
Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
Ap.sort_indices()
b = prob.b
for rep in range(3):
    t0 = time.perf_counter()
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    t1 = time.perf_counter()
    x, h = hs.gmres(Ap, b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True, A_is_factored=True)
    t2 = time.perf_counter()
    st = F.stats()
    print(f"rep {rep}: factor call {1e3*(t1-t0):.1f} ms (analyze {st['ms_analyze']:.1f}, h2d {st['ms_h2d']:.1f}, numeric {st['ms_factor_total']:.1f}), "
          f"gmres call {1e3*(t2-t1):.1f} ms, iters {h.iters}", flush=True)
    t3 = time.perf_counter()
    del F
    print(f"        free {1e3*(time.perf_counter()-t3):.1f} ms", flush=True)
