"""Residual of the direct solve on 2D complex Helmholtz grids of growing size:  python tools/helm_check.py 256 512 1024 2048"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import _pkg
hs = _pkg.load()
for g in [int(a) for a in sys.argv[1:]] or [256, 512, 1024]:
    for kind in ("helmholtz", "poisson"):
        prob = hs.grid_problem((g, g), kind)
        Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
        Ap.sort_indices()
        F = hs.factor(Ap, nd, nd_loc, swlevel=0)
        x = hs.ldiv(F, prob.b)
        r1 = np.linalg.norm(Ap @ x - prob.b) / np.linalg.norm(prob.b)
        xg, h = hs.gmres(Ap, prob.b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True, A_is_factored=True)
        r2 = np.linalg.norm(Ap @ xg - prob.b) / np.linalg.norm(prob.b)
        B = np.stack([prob.b, prob.b[::-1]], axis=1)
        X = hs.ldiv(F, B)
        r3 = np.linalg.norm(Ap @ X - B) / np.linalg.norm(B)
        print(f"{kind} {g}: ldiv resid {r1:.2e}, gmres iters {h.iters} resid {r2:.2e}, 2-rhs resid {r3:.2e}, max_ni {F.stats()['max_ni']}", flush=True)
        del F
