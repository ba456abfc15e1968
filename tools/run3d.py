"""3D grid problems (7-point Poisson / complex Helmholtz): factor + GMRES, dense and with compressed upper fronts.
    python tools/run3d.py [side] [kind] [swsize] [tol]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import _pkg  # noqa: E402

hs = _pkg.load()
side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = sys.argv[2] if len(sys.argv) > 2 else "helmholtz"
swsize = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-4
t0 = time.perf_counter()
prob = hs.grid_problem((side, side, side), kind)
Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
print(f"{side}^3 {kind}: N = {Ap.shape[0]}, nodes {nd.nnodes}, setup {time.perf_counter() - t0:.1f} s", flush=True)
for label, opts in (("dense", dict(swlevel=0)), ("compressed", dict(swlevel=-2, swsize=swsize, atol=tol, rtol=tol))):
    F = hs.factor(Ap, nd, nd_loc, **opts)
    F.refactor(Ap)
    st = F.stats()
    x = hs.ldiv(F, prob.b)
    st2 = F.stats()
    res = np.linalg.norm(Ap @ x - prob.b) / np.linalg.norm(prob.b)
    xs, ch = hs.gmres(Ap, prob.b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True)
    print(f"  {label:10s}: factor {st['ms_factor_total']:.1f} ms = {st['factor_flops'] / st['ms_factor_total'] / 1e9:.2f} TFLOP/s (dense flop count "
          f"{st['factor_flops'] / 1e12:.2f} TF), max front ni {st['max_ni']} nb {st['max_nb']}, fronts {st['front_bytes'] / 1e9:.2f} GB + {st['lowrank_bytes'] / 1e9:.2f} GB, "
          f"apply {st2['ms_solve_total']:.2f} ms resid {res:.1e}, maxrank {hs.maxrank(F)}, gmres {ch.iters} iters conv {ch.isconverged} "
          f"final {np.linalg.norm(Ap @ xs - prob.b) / np.linalg.norm(prob.b):.1e}", flush=True)
    del F
