"""Compressed factorization with HSS Schur complements at scale:  python tools/hss_run.py [grid] [kind] [swsize] [tol] [leafsize] [hss]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import _pkg
hs = _pkg.load()
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kind = sys.argv[2] if len(sys.argv) > 2 else "poisson"
swsize = int(sys.argv[3]) if len(sys.argv) > 3 else 128
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-5
leaf = int(sys.argv[5]) if len(sys.argv) > 5 else 32
hss = int(sys.argv[6]) if len(sys.argv) > 6 else 1
dim3 = kind.endswith("3d")
kind = kind.replace("3d", "")
prob = hs.grid_problem((grid, grid, grid) if dim3 else (grid, grid), kind, nmax=100)
Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
Ap.sort_indices()
b = prob.b
t0 = time.perf_counter()
F = hs.factor(Ap, nd, nd_loc, swlevel=-2, swsize=swsize, atol=tol, rtol=tol, leafsize=leaf, hss=bool(hss))
t1 = time.perf_counter()
st = F.stats()
print(f"grid {grid} {kind} swsize {swsize} tol {tol} leafsize {leaf} hss {hss}: first factor call {t1 - t0:.3f} s, numeric {st['ms_factor_total']:.1f} ms, "
      f"maxrank {hs.maxrank(F)}, hss_maxrank {st['hss_maxrank']}, rounds {st['hss_rounds']}, hss nodes {st['hss_nodes']}, front GB {st['front_bytes'] / 1e9:.2f}, "
      f"lowrank GB {st['lowrank_bytes'] / 1e9:.2f}, hss GB {st['hss_bytes'] / 1e9:.3f}, launches {st['launches_factor']}", flush=True)
F.refactor(Ap)
st = F.stats()
print(f"  refactor: numeric {st['ms_factor_total']:.1f} ms, launches {st['launches_factor']}", flush=True)
x, h = hs.gmres(Ap, b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True, A_is_factored=True)
print(f"  gmres iters {h.iters} converged {h.isconverged} residual {np.linalg.norm(Ap @ x - b) / np.linalg.norm(b):.2e}, apply {F.stats()['ms_solve_total']:.2f} ms", flush=True)
if os.environ.get("HS_PROFILE"):
    print("  phases:", {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")})
print("ok")
