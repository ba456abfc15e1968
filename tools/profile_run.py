"""One factorization + one preconditioner application, for ncu launch lists / captures.
    python tools/profile_run.py [grid | grid^3] [kind] [repeats]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import _pkg  # noqa: E402

hs = _pkg.load()
garg = sys.argv[1] if len(sys.argv) > 1 else "1024"
kind = sys.argv[2] if len(sys.argv) > 2 else "poisson"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cube = garg.endswith("^3")
grid = int(garg[:-2]) if cube else int(garg)
prob = hs.grid_problem((grid,) * (3 if cube else 2), kind)
Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
for _ in range(reps - 1):   # warm repeats: only the last factorization is reported in full
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    print(f"  warm-up factor {F.stats()['ms_factor_total']:.2f} ms")
    del F
t0 = time.perf_counter()
F = hs.factor(Ap, nd, nd_loc, swlevel=0)
t1 = time.perf_counter()
x = hs.ldiv(F, prob.b)
t2 = time.perf_counter()
st = F.stats()
print(f"grid {garg} {kind}: factor {st['ms_factor_total']:.2f} ms (call {1e3 * (t1 - t0):.0f} ms), solve {st['ms_solve_total']:.2f} ms "
      f"(call {1e3 * (t2 - t1):.0f} ms), resid {np.linalg.norm(Ap @ x - prob.b) / np.linalg.norm(prob.b):.2e}, "
      f"{st['factor_flops'] / st['ms_factor_total'] / 1e9:.2f} TFLOP/s")
if os.environ.get("HS_PROFILE"):
    print("  phases(ms):", {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")}, "gemm TF/s",
          round(st["gemm_flops"] / max(st["ms_gemm"], 1e-9) / 1e9, 2), "launches", st["launches_factor"], st["launches_solve"])
