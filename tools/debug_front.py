import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scipy.linalg as sla
import _pkg
hs = _pkg.load()
from test_gpu_parity import _single_front_problem, rel
ni, nb, cx = int(sys.argv[1]), int(sys.argv[2]), bool(int(sys.argv[3]))
prob, A = _single_front_problem(hs, ni, nb, cx)
Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
F = hs.factor(Ap, nd, nd_loc, swlevel=0)
Aii = A[:ni, :ni]
print("D", rel(F.D, Aii))
piv = F.piv
lu, p = sla.lu_factor(Aii)
print("piv equal", np.array_equal(piv, p), "first diff", np.nonzero(piv != p)[0][:10], piv[:8], p[:8])
if nb:
    R = np.linalg.solve(Aii, A[:ni, ni:]); L = np.linalg.solve(Aii.T, A[ni:, :ni].T).T
    S = A[ni:, ni:] - A[ni:, :ni] @ R
    for nm, got, ref in (("R", F.R, R), ("L", F.L, L), ("S", F.S, S)):
        e = np.abs(got - ref)
        bad = np.argwhere(e > 1e-8 * np.abs(ref).max())
        print(nm, rel(got, ref), "bad entries", len(bad), "rows", np.unique(bad[:, 0])[:10], "cols", np.unique(bad[:, 1])[:10] if len(bad) else "")
x = hs.ldiv(F, prob.b)
print("resid", rel(A @ x, prob.b))
