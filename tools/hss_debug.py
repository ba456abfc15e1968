"""Small HSS factorizations for compute-sanitizer / debugging:  python tools/hss_debug.py [kind] [n] [swlevel]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, scipy.sparse as sp
import _pkg
hs = _pkg.load()
kind = sys.argv[1] if len(sys.argv) > 1 else "helmholtz"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 33
sw = int(sys.argv[3]) if len(sys.argv) > 3 else -2
prob = hs.grid_problem((n, n), kind, nmax=40)
rng = np.random.default_rng(0)
A = sp.csr_matrix(prob.A).copy(); A.data = A.data * (1.0 + 0.3 * rng.random(A.nnz)); A = sp.csc_matrix(A)
Ap, nd, nd_loc, perm = hs.prepare(A, prob.elim_tree)
for sk in (True, False):
    nb = int(prob.elim_tree.nbound().max())
    sketches = (rng.standard_normal((nb, 120)), rng.standard_normal((nb, 120))) if sk else None
    F = hs.factor(Ap, nd, nd_loc, swlevel=sw, swsize=12, atol=1e-5, rtol=1e-5, leafsize=8, hss=True, sketches=sketches)
    x = hs.ldiv(F, prob.b)
    st = F.stats()
    print(kind, n, "sketches" if sk else "generated", "maxrank", hs.maxrank(F), "hss_maxrank", st["hss_maxrank"], "rounds", st["hss_rounds"], "nodes", st["hss_nodes"],
          "resid", np.linalg.norm(Ap @ x - prob.b) / np.linalg.norm(prob.b), flush=True)
    for k in range(nd.nnodes):
        g = F.node(k).hss()
        if g is not None:
            S = F.node(k).S
            print("  node", k, "hss nodes", len(g), "hssrank", F.node(k).hssrank(), "ranks", F.node(k).ranks(), "S", S.shape, flush=True)
print("ok")
