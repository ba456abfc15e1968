"""Generates tests/golden/*.npz from the oracle (the reference itself cannot run here: no Julia).
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkg  # noqa: E402
import hs_oracle as orc  # noqa: E402

hs = _pkg.load()
prob = hs.grid_problem((17, 17), "poisson", nmax=40)
Ao, nd, nd_loc, perm = orc.prepare(prob.A, prob.elim_tree)
F = orc.factor(Ao, nd, nd_loc)
x = orc.ldiv(F, prob.b)
nodes = orc.nodes_postorder(F)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "poisson2d_17x17.npz"), perm=perm, x=x, root_S=F.S,
                    leaf0_L=nodes[0].L, leaf0_D=nodes[0].D, leaf0_S=nodes[0].S)
print("wrote poisson2d_17x17.npz", x.shape)

# compressed branch (oracle/hs_oracle_hss.py): ranks per node, the preconditioner application and the GMRES history
import hs_oracle_hss as oh  # noqa: E402
import scipy.sparse as sp  # noqa: E402

prob = hs.grid_problem((33, 33), "helmholtz", nmax=40)
rng = np.random.default_rng(7)
A = sp.csr_matrix(prob.A).copy()
A.data = A.data * (1.0 + 0.3 * rng.random(A.nnz))          # no tied column norms: the pivot order is unambiguous
A = sp.csc_matrix(A)
Ao, nd, nd_loc, perm = orc.prepare(A, prob.elim_tree)
opts = dict(swlevel=-1, swsize=8, atol=1e-4, rtol=1e-4)
F = orc.factor(Ao, nd, nd_loc, **opts)
x = orc.ldiv(F, prob.b)
_, res, conv = orc.gmres(Ao, prob.b, Pr=lambda v: orc.ldiv(F, v), reltol=1e-9, restart=30, maxiter=30)
ranks = np.asarray(oh.node_ranks(F))
kc = int(np.nonzero(ranks.max(axis=1) > 0)[0][-1])          # the highest compressed node (post-order id)
nodec = orc.nodes_postorder(F)[kc]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "helmholtz2d_33x33_compressed.npz"), scale=A.data, perm=perm,
                    ranks=ranks, x=x, resnorm=np.asarray(res), converged=conv, node=kc, node_L=nodec.L_dense(),
                    node_R=nodec.R_dense(), **{k: np.asarray(v) for k, v in opts.items()})
print("wrote helmholtz2d_33x33_compressed.npz  maxrank", orc.maxrank(F), "gmres", len(res))
