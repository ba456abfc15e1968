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
