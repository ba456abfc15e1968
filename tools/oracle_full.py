import sys, time, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
g = int(sys.argv[1]); nt = int(sys.argv[2])
from threadpoolctl import threadpool_limits
import importlib.util
# problem generation without the product .so? use product package python only
import _pkg
hs = _pkg.load()
import hs_oracle as orc
t0=time.perf_counter()
prob = hs.grid_problem((g, g), "poisson", nmax=100)
Ao, nd, nd_loc, _ = orc.prepare(prob.A, prob.elim_tree)
print("setup", time.perf_counter()-t0, flush=True)
with threadpool_limits(limits=nt):
    t0=time.perf_counter()
    F = orc.factor(Ao, nd, nd_loc)
    t1=time.perf_counter()
    print("factor", t1-t0, flush=True)
    x,res,conv = orc.gmres(Ao, prob.b, Pr=lambda v: orc.ldiv(F, v), reltol=1e-9, restart=30, maxiter=30)
    t2=time.perf_counter()
print(json.dumps({"grid":g,"threads":nt,"factor_s":t1-t0,"solve_s":t2-t1,"iters":len(res)}), flush=True)
import psutil; print("rss GB", psutil.Process().memory_info().rss/1e9)
