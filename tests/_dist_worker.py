"""Worker for the multi-rank tests: `python tests/_dist_worker.py <backend> <rank> <world> <port> <grid> <kind>`.
backend = gloo  → CPU tensors, numerics by an engine built on the oracle (host logic of parallel.py under test)
backend = nccl  → one GPU per rank, numerics by the CUDA library."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def to_oracle_tree(orc, nd, loc, k=None):
    k = nd.root if k is None else k
    l, r = int(nd.left[k]), int(nd.right[k])
    nl = to_oracle_tree(orc, nd, loc, l) if l >= 0 else (None, None)
    nr = to_oracle_tree(orc, nd, loc, r) if r >= 0 else (None, None)
    a = orc.NDNode(nd.node(k).int, nd.node(k).bnd, nl[0], nr[0])
    b = orc.NDNode(loc.node(k).int, loc.node(k).bnd, nl[1], nr[1])
    return a, b


class OracleEngine:
    """Same interface as parallel.CudaEngine, numerics by oracle/hs_oracle.py on CPU tensors."""

    def __init__(self):
        import torch
        import hs_oracle
        self.torch, self.orc = torch, hs_oracle
        self.t_dtype = torch.float64

    def factor_subtree(self, A, nd, loc, opts):
        import scipy.sparse as sp
        a, b = to_oracle_tree(self.orc, nd, loc)
        return {"F": self.orc._factor(sp.csr_matrix(A), a, b, 1), "top": False, "A": sp.csr_matrix(A), "a": a, "b": b}

    def export_schur(self, h, node, nb, pad):
        F = h["F"]
        perm = np.concatenate([F.int_loc, F.bnd_loc]) - 1
        S = np.zeros((nb, nb))
        S[np.ix_(perm, perm)] = F.S                      # undo S[perm,perm] (factorization.jl:41,74)
        buf = self.torch.zeros((pad, pad), dtype=self.t_dtype)
        buf[:nb, :nb] = self.torch.from_numpy(S.T.copy())  # column-major block
        return buf

    def analyze_top(self, A, nd, loc, opts, subtree=False):
        import scipy.sparse as sp
        return {"A": sp.csr_matrix(A), "nd": nd, "loc": loc, "ext": {}, "top": True, "subtree": subtree}

    def import_schur(self, h, node, buf, pad):
        h["ext"][node] = buf

    def numeric(self, h):
        if not h["top"]:
            h["F"] = self.orc._factor(h["A"], h["a"], h["b"], 1)
            return
        orc, nd, loc = self.orc, h["nd"], h["loc"]

        def build(k):
            nv, lv = nd.node(k), loc.node(k)
            if k in h["ext"]:
                nb = len(nv.bnd)
                S = h["ext"][k][:nb, :nb].numpy().T.copy()
                perm = np.concatenate([lv.int, lv.bnd]) - 1
                e = np.zeros(0, dtype=np.int64)
                node_nd = orc.NDNode(e, nv.bnd)
                return orc.FactorNode(np.zeros((0, 0)), S[np.ix_(perm, perm)], np.zeros((nb, 0)), np.zeros((0, nb)), e, nv.bnd,
                                      lv.int, lv.bnd), node_nd, orc.NDNode(lv.int, lv.bnd)
            l, r = int(nd.left[k]), int(nd.right[k])
            if l < 0:
                a, b = orc.NDNode(nv.int, nv.bnd), orc.NDNode(lv.int, lv.bnd)
                return orc._factor_leaf(h["A"], a, b), a, b
            Fl, al, bl = build(l)
            Fr, ar, br = build(r)
            a, b = orc.NDNode(nv.int, nv.bnd, al, ar), orc.NDNode(lv.int, lv.bnd, bl, br)
            return orc._factor_branch(h["A"], Fl, Fr, a, b), a, b

        h["F"] = build(nd.root)[0]

    def to_device(self, b):
        return self.torch.from_numpy(np.array(b, dtype=np.float64))

    def to_host(self, x):
        return x.numpy().copy()

    def sweep(self, h, x, which):
        orc, F = self.orc, h["F"]
        v = x.numpy().reshape(-1, 1)
        if which & 1:
            orc._lsolve(F, v)
            orc._dsolve(F, v)
        if which == 3 and len(F.bnd) and not h.get("subtree", False) and h["top"]:
            v[F.bnd - 1] = orc._solve(F.S, v[F.bnd - 1])
        if which & 2:
            orc._rsolve_tree(F, v)

    def index(self, idx):
        return self.torch.from_numpy(np.asarray(idx, dtype=np.int64) - 1)

    def zeros(self, n):
        return self.torch.zeros(n, dtype=self.t_dtype)


def main():
    backend, rank, world, port, grid, kind = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), sys.argv[6]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import _pkg
    hs = _pkg.load()
    from hsolve_b200.parallel import CudaEngine, DistributedFactor
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        eng = CudaEngine(rank)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        eng = OracleEngine()
    prob = hs.grid_problem((grid, grid), kind, nmax=40)
    Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
    mode = sys.argv[7] if len(sys.argv) > 7 else "tree"
    DF = DistributedFactor(Ap, nd, nd_loc, engine=eng, swlevel=0, top=mode)
    x = DF.ldiv(prob.b)
    DF.refactor()
    x2 = DF.ldiv(prob.b)
    assert np.array_equal(x, x2), "refactor changed the solution"
    res = np.linalg.norm(Ap @ x - prob.b) / np.linalg.norm(prob.b)
    # every rank must hold the same, correct solution
    import scipy.sparse.linalg as spla
    xr = spla.splu(Ap.tocsc()).solve(prob.b)
    err = np.linalg.norm(x - xr) / np.linalg.norm(xr)
    xs = [None] * world
    dist.all_gather_object(xs, x)
    same = all(np.array_equal(xs[0], v) for v in xs)
    print(f"RESULT rank={rank} res={res:.3e} err={err:.3e} same={same} cut={DF.part.cut} schur_bytes={DF.schur_bytes}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if (res < 1e-10 and err < 1e-10 and same) else 3)


if __name__ == "__main__":
    main()
