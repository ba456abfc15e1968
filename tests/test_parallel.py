"""Subtree-per-GPU layer (hierarchicalsolvers.jl_b200/parallel.py): tree cut invariants on CPU, the N>1 host logic
with world_size-2/4 `gloo` processes on CPU (numerics by an oracle-backed engine), and the real thing on ≥2 GPUs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_dist_worker.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return str(p)


def _run_ranks(backend, world, grid, kind, mode="tree", timeout=600):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, WORKER, backend, str(r), str(world), port, str(grid), kind, mode],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append((p.returncode, out))
    for rc, out in outs:
        assert rc == 0, out[-3000:]
        assert "RESULT" in out
    return outs


@pytest.mark.parametrize("nparts", [1, 2, 3, 4, 8])
def test_partition_tree_invariants(hs, nparts):
    from hsolve_b200.parallel import partition_tree
    prob = hs.grid_problem((65, 65), "poisson", nmax=40)
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    part = partition_tree(nd, nd_loc, nparts)
    assert len(part.cut) == nparts == len(part.sub_nd)
    # subtrees + top nodes partition the tree; the cut nodes appear in both (root of a subtree, leaf of the top)
    nsub = sum(s.nnodes for s in part.sub_nd)
    assert nsub + part.top_nd.nnodes - nparts == nd.nnodes
    # every DOF is eliminated exactly once: inside one subtree or above the cut
    top_int = part.top_nd.int_idx
    allint = np.concatenate(part.int_idx + [top_int])
    assert np.array_equal(np.sort(allint), np.arange(1, Ap.shape[0] + 1))
    # boundaries of different subtree roots are disjoint (the exchange is a concatenation)
    allb = np.concatenate(part.bnd_idx)
    assert len(np.unique(allb)) == len(allb)
    for r, leaf in enumerate(part.top_leaf):
        v = part.top_nd.node(leaf)
        assert len(v.int) == 0 and np.array_equal(v.bnd, part.bnd_idx[r])
        assert part.top_nd.left[leaf] == -1 and part.top_nd.right[leaf] == -1
        assert np.array_equal(part.sub_nd[r].node().bnd, part.bnd_idx[r])
    if nparts in (2, 4, 8):  # balanced tree: the cut is one whole level
        lv = np.zeros(nd.nnodes, int)
        for k in range(nd.nnodes - 1, -1, -1):
            for c in (nd.left[k], nd.right[k]):
                if c >= 0:
                    lv[c] = lv[k] + 1
        assert len(set(lv[part.cut])) == 1


@pytest.mark.parametrize("world,mode", [(2, "tree"), (4, "tree"), (3, "tree"), (4, "replicated"), (8, "tree")])
def test_gloo_distributed_matches_direct_solve(world, mode):
    outs = _run_ranks("gloo", world, 65 if world > 4 else 33, "poisson", mode)
    assert all("same=True" in o for _, o in outs)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["poisson", "helmholtz"])
def test_nccl_two_gpus(hs, kind):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    outs = _run_ranks("nccl", 2, 257, kind)
    assert all("same=True" in o for _, o in outs)
