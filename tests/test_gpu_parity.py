"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.
Tolerance: 1e-10 relative (Frobenius / 2-norm) — the figure BASELINE.json's north_star states for FP64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    d = np.linalg.norm(np.asarray(a) - np.asarray(b))
    s = np.linalg.norm(np.asarray(b))
    return d / s if s > 0 else d


def both(hs, orc, shape, kind, nmax=100):
    prob = hs.grid_problem(shape, kind, nmax=nmax)
    Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    Ao, ndo, ndo_loc, _ = orc.prepare(prob.A, prob.elim_tree)
    Fo = orc.factor(Ao, ndo, ndo_loc)
    return prob, Ap, F, Fo


# (65, 65) / (129, 129) with nmax = 100 are the stand-ins of BASELINE configs 1-3 (poisson2d / helmholtz2d p1 h64 / h128, whose
# .mat fixtures are stripped from the reference checkout)
CASES = [("poisson", (33, 33), 40), ("helmholtz", (33, 33), 40), ("poisson", (65, 65), 100),
         ("helmholtz", (65, 65), 100), ("poisson", (129, 129), 100), ("helmholtz", (129, 129), 100),
         ("poisson", (9, 8, 7), 60), ("helmholtz", (10, 9, 8), 60),
         ("poisson", (37, 3), 20), ("poisson", (4, 4), 100)]


@pytest.mark.parametrize("kind,shape,nmax", CASES)
def test_every_node_matches_oracle(hs, orc, kind, shape, nmax):
    """D, L, R, S of every FactorNode (factornode.jl:8-11) against the oracle's."""
    prob, Ap, F, Fo = both(hs, orc, shape, kind, nmax)
    onodes = orc.nodes_postorder(Fo)
    assert len(onodes) == F._hd.nd.nnodes
    worst = 0.0
    for k, on in enumerate(onodes):
        gn = F.node(k)
        assert np.array_equal(gn.int, on.int) and np.array_equal(gn.bnd, on.bnd)
        assert np.array_equal(gn.int_loc, on.int_loc) and np.array_equal(gn.bnd_loc, on.bnd_loc)
        for name, ref in (("D", on.D_dense()), ("L", on.L_dense()), ("R", on.R_dense()), ("S", on.S)):
            got = getattr(gn, name)
            assert got.shape == ref.shape, (k, name, got.shape, ref.shape)
            if ref.size:
                e = rel(got, ref)
                worst = max(worst, e)
                assert e < TOL, (k, name, e)
    assert hs.maxrank(F) == orc.maxrank(Fo) == 0
    assert hs.isleaf(F.node(0)) and hs.isbranch(F) or F._hd.nd.nnodes == 1


@pytest.mark.parametrize("kind,shape,nmax", CASES)
def test_ldiv_matches_oracle(hs, orc, kind, shape, nmax):
    prob, Ap, F, Fo = both(hs, orc, shape, kind, nmax)
    x = hs.ldiv(F, prob.b)
    xo = orc.ldiv(Fo, prob.b)
    assert rel(x, xo) < TOL
    assert rel(Ap @ x, prob.b) < TOL
    # 3-argument form writes into C; 2-argument form leaves B untouched (factornode.jl:62-67)
    b0 = prob.b.copy()
    C = np.empty_like(x)
    out = hs.ldiv(C, F, prob.b)
    assert out is C and np.array_equal(prob.b, b0) and np.array_equal(C, x)
    # matrices (factornode.jl:68-74)
    B = np.random.default_rng(7).standard_normal((Ap.shape[0], 3)).astype(x.dtype)
    X = hs.ldiv(F, B)
    assert rel(X, orc.ldiv(Fo, B)) < TOL


@pytest.mark.parametrize("kind,shape", [("poisson", (65, 65)), ("helmholtz", (65, 65))])
def test_gmres_iterations_match_oracle(hs, orc, kind, shape):
    """rungmres.jl:47: same iteration count and residual history as the oracle's GMRES with the oracle's factors."""
    prob, Ap, F, Fo = both(hs, orc, shape, kind)
    xo, reso, convo = orc.gmres(Ap, prob.b, Pr=lambda v: orc.ldiv(Fo, v), reltol=1e-9, restart=30, maxiter=30)
    for dev in (True, False):
        x, hist = hs.gmres(Ap, prob.b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True, device_resident=dev)
        assert hist.isconverged == convo and hist.iters == len(reso)
        assert rel(Ap @ x, prob.b) < 1e-9
        assert rel(x, xo) < 1e-8
    # no preconditioner: same (non-)convergence history as the oracle to 1e-6 relative per entry
    x0, h0 = hs.gmres(Ap, prob.b, Pr=None, reltol=1e-9, restart=30, maxiter=30, log=True)
    _, r0, c0 = orc.gmres(Ap, prob.b, Pr=None, reltol=1e-9, restart=30, maxiter=30)
    assert h0.iters == len(r0) and h0.isconverged == c0
    assert np.allclose(h0.resnorm, r0, rtol=1e-6)


@pytest.mark.parametrize("kind,shape", [("poisson", (257, 257)), ("helmholtz", (257, 257)), ("poisson", (513, 513)),
                                        ("helmholtz", (300, 200)), ("poisson", (40, 40, 40))])
def test_large_residual_property(hs, kind, shape):
    """Sizes past what the oracle finishes in seconds: the factorization is an exact direct solver
    (‖A·x − b‖/‖b‖ ≤ 1e-10) and preconditioned GMRES needs ≤ 2 steps; exercises the multi-CTA cluster panels."""
    prob = hs.grid_problem(shape, kind)
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    x = hs.ldiv(F, prob.b)
    assert rel(Ap @ x, prob.b) < TOL
    _, hist = hs.gmres(Ap, prob.b, Pr=F, log=True)
    assert hist.isconverged and hist.iters <= 2
    # linearity of the solve
    y = hs.ldiv(F, 2.0 * prob.b)
    assert rel(y, 2.0 * x) < 1e-12
    st = F.stats()
    assert st["factor_flops"] > 0 and st["launches_factor"] > 0 and st["ms_factor_total"] > 0


def test_refactor_same_pattern(hs):
    prob = hs.grid_problem((65, 65), "poisson")
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    A2 = Ap.copy()
    A2.data = A2.data * 3.0
    F.refactor(A2)
    x = hs.ldiv(F, prob.b)
    assert rel(A2 @ x, prob.b) < TOL


def _subtree_problem(hs, shape, kind):
    """The left subtree of a grid problem as a problem of its own: its root keeps a non-empty boundary, which
    exercises the root Schur solve `C[F.bnd,:] = F.S \\ C[F.bnd,:]` (factornode.jl:72)."""
    from hsolve_b200.problems import ElimTree, Problem
    full = hs.grid_problem(shape, kind, nmax=40)
    et = full.elim_tree
    root = int(np.nonzero(et.fathers == -1)[0][0])
    sub_root = int(et.lsons[root]) - 1
    keep, stack = [], [sub_root]
    while stack:
        i = stack.pop()
        keep.append(i)
        if et.lsons[i] != -1:
            stack += [int(et.lsons[i]) - 1, int(et.rsons[i]) - 1]
    keep = sorted(keep)
    newid = {old: k + 1 for k, old in enumerate(keep)}
    dofs = np.unique(np.concatenate([np.concatenate([et.inter(i), et.bound(i)]) for i in keep if et.lsons[i] == -1]))
    dmap = np.zeros(full.A.shape[0] + 1, dtype=np.int64)
    dmap[dofs] = np.arange(1, len(dofs) + 1)
    mp = lambda v: -1 if v == -1 else newid[int(v) - 1]
    fathers = np.array([-1 if i == sub_root else mp(et.fathers[i]) for i in keep], dtype=np.int64)
    lsons = np.array([mp(et.lsons[i]) for i in keep], dtype=np.int64)
    rsons = np.array([mp(et.rsons[i]) for i in keep], dtype=np.int64)
    iptr = np.concatenate([[0], np.cumsum([len(et.inter(i)) for i in keep])]).astype(np.int64)
    bptr = np.concatenate([[0], np.cumsum([len(et.bound(i)) for i in keep])]).astype(np.int64)
    iidx = dmap[np.concatenate([et.inter(i) for i in keep])]
    bidx = dmap[np.concatenate([et.bound(i) for i in keep])]
    A = full.A.tocsr()[dofs - 1][:, dofs - 1].tocsc()
    b = full.b[dofs - 1]
    return Problem(A, b, ElimTree(fathers, lsons, rsons, iptr, iidx, bptr, bidx))


@pytest.mark.parametrize("kind", ["poisson", "helmholtz"])
def test_nonempty_root_boundary(hs, orc, kind):
    prob = _subtree_problem(hs, (33, 33), kind)
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    assert len(nd.node().bnd) > 0
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    Ao, ndo, ndo_loc, _ = orc.prepare(prob.A, prob.elim_tree)
    Fo = orc.factor(Ao, ndo, ndo_loc)
    x, xo = hs.ldiv(F, prob.b), orc.ldiv(Fo, prob.b)
    assert rel(x, xo) < TOL and rel(Ap @ x, prob.b) < TOL
    assert rel(F.S, Fo.S) < TOL


def test_errors_match_reference_exceptions(hs):
    prob = hs.grid_problem((17, 17), "poisson", nmax=40)
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    with pytest.raises(hs.ArgumentError):
        hs.factor(Ap, nd, nd_loc, swlevel=0, swsize=0)          # chkopts! HierarchicalSolvers.jl:74
    Z = Ap.copy()
    Z.data[:] = 0.0
    with pytest.raises(hs.SingularException):                     # `D \ …` on a singular pivot block
        hs.factor(Z, nd, nd_loc, swlevel=0)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    with pytest.raises(hs.DimensionMismatch):
        hs.ldiv(F, np.zeros(Ap.shape[0] + 1))
    Fc = hs.factor(Ap, nd, nd_loc, swlevel=3, swsize=1)           # compressed path (tests/test_gpu_compress.py)
    assert hs.maxrank(Fc) > 0 and Fc.resolved_swlevel() == 3


def test_pivoting_is_exercised(hs, orc):
    """A matrix whose pivot blocks need row interchanges: scale rows of an indefinite Helmholtz operator so that the
    diagonal is tiny; the result must still match the oracle (LAPACK partial pivoting) to 1e-10."""
    prob = hs.grid_problem((33, 33), "helmholtz", nmax=40)
    A = prob.A.tolil()
    rng = np.random.default_rng(3)
    idx = rng.choice(A.shape[0], size=200, replace=False)
    for i in idx:
        A[i, i] = 1e-9 * (1 + 1j)
    prob.A = A.tocsc()
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    Ao, ndo, ndo_loc, _ = orc.prepare(prob.A, prob.elim_tree)
    Fo = orc.factor(Ao, ndo, ndo_loc)
    piv_moved = sum(int(np.any(F.node(k).piv != np.arange(len(F.node(k).piv)))) for k in range(nd.nnodes))
    assert piv_moved > 0
    x, xo = hs.ldiv(F, prob.b), orc.ldiv(Fo, prob.b)
    assert rel(Ap @ x, prob.b) < 1e-9 and rel(x, xo) < 1e-8


def _single_front_problem(hs, ni, nb, cx, seed=0):
    """One dense front: a tree whose only node is a leaf with ni interior and nb boundary DOFs and a dense A.
    Isolates the partial-LU kernels (panel / TRSM / DMMA update) at sizes the grid problems reach only at 1024²+."""
    import scipy.sparse as sp
    from hsolve_b200.problems import ElimTree, Problem
    rng = np.random.default_rng(seed)
    n = ni + nb
    A = rng.standard_normal((n, n))
    if cx:
        A = A + 1j * rng.standard_normal((n, n))
    A = A + np.diag(np.full(n, 0.5 * np.sqrt(n)))   # keeps the pivot block well conditioned but not pivot-free
    b = rng.standard_normal(n).astype(A.dtype)
    et = ElimTree(np.array([-1]), np.array([-1]), np.array([-1]), np.array([0, ni]), np.arange(1, ni + 1),
                  np.array([0, nb]), np.arange(ni + 1, n + 1))
    return Problem(sp.csc_matrix(A), b, et), A


@pytest.mark.parametrize("ni,nb,cx", [(100, 60, False), (300, 200, False), (700, 400, False), (1100, 300, False),
                                      (100, 60, True), (300, 200, True), (700, 400, True), (1100, 300, True),
                                      (513, 0, True), (65, 1000, False), (33, 700, True)])
def test_single_dense_front(hs, ni, nb, cx):
    prob, A = _single_front_problem(hs, ni, nb, cx)
    Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
    assert np.array_equal(perm, np.arange(1, ni + nb + 1))
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    Aii, Aib, Abi, Abb = A[:ni, :ni], A[:ni, ni:], A[ni:, :ni], A[ni:, ni:]
    R = np.linalg.solve(Aii, Aib) if nb else np.zeros((ni, 0), A.dtype)
    L = np.linalg.solve(Aii.T, Abi.T).T if nb else np.zeros((0, ni), A.dtype)
    S = Abb - Abi @ R
    assert rel(F.D, Aii) < TOL
    if nb:
        assert rel(F.R, R) < TOL and rel(F.L, L) < TOL and rel(F.S, S) < TOL
    x = hs.ldiv(F, prob.b)
    assert rel(A @ x, prob.b) < TOL


@pytest.mark.parametrize("ni,nb,cx", [(1, 0, False), (1, 1, True), (0, 5, False), (64, 10, False), (65, 10, False),
                                      (63, 65, True), (128, 0, False), (129, 0, False), (96, 0, True), (97, 3, True),
                                      (256, 1, False), (257, 255, False),
                                      # row-per-thread leaf kernel (real, n ≤ 64) and its 8-column rotation blocks
                                      (8, 0, False), (9, 55, False), (16, 48, False), (49, 15, False), (64, 0, False), (33, 31, False),
                                      (40, 25, False),
                                      # block sizes of the diagonal-block inversion: ni ≤ 16 / 32 / 50 / 64, both scalar types
                                      (14, 20, False), (16, 10, True), (17, 3, True), (30, 60, False), (32, 5, True), (33, 10, True),
                                      (50, 10, False), (50, 30, True), (51, 5, False), (51, 40, True)])
def test_front_size_thresholds(hs, ni, nb, cx):
    """Sizes that sit on the switch points of the kernels: fused register kernels (n ≤ 64 rows-per-thread, n ≤ 128 f64 /
    96 c64 tiles), solve block (64) and the smaller inversion blocks of the bottom levels, panel width (64 / 32), outer
    block (256), one CTA vs cluster (256 rows); plus empty interior / boundary."""
    prob, A = _single_front_problem(hs, ni, nb, cx, seed=ni + 7 * nb)
    Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    x = hs.ldiv(F, prob.b)
    assert rel(A @ x, prob.b) < TOL
    if ni and nb:
        Aii, Aib, Abi, Abb = A[:ni, :ni], A[:ni, ni:], A[ni:, :ni], A[ni:, ni:]
        R = np.linalg.solve(Aii, Aib)
        assert rel(F.R, R) < TOL and rel(F.S, Abb - Abi @ R) < TOL and rel(F.D, Aii) < TOL
    X = hs.ldiv(F, np.stack([prob.b, 2 * prob.b, -prob.b], axis=1))
    assert rel(X[:, 1], 2 * x) < 1e-12 and rel(X[:, 2], -x) < 1e-12


def test_repeated_factor_free_cycles(hs):
    """Handles are released (no device-memory growth across factor / free cycles)."""
    import torch
    prob = hs.grid_problem((129, 129), "poisson")
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    del F
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(5):
        F = hs.factor(Ap, nd, nd_loc, swlevel=0)
        x = hs.ldiv(F, prob.b)
        del F
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 * 2 ** 20


def test_gmres_edge_cases(hs):
    prob = hs.grid_problem((33, 33), "poisson", nmax=40)
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    x, h = hs.gmres(Ap, prob.b, Pr=F, maxiter=0, log=True)           # no iterations allowed: x0 = 0 comes back
    assert h.iters == 0 and not h.isconverged and np.all(x == 0)
    x, h = hs.gmres(Ap, prob.b, Pr=None, restart=5, maxiter=12, log=True)   # restarts without preconditioner
    assert h.iters == 12 and len(h.resnorm) == 12
    x, h = hs.gmres(Ap, np.zeros_like(prob.b), Pr=F, log=True)        # zero right-hand side converges immediately
    assert h.iters == 0 and np.all(x == 0)


@pytest.mark.parametrize("kind", ["poisson", "helmholtz"])
def test_spmv_with_the_resident_matrix(hs, kind):
    """hs_spmv: y = A·x with the matrix a factorization holds (the mat-vec of the host-driven replicated GMRES)."""
    import ctypes as C
    import torch
    prob = hs.grid_problem((33, 31), kind, nmax=40)
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(Ap.shape[0]) + (1j * rng.standard_normal(Ap.shape[0]) if kind == "helmholtz" else 0)
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    yd = torch.empty_like(xd)
    torch.cuda.synchronize()
    hs._lib.check(hs._lib.lib.hs_spmv(F._hd.h, C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr())))
    hs.ldiv(F, prob.b)   # a synchronising call on the library's stream
    torch.cuda.synchronize()
    assert rel(yd.cpu().numpy(), Ap @ x) < 1e-13
    with pytest.raises(hs.ArgumentError):
        hs._lib.check(hs._lib.lib.hs_spmv(F._hd.h, C.c_void_p(xd.data_ptr()), C.c_void_p(xd.data_ptr())))


@pytest.mark.parametrize("kind", ["poisson", "helmholtz"])
def test_split_pivot_block_matches_unsplit(hs, kind):
    """Pivot blocks taller than one panel cluster (the 32 768-row root of the 128^3 problem in complex arithmetic) are
    eliminated as the reference's 2x2 `blockfactor` (blockmatrix.jl:115-120): pivoting inside A11 and inside S22.
    HS_PROW_CAP forces that path on a small problem; D, L, R, S are pivot-order independent, so both factorizations
    must agree to rounding."""
    import os
    prob = hs.grid_problem((513, 513), kind, nmax=100)
    Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
    F0 = hs.factor(Ap, nd, nd_loc, swlevel=0)
    x0 = hs.ldiv(F0, prob.b)
    os.environ["HS_PROW_CAP"] = "600"
    try:
        F1 = hs.factor(Ap, nd, nd_loc, swlevel=0)
    finally:
        os.environ.pop("HS_PROW_CAP", None)
    x1 = hs.ldiv(F1, prob.b)
    assert rel(Ap @ x1, prob.b) < TOL and rel(x1, x0) < 1e-9
    root = nd.nnodes - 1
    assert rel(F1.node(root).D, F0.node(root).D) < TOL
    k2 = int(nd.left[root])
    for name in ("L", "R", "S"):
        assert rel(getattr(F1.node(k2), name), getattr(F0.node(k2), name)) < 1e-9, name


def test_large_pivot_block_with_interchanges(hs):
    """Complex Helmholtz 1536^2: the root pivot block has 3072 rows and its LU really interchanges rows, so the forward
    sweep gathers P·x_int across super-blocks (rows pivoted out of the first 128 are read by other CTAs of the same
    launch).  Round 2 shipped a step-0 kernel that overwrote those entries of x too early; Poisson never pivots and
    did not see it.  Direct solve, two right-hand sides and one GMRES iteration must all reach rounding level."""
    prob = hs.grid_problem((1536, 1536), "helmholtz", nmax=100)
    Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    assert F.stats()["max_ni"] >= 3072
    for _ in range(3):
        x = hs.ldiv(F, prob.b)
        assert rel(Ap @ x, prob.b) < 1e-10
    B = np.stack([prob.b, prob.b[::-1]], axis=1)
    assert rel(Ap @ hs.ldiv(F, B), B) < 1e-10
    xg, h = hs.gmres(Ap, prob.b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True)
    assert h.isconverged and h.iters == 1 and rel(Ap @ xg, prob.b) < 1e-9
