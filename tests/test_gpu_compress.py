"""GPU: compressed fronts (low-rank Gauss transforms, factorization.jl:78-112,171-182,228-249) against the oracle's
restatement of the same branch (oracle/hs_oracle_hss.py).  Both keep the Schur complements dense (HSS tolerance → 0).

Tolerances.  The oracle truncates with LAPACK's Householder QR with column pivoting, the library with a pivoted
Cholesky factorization of the Gram matrix: the same pivots and R in exact arithmetic.  On problems with randomly
perturbed coefficients (no exactly tied column norms) ranks must agree exactly and the factors to 1e-7; on symmetric
model problems a tie may be broken differently, which changes the truncated factors at the level of the compression
tolerance — there the ranks may differ by one and both are checked against the uncompressed factors instead."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _perturbed(hs, shape, kind, nmax, seed=0):
    prob = hs.grid_problem(shape, kind, nmax=nmax)
    rng = np.random.default_rng(seed)
    A = sp.csr_matrix(prob.A).copy()
    A.data = A.data * (1.0 + 0.3 * rng.random(A.nnz))
    prob.A = sp.csc_matrix(A)
    return prob


def hs_front_bytes_dense(hs, prob):
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    return hs.factor(Ap, nd, nd_loc, swlevel=0).stats()["front_bytes"]


def _both(hs, orc, prob, keep_schur=False, **opts):
    """Oracle and GPU factorizations of the same problem.  ``keep_schur``: keep the dense Schur blocks of compressed
    fronts (HS_KEEP_SCHUR=1) so that ``F.S`` of those nodes can be read back; by default their slots are transient."""
    import os
    opts.setdefault("hss", False)   # this file pins the low-rank Gauss transforms with dense Schur complements; test_gpu_hss.py the HSS form
    Ap, nd, nd_loc, perm = orc.prepare(prob.A, prob.elim_tree)
    Fo = orc.factor(Ap, nd, nd_loc, **opts)
    hnd = hs.from_elimtree(prob.elim_tree)
    hnd, hloc = hs.symfact(hnd)
    p = hs.postorder(hnd)
    A = hs.permute(prob.A, p, p)
    hnd = hs.permuted(hnd, hs.invperm(p))
    old = os.environ.pop("HS_KEEP_SCHUR", None)
    if keep_schur:
        os.environ["HS_KEEP_SCHUR"] = "1"
    try:
        F = hs.factor(A, hnd, hloc, **opts)
    finally:
        os.environ.pop("HS_KEEP_SCHUR", None)
        if old is not None:
            os.environ["HS_KEEP_SCHUR"] = old
    return Ap, Fo, F


# 1e-6 is the reference's default atol = rtol (HierarchicalSolvers.jl:45-46), halved to 5e-7 for the Gauss transforms
# (factorization.jl:99-100): within a factor 5 of the Gram-matrix floor of the library's pivoted QR, still rank-exact
@pytest.mark.parametrize("kind,shape,tol", [("poisson", (65, 65), 1e-2), ("poisson", (65, 65), 1e-5), ("poisson", (65, 65), 1e-6),
                                            ("helmholtz", (65, 65), 1e-6),
                                            ("helmholtz", (65, 65), 1e-3), ("poisson", (12, 11, 10), 1e-3)])
def test_compressed_nodes_match_oracle(hs, orc, kind, shape, tol):
    import hs_oracle_hss as oh
    prob = _perturbed(hs, shape, kind, nmax=40)
    Ap, Fo, F = _both(hs, orc, prob, keep_schur=True, swlevel=-2, swsize=16, atol=tol, rtol=tol)
    ranks_o = oh.node_ranks(Fo)
    nodes_o = orc.nodes_postorder(Fo)
    ncomp = 0
    for k, (no, ro) in enumerate(zip(nodes_o, ranks_o)):
        nk = F.node(k)
        assert nk.ranks() == ro, f"node {k}: ranks {nk.ranks()} vs oracle {ro}"
        if ro == (0, 0):
            continue
        ncomp += 1
        for name, Xo in (("D", no.D_dense()), ("L", no.L_dense()), ("R", no.R_dense()), ("S", no.S)):
            Xg = getattr(nk, name)
            assert Xg.shape == Xo.shape
            err = np.linalg.norm(Xg - Xo) / max(np.linalg.norm(Xo), 1e-300)
            assert err < 1e-7, f"node {k} {name}: rel err {err:.2e}"
    assert ncomp > 0
    assert hs.maxrank(F) == orc.maxrank(Fo) > 0
    # default mode: the dense slots of compressed fronts are transient (recycled two levels up) — same preconditioner,
    # F.S of a compressed node is no longer available
    _, _, F = _both(hs, orc, prob, swlevel=-2, swsize=16, atol=tol, rtol=tol)
    kc = next(k for k, r in enumerate(ranks_o) if r != (0, 0) and k != len(ranks_o) - 1)
    with pytest.raises(hs.ArgumentError):
        F.node(kc).S
    assert F.stats()["front_bytes"] < hs_front_bytes_dense(hs, prob)
    # the preconditioner application and the GMRES history
    b = prob.b
    xo, xg = orc.ldiv(Fo, b), hs.ldiv(F, b)
    assert np.linalg.norm(xg - xo) / np.linalg.norm(xo) < 1e-7
    B = np.random.default_rng(1).standard_normal((len(b), 3)).astype(xo.dtype)
    assert np.linalg.norm(hs.ldiv(F, B) - orc.ldiv(Fo, B)) / np.linalg.norm(B) < 1e-6
    _, reso, convo = orc.gmres(Ap, b, Pr=lambda v: orc.ldiv(Fo, v), reltol=1e-9, restart=30, maxiter=30)
    xs, ch = hs.gmres(sp.csc_matrix(Ap), b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True)
    assert ch.isconverged == convo and ch.iters == len(reso)
    # history: entries agree to 1e-3 relative, or absolutely below GMRES's own stopping threshold reltol*|b| — with a
    # preconditioner that is only accurate to `tol`, the late residuals are set by the last digits of the truncated factors
    # (which differ at the 1e-7 level between a Gram-matrix and a Householder pivoted QR)
    assert np.allclose(np.asarray(ch.resnorm), np.asarray(reso), rtol=1e-3, atol=1e-9 * np.linalg.norm(b))


@pytest.mark.parametrize("kind", ["poisson", "helmholtz"])
def test_compressed_symmetric_model_problem(hs, orc, kind):
    """Unperturbed grid operators have tied column norms: compare against the exact factors at the tolerance."""
    tol = 1e-3
    prob = hs.grid_problem((65, 65), kind, nmax=40)
    Ap, Fo, F = _both(hs, orc, prob, swlevel=-2, swsize=16, atol=tol, rtol=tol)
    import hs_oracle_hss as oh
    ro = oh.node_ranks(Fo)
    for k, r in enumerate(ro):
        rg = F.node(k).ranks()
        assert abs(rg[0] - r[0]) <= 1 and abs(rg[1] - r[1]) <= 1, (k, rg, r)
    assert abs(hs.maxrank(F) - orc.maxrank(Fo)) <= 1 and hs.maxrank(F) > 0
    b = prob.b
    xg, xo = hs.ldiv(F, b), orc.ldiv(Fo, b)
    rg = np.linalg.norm(Ap @ xg - b) / np.linalg.norm(b)
    rr = np.linalg.norm(Ap @ xo - b) / np.linalg.norm(b)
    assert rg < 20 * max(rr, tol)          # an approximate inverse of the quality the oracle reaches
    _, reso, convo = orc.gmres(Ap, b, Pr=lambda v: orc.ldiv(Fo, v), reltol=1e-9, restart=30, maxiter=30)
    xs, ch = hs.gmres(sp.csc_matrix(Ap), b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True)
    assert ch.isconverged and abs(ch.iters - len(reso)) <= 1
    assert np.linalg.norm(Ap @ xs - b) / np.linalg.norm(b) < 1e-8


def test_compression_options_semantics(hs, orc):
    """swlevel / swsize decide which nodes are compressed exactly as factorization.jl:8,15."""
    prob = _perturbed(hs, (65, 65), "poisson", nmax=40)
    Ap, nd, nd_loc, _ = orc.prepare(prob.A, prob.elim_tree)
    hnd, hloc = hs.symfact(hs.from_elimtree(prob.elim_tree))
    p = hs.postorder(hnd)
    A = hs.permute(prob.A, p, p)
    hnd = hs.permuted(hnd, hs.invperm(p))
    import hs_oracle_hss as oh
    for sw, size in [(2, 1), (3, 40), (-3, 1), (50, 10 ** 6)]:
        Fo = orc.factor(Ap, nd, nd_loc, swlevel=sw, swsize=size, atol=1e-4, rtol=1e-4)
        F = hs.factor(A, hnd, hloc, swlevel=sw, swsize=size, atol=1e-4, rtol=1e-4, hss=False)
        ro = oh.node_ranks(Fo)
        rg = [F.node(k).ranks() for k in range(len(ro))]
        assert [r != (0, 0) for r in rg] == [r != (0, 0) for r in ro]
        assert hs.maxrank(F) == orc.maxrank(Fo)
    # refactor keeps working on a compressed factorization
    F = hs.factor(A, hnd, hloc, swlevel=3, swsize=1, atol=1e-6, rtol=1e-6, hss=False)
    x1 = hs.ldiv(F, prob.b)
    F.refactor(A)
    assert np.allclose(hs.ldiv(F, prob.b), x1, rtol=1e-12, atol=0)


def test_baseline_config2_standin(hs, orc):
    """BASELINE config 2 (poisson2d_p1_h128_nmax100 + compression, test/rungmres.jl:39) on the 129×129 stand-in.  With
    rungmres.jl's swsize = 480 no boundary of the 5-point stand-in qualifies, so swsize = 64 is used;
    atol = rtol = 1e-2, swlevel = -2 as in the script."""
    import hs_oracle_hss as oh
    prob = hs.grid_problem((129, 129), "poisson", nmax=100)
    opts = dict(swlevel=-2, swsize=64, atol=1e-2, rtol=1e-2, kest=200, stepsize=100, leafsize=120)
    Ap, Fo, F = _both(hs, orc, prob, **opts)
    assert hs.maxrank(F) > 0 and abs(hs.maxrank(F) - orc.maxrank(Fo)) <= 1
    ro = oh.node_ranks(Fo)
    assert sum(r != (0, 0) for r in ro) == sum(F.node(k).ranks() != (0, 0) for k in range(len(ro))) > 0
    b = prob.b
    _, reso, convo = orc.gmres(Ap, b, Pr=lambda v: orc.ldiv(Fo, v), reltol=1e-9, restart=30, maxiter=30)
    xs, ch = hs.gmres(sp.csc_matrix(Ap), b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True)
    assert ch.isconverged == convo and abs(ch.iters - len(reso)) <= 1
    # same preconditioner quality: residual histories agree to the compression tolerance
    m = min(ch.iters, len(reso))
    assert np.allclose(np.log10(np.asarray(ch.resnorm)[:m]), np.log10(np.asarray(reso)[:m]), atol=0.5)


@pytest.mark.parametrize("kind", ["poisson", "helmholtz"])
def test_compressed_root_with_boundary(hs, orc, kind):
    """A root that keeps a boundary is compressed like any other node (factorization.jl:15) and its dense Schur
    complement is then factored for `C[F.bnd,:] = F.S \\ C[F.bnd,:]` (factornode.jl:72)."""
    from test_gpu_parity import _subtree_problem
    prob = _subtree_problem(hs, (49, 49), kind)
    rng = np.random.default_rng(5)
    A = sp.csr_matrix(prob.A).copy()
    A.data = A.data * (1.0 + 0.3 * rng.random(A.nnz))
    prob.A = sp.csc_matrix(A)
    opts = dict(swlevel=2, swsize=8, atol=1e-4, rtol=1e-4)
    Ap, Fo, F = _both(hs, orc, prob, **opts)
    assert F.ranks() == (Fo.L.rank, Fo.R.rank) and F.ranks()[0] > 0        # the root itself is compressed
    assert len(Fo.bnd) > 0
    for name, Xo in (("D", Fo.D_dense()), ("L", Fo.L_dense()), ("R", Fo.R_dense()), ("S", Fo.S)):
        Xg = getattr(F, name)
        assert np.linalg.norm(Xg - Xo) / np.linalg.norm(Xo) < 1e-7, name
    B = rng.standard_normal((Ap.shape[0], 2)).astype(Fo.S.dtype)
    assert np.linalg.norm(hs.ldiv(F, B) - orc.ldiv(Fo, B)) / np.linalg.norm(B) < 1e-7


@pytest.mark.parametrize("cplx", [False, True])
def test_hss_solve_through_the_multifrontal_kernels(hs, orc, cplx):
    """Round-2 groundwork: an HSS matrix (oracle/hs_hss.py) embedded as a sparse system whose elimination tree is the
    HSS tree is factored and solved by the CUDA multifrontal path as it is — O(n·r²) instead of O(n³) for the block."""
    import hs_hss as H
    from test_oracle import _embedding_problem
    h, Aext, nx, et, b, bext = _embedding_problem(hs, H, cplx, n=1200, leafsize=64, tol=1e-9)
    Ap, nd, nd_loc, perm = hs.prepare(Aext, et)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    xp = hs.ldiv(F, bext[perm - 1])
    x = np.empty_like(xp)
    x[perm - 1] = xp
    xh = H.solve(h, b)
    assert np.linalg.norm(x[:nx] - xh) / np.linalg.norm(xh) < 1e-10
    assert F.stats()["max_ni"] + F.stats()["max_nb"] < nx // 3
