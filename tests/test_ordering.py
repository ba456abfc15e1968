"""``nested_dissection`` (hs_nd_create, csrc/hs_ordering.cpp): elimination trees for matrices that come without an
``elim_tree``.  CPU: schema invariants the reference's ``_symfact!`` / ``_factor_*`` rely on, and a direct
solve through the oracle against SuperLU.  GPU: the same trees through the CUDA path."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _mesh_matrix(n=2500, seed=0, unsym=False, cplx=False):
    """Graph Laplacian + identity of a random geometric graph (an unstructured 2D 'mesh'), optionally with an
    unsymmetric pattern (one triangle of some couplings dropped) and complex shifts."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2))
    pairs = cKDTree(pts).query_pairs(1.75 / np.sqrt(n), output_type="ndarray")
    w = rng.random(len(pairs)) + 0.5
    i, j = pairs[:, 0], pairs[:, 1]
    lo = -w.copy()
    if unsym:
        lo[rng.random(len(w)) < 0.3] = 0.0
    A = sp.coo_matrix((np.r_[-w, lo], (np.r_[i, j], np.r_[j, i])), shape=(n, n)).tocsr()
    A.eliminate_zeros()
    d = np.asarray(abs(A).sum(1)).ravel() + 1.0
    A = (A + sp.diags(d)).astype(np.complex128 if cplx else np.float64)
    if cplx:
        A = A + 0.3j * sp.diags(rng.random(n))
    return sp.csc_matrix(A), (rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0))


def _check_schema(A, et):
    n = A.shape[0]
    P = sp.csr_matrix((abs(A) + abs(A).T) != 0)
    nn = et.nnodes
    leaves = [i for i in range(nn) if et.lsons[i] == -1]
    allv = np.concatenate([np.r_[et.inter(i), et.bound(i)] for i in leaves])
    assert np.array_equal(np.sort(allv), np.arange(1, n + 1))            # leaves partition the DOFs
    assert int(np.sum(et.fathers == -1)) == 1
    root = int(np.nonzero(et.fathers == -1)[0][0])
    assert len(et.bound(root)) == 0                                        # nothing outside the root
    sub = {}
    for i in range(nn - 1, -1, -1):                                        # children have larger ids than fathers
        if et.lsons[i] == -1:
            sub[i] = set((np.r_[et.inter(i), et.bound(i)] - 1).tolist())
        else:
            l, r = int(et.lsons[i]) - 1, int(et.rsons[i]) - 1
            assert et.fathers[l] == i + 1 and et.fathers[r] == i + 1
            sub[i] = sub[l] | sub[r]
            both = np.sort(np.r_[et.inter(i), et.bound(i)])
            assert np.array_equal(both, np.sort(np.r_[et.bound(l), et.bound(r)]))   # factorization.jl:63-64
    for i in range(nn):
        inside = sub[i]
        for v in et.inter(i) - 1:                                          # A[inter, outside the subtree] = 0
            assert all(int(u) in inside for u in P.indices[P.indptr[v]:P.indptr[v + 1]])
        for v in et.bound(i) - 1:
            assert any(int(u) not in inside for u in P.indices[P.indptr[v]:P.indptr[v + 1]])


@pytest.mark.parametrize("unsym,cplx", [(False, False), (True, False), (False, True)])
def test_nested_dissection_schema_and_oracle_solve(hs, orc, unsym, cplx):
    A, b = _mesh_matrix(1500, seed=1, unsym=unsym, cplx=cplx)
    et = hs.nested_dissection(A, nmax=50)
    assert et.nnodes > 15 and max(len(et.inter(i)) + len(et.bound(i)) for i in range(et.nnodes) if et.lsons[i] == -1) <= 50
    _check_schema(A, et)
    Ap, nd, nd_loc, perm = orc.prepare(A, et)
    x = orc.ldiv(orc.factor(Ap, nd, nd_loc), b)
    xr = spla.splu(sp.csc_matrix(Ap)).solve(b)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-10


def test_nested_dissection_degenerate_inputs(hs):
    # diagonal matrix: no edges at all — halved by position
    et = hs.nested_dissection(sp.identity(37, format="csc"), nmax=5)
    _check_schema(sp.identity(37, format="csc"), et)
    # already small: a single leaf
    A, _ = _mesh_matrix(40, seed=2)
    et = hs.nested_dissection(A, nmax=100)
    assert et.nnodes == 1 and len(et.bound(0)) == 0 and len(et.inter(0)) == 40
    with pytest.raises(hs.ArgumentError):
        hs.nested_dissection(A, nmax=0)
    with pytest.raises(hs.DimensionMismatch):
        hs.nested_dissection(sp.csc_matrix(np.ones((3, 4))))
    # reproducible
    A, _ = _mesh_matrix(800, seed=3)
    e1, e2 = hs.nested_dissection(A, nmax=40), hs.nested_dissection(A, nmax=40)
    assert np.array_equal(e1.inter_idx, e2.inter_idx) and np.array_equal(e1.bound_idx, e2.bound_idx)


@pytest.mark.gpu
@pytest.mark.parametrize("unsym,cplx", [(False, False), (True, False), (False, True)])
def test_gpu_factor_on_metis_tree_matches_oracle(hs, orc, unsym, cplx):
    A, b = _mesh_matrix(2500, seed=4, unsym=unsym, cplx=cplx)
    et = hs.nested_dissection(A, nmax=60)
    Ap, nd, nd_loc, perm = hs.prepare(A, et)
    F = hs.factor(Ap, nd, nd_loc, swlevel=0)
    Ao, ndo, ndo_loc, _ = orc.prepare(A, et)
    Fo = orc.factor(Ao, ndo, ndo_loc)
    x, xo = hs.ldiv(F, b), orc.ldiv(Fo, b)
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
    assert np.linalg.norm(Ap @ x - b) / np.linalg.norm(b) < 1e-10
    for k in (0, F._hd.nd.nnodes // 2, F._hd.nd.nnodes - 1):
        nk, no = F.node(k), orc.nodes_postorder(Fo)[k]
        for name, Xo in (("D", no.D_dense()), ("L", no.L_dense()), ("R", no.R_dense()), ("S", no.S)):
            Xg = getattr(nk, name)
            assert Xg.shape == Xo.shape
            if Xo.size:
                assert np.linalg.norm(Xg - Xo) / max(np.linalg.norm(Xo), 1e-300) < 1e-10, (k, name)
    xs, ch = hs.gmres(Ap, b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True)
    assert ch.isconverged and ch.iters <= 2
    # compressed upper fronts on an unstructured tree
    Fc = hs.factor(Ap, nd, nd_loc, swlevel=-2, swsize=24, atol=1e-6, rtol=1e-6)
    Foc = orc.factor(Ao, ndo, ndo_loc, swlevel=-2, swsize=24, atol=1e-6, rtol=1e-6)
    assert hs.maxrank(Fc) > 0 and abs(hs.maxrank(Fc) - orc.maxrank(Foc)) <= 1
    xc, xoc = hs.ldiv(Fc, b), orc.ldiv(Foc, b)
    assert np.linalg.norm(xc - xoc) / np.linalg.norm(xoc) < 1e-4


@pytest.mark.parametrize("seed,unsym", [(11, False), (12, True), (13, False)])
def test_symbolic_phase_on_metis_trees_matches_oracle(hs, orc, seed, unsym):
    """The C++ symbolic phase (hs_symfact) against the oracle's ``symfact!`` / ``postorder`` / ``permuted!`` on the
    irregular trees METIS produces (unbalanced subtrees, leaves at different depths)."""
    A, _ = _mesh_matrix(900 + 37 * seed, seed=seed, unsym=unsym)
    et = hs.nested_dissection(A, nmax=35)
    Ap, nd, nd_loc, perm = hs.prepare(A, et)
    Ao, ndo, ndo_loc, permo = orc.prepare(A, et)
    assert np.array_equal(perm, permo) and abs(Ap - Ao).max() == 0
    nodes, locs = orc._postorder_nodes(ndo), orc._postorder_nodes(ndo_loc)
    assert len(nodes) == nd.nnodes and hs.depth(nd) == orc.depth(ndo)
    for k, (a, l) in enumerate(zip(nodes, locs)):
        v, vl = nd.node(k), nd_loc.node(k)
        assert np.array_equal(v.int, a.int) and np.array_equal(v.bnd, a.bnd)
        assert np.array_equal(vl.int, l.int) and np.array_equal(vl.bnd, l.bnd)
