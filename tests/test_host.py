"""CPU: host logic of the product — C ABI surface, symbolic phase (C++) against the oracle, problem I/O, options."""
import ctypes
import os
import sys
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(hs):
    hdr = open(os.path.join(ROOT, "include", "hsolve_cuda.h")).read()
    body = hdr[hdr.index("int32_t hs_version"):]
    names = set(re.findall(r"\b(hs_[a-z_0-9]+)\s*\(", body))
    assert len(names) >= 20
    lib = ctypes.CDLL(hs._lib.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/hsolve_cuda.h but not exported"
    assert {p[0] for p in hs._lib.PROTOTYPES} == names
    assert lib.hs_version() == int(re.search(r"#define HS_VERSION (\d+)", hdr).group(1))


def test_no_cpu_fallback(hs):
    """Without a GPU the compute entry points must fail loudly, never compute on the CPU."""
    if hs._lib.lib.hs_device_count() > 0:
        pytest.skip("GPU present")
    prob = hs.grid_problem((9, 9), "poisson", nmax=20)
    Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
    with pytest.raises(hs.HSolveError, match="no CUDA device"):
        hs.factor(Ap, nd, nd_loc, swlevel=0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hierarchicalsolvers.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "hs_oracle" not in txt and "oracle/" not in txt.replace("# oracle/", ""), f


@pytest.mark.parametrize("kind,shape,nmax", [("poisson", (65, 65), 100), ("helmholtz", (40, 23), 30),
                                             ("poisson", (12, 10, 9), 100), ("poisson", (5, 1), 100)])
def test_symfact_matches_oracle(hs, orc, kind, shape, nmax):
    prob = hs.grid_problem(shape, kind, nmax=nmax)
    Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
    Ao, ndo, ndo_loc, permo = orc.prepare(prob.A, prob.elim_tree)
    assert np.array_equal(perm, permo)
    assert abs(Ap - Ao).max() == 0
    nodes, locs = orc._postorder_nodes(ndo), orc._postorder_nodes(ndo_loc)
    assert len(nodes) == nd.nnodes and hs.depth(nd) == orc.depth(ndo)
    for k, (a, l) in enumerate(zip(nodes, locs)):
        v, vl = nd.node(k), nd_loc.node(k)
        assert np.array_equal(v.int, a.int) and np.array_equal(v.bnd, a.bnd)
        assert np.array_equal(vl.int, l.int) and np.array_equal(vl.bnd, l.bnd)
    # postorder makes interiors contiguous: getinterior's shortcut (nesteddissection.jl:100) holds
    assert list(hs.getinterior(nd)) == list(range(1, Ap.shape[0] + 1))
    assert len(hs.getboundary(nd)) == 0


def test_symfact_rejects_two_roots(hs):
    prob = hs.grid_problem((17, 17), "poisson", nmax=40)
    et = prob.elim_tree
    et.fathers = et.fathers.copy()
    et.fathers[1] = -1
    with pytest.raises(hs.ArgumentError):
        hs.from_elimtree(et)


def test_mat_roundtrip_and_parse(hs, tmp_path):
    prob = hs.grid_problem((17, 13), "helmholtz", nmax=30)
    p = str(tmp_path / "helmholtz2d_test.mat")
    hs.write_problem(p, prob)
    back = hs.read_problem(p)
    assert abs(back.A - prob.A).max() == 0 and np.array_equal(back.b, prob.b)
    for f in ("fathers", "lsons", "rsons", "inter_ptr", "inter_idx", "bound_ptr", "bound_idx"):
        assert np.array_equal(getattr(back.elim_tree, f), getattr(prob.elim_tree, f)), f
    pad = prob.elim_tree.to_padded()
    nd = hs.parse_elimtree(pad["fathers"], pad["lsons"], pad["rsons"], pad["ninter"], pad["inter"], pad["nbound"],
                           pad["bound"])
    nd, nd_loc = hs.symfact(nd)
    assert nd.nnodes == prob.elim_tree.nnodes


def test_golden_fixture(hs, orc):
    """tests/golden/poisson2d_17x17.npz was produced by tests/golden/make_golden.py from the oracle."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "poisson2d_17x17.npz"))
    prob = hs.grid_problem((17, 17), "poisson", nmax=40)
    Ao, ndo, ndo_loc, perm = orc.prepare(prob.A, prob.elim_tree)
    assert np.array_equal(perm, g["perm"])
    F = orc.factor(Ao, ndo, ndo_loc)
    x = orc.ldiv(F, prob.b)
    assert np.allclose(x, g["x"], rtol=0, atol=1e-13)
    assert np.allclose(F.S, g["root_S"]) and np.allclose(orc.nodes_postorder(F)[0].L, g["leaf0_L"], atol=1e-14)


def _golden_compressed(hs):
    import scipy.sparse as sp
    g = np.load(os.path.join(ROOT, "tests", "golden", "helmholtz2d_33x33_compressed.npz"))
    prob = hs.grid_problem((33, 33), "helmholtz", nmax=40)
    A = sp.csc_matrix(prob.A).copy()
    A.sort_indices()
    A.data = g["scale"].copy()          # stored in CSC order with sorted row indices
    opts = dict(swlevel=int(g["swlevel"]), swsize=int(g["swsize"]), atol=float(g["atol"]), rtol=float(g["rtol"]))
    return g, prob, A, opts


def test_golden_fixture_compressed(hs, orc):
    """tests/golden/helmholtz2d_33x33_compressed.npz: the oracle's compressed branch reproduces its committed output."""
    import hs_oracle_hss as oh
    g, prob, A, opts = _golden_compressed(hs)
    Ao, ndo, ndo_loc, perm = orc.prepare(A, prob.elim_tree)
    assert np.array_equal(perm, g["perm"])
    F = orc.factor(Ao, ndo, ndo_loc, **opts)
    assert np.array_equal(np.asarray(oh.node_ranks(F)), g["ranks"]) and g["ranks"].max() > 0
    assert np.allclose(orc.ldiv(F, prob.b), g["x"], rtol=0, atol=1e-12)
    nodec = orc.nodes_postorder(F)[int(g["node"])]
    assert np.allclose(nodec.L_dense(), g["node_L"], atol=1e-12) and np.allclose(nodec.R_dense(), g["node_R"], atol=1e-12)


@pytest.mark.gpu
def test_gpu_matches_golden_compressed(hs):
    """The CUDA path against the committed fixture (no oracle at run time)."""
    g, prob, A, opts = _golden_compressed(hs)
    Ap, nd, nd_loc, perm = hs.prepare(A, prob.elim_tree)
    assert np.array_equal(perm, g["perm"])
    F = hs.factor(Ap, nd, nd_loc, hss=False, **opts)     # the fixture holds the dense-Schur-complement form
    assert np.array_equal(np.asarray([F.node(k).ranks() for k in range(len(g["ranks"]))]), g["ranks"])
    x = hs.ldiv(F, prob.b)
    assert np.linalg.norm(x - g["x"]) / np.linalg.norm(g["x"]) < 1e-8
    nk = F.node(int(g["node"]))
    assert np.linalg.norm(nk.L - g["node_L"]) / np.linalg.norm(g["node_L"]) < 1e-8
    assert np.linalg.norm(nk.R - g["node_R"]) / np.linalg.norm(g["node_R"]) < 1e-8
    xs, ch = hs.gmres(Ap, prob.b, Pr=F, reltol=1e-9, restart=30, maxiter=30, log=True)
    assert ch.isconverged == bool(g["converged"]) and ch.iters == len(g["resnorm"])
    assert np.allclose(ch.resnorm, g["resnorm"], rtol=1e-5, atol=1e-12 * np.linalg.norm(prob.b))


def test_solver_options(hs):
    o = hs.SolverOptions()
    assert (o.swlevel, o.swsize, o.atol, o.rtol, o.c_tol, o.leafsize, o.kest, o.stepsize, o.verbose) == \
        (5, 1, 1e-6, 1e-6, 0.5, 32, -1, 10, False)                       # HierarchicalSolvers.jl:43-54
    o2 = o.copy(swlevel=0, atol=1e-2)
    assert o2.swlevel == 0 and o.swlevel == 5
    with pytest.raises(TypeError):
        o.copy(nonsense=1)
    for bad in (dict(swsize=0), dict(atol=-1.0), dict(rtol=-1.0), dict(c_tol=0.0), dict(c_tol=1.5), dict(leafsize=0)):
        with pytest.raises(hs.ArgumentError):
            hs.chkopts(o.copy(**bad))                                    # :73-79


def test_generator_invariants(hs):
    """Appendix A of SURVEY: leaves partition the DOFs, a branch's sets are the disjoint union of its children's
    boundaries, the root boundary is empty, A[leaf.inter, outside leaf] = 0."""
    prob = hs.grid_problem((21, 18), "poisson", nmax=30)
    et, A = prob.elim_tree, prob.A.tocsr()
    n = A.shape[0]
    seen = np.zeros(n, int)
    for i in range(et.nnodes):
        if et.lsons[i] == -1:
            ids = np.concatenate([et.inter(i), et.bound(i)]) - 1
            seen[ids] += 1
            inside = np.zeros(n, bool)
            inside[ids] = True
            rows = A[et.inter(i) - 1]
            assert inside[rows.indices].all()
        else:
            l, r = et.lsons[i] - 1, et.rsons[i] - 1
            u = np.sort(np.concatenate([et.bound(l), et.bound(r)]))
            assert np.array_equal(u, np.sort(np.concatenate([et.inter(i), et.bound(i)])))
            assert len(np.unique(u)) == len(u)
    assert (seen == 1).all()
    root = int(np.nonzero(et.fathers == -1)[0][0])
    assert len(et.bound(root)) == 0


def test_reference_arm_is_measured_and_never_loads_the_product_library():
    """`bench.py --impl reference` (the CPU arm): its value is the measured time of the grid its config names — no scaling to
    the full workload — and the process never maps libhsolve_cuda.so (only oracle/ and the pure-Python problem generator)."""
    import json
    import subprocess
    code = ("import sys, json, io, contextlib; sys.argv = ['bench.py', '--impl', 'reference', '--cpu-grid', '48', '--steps', '1', "
            "'--warmup', '0', '--cpu-extra', 'none'];\n"
            "import bench\n"
            "buf = io.StringIO()\n"
            "with contextlib.redirect_stdout(buf): bench.main()\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libhsolve_cuda' not in maps, 'product library mapped in the reference arm'\n"
            "assert 'hsolve_b200' not in sys.modules\n"
            "print(buf.getvalue().strip().splitlines()[-1])\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port"
    assert "48x48" in line["config"]["workload"] and line["config"]["n"] == 48 * 48
    assert line["value"] == line["e2e"]["value"] == line["cpu_baseline"]["value"] > 0
    assert line["residual"] < 1e-10 and line["gmres_iters"] <= 2
    assert "measured_sample_s" not in line and "flop ratio" not in json.dumps(line)


def test_solver_options_extensions_and_sketch_marshalling(hs):
    """The library's extensions of SolverOptions (hss, sketches, sketch_seed) reach the C struct; the reference's nine fields
    keep their defaults (HierarchicalSolvers.jl:43-54)."""
    from hsolve_b200.options import to_c
    o = hs.SolverOptions()
    assert o.hss is True and o.sketches is None and o.sketch_seed == 123
    c, keep = to_c(o)
    assert c.hss == 1 and not c.sketch_omega and c.sketch_rows == 0 and c.sketch_seed == 123 and keep == []
    Om, Ps = np.arange(12.0).reshape(4, 3), np.ones((4, 3))
    c, keep = to_c(o.copy(sketches=(Om, Ps), hss=False), dtype=np.complex128)
    assert c.hss == 0 and c.sketch_rows == 4 and c.sketch_cols == 3 and c.sketch_omega and c.sketch_psi
    assert keep[0].dtype == np.complex128 and keep[0].flags.f_contiguous and np.array_equal(keep[0].real, Om)
    with pytest.raises(hs.DimensionMismatch):
        to_c(o.copy(sketches=(Om, Ps[:3])))
