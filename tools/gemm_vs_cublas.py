"""k_gemm against cuBLAS on the launch shape that carries the flops: the trailing update C(m×m) -= A(m×256)·B(256×m).
A single dense front with ni = 256 pivot rows and nb = m boundary rows does exactly one such launch; HS_PROFILE=1 event-times
it inside the library.  cuBLAS: torch.addmm_ on the same shape (torch is tooling here, not on the product path).
    HS_PROFILE=1 python tools/gemm_vs_cublas.py [m ...]"""
import os
import sys

os.environ.setdefault("HS_PROFILE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402
import _pkg  # noqa: E402

hs = _pkg.load()
from hsolve_b200.problems import ElimTree, Problem  # noqa: E402

K = 256


def front(ni, nb, cx):
    rng = np.random.default_rng(1)
    n = ni + nb
    A = rng.standard_normal((n, n))
    if cx:
        A = A + 1j * rng.standard_normal((n, n))
    A = A + np.diag(np.full(n, 2.0 * np.sqrt(n)))
    et = ElimTree(np.array([-1]), np.array([-1]), np.array([-1]), np.array([0, ni]), np.arange(1, ni + 1),
                  np.array([0, nb]), np.arange(ni + 1, n + 1))
    return Problem(sp.csc_matrix(A), np.ones(n, dtype=A.dtype), et)


def factor_subtree(Ap, nd, nd_loc, cx):
    """hs_factor in subtree mode (the root keeps its Schur block unfactored, as on a rank of the multi-GPU path): the only
    big update of the factorization is the one measured here."""
    import ctypes as C
    from hsolve_b200 import _lib
    from hsolve_b200.options import SolverOptions, to_c
    from hsolve_b200.parallel import _tree_struct
    dtype = np.complex128 if cx else np.float64
    nd._need_analyzed()
    tree, keep = _tree_struct(nd, nd_loc)
    ctx = _lib.default_context(0)
    copts, keep2 = to_c(SolverOptions(swlevel=0), subtree=True, dtype=dtype)
    h = C.c_void_p()
    A = Ap.tocsc()
    colptr, rowval = _lib.as_i64(A.indptr), _lib.as_i64(A.indices)
    nzval = np.ascontiguousarray(A.data, dtype=dtype)
    _lib.check(_lib.lib.hs_factor(ctx, _lib.HS_C64 if cx else _lib.HS_F64, A.shape[0], colptr.ctypes.data_as(C.c_void_p),
                                  rowval.ctypes.data_as(C.c_void_p), nzval.ctypes.data_as(C.c_void_p), C.byref(tree),
                                  C.byref(copts), _lib.HS_CSC_ZERO_BASED, C.byref(h)))
    s = _lib.hs_stats_t()
    _lib.check(_lib.lib.hs_stats(h, C.byref(s)))
    out = s.asdict()
    _lib.lib.hs_factor_free(h)
    return out


def cublas(m, cx, reps=20):
    dt = torch.complex128 if cx else torch.float64
    C = torch.randn(m, m, dtype=dt, device="cuda"); A = torch.randn(m, K, dtype=dt, device="cuda"); B = torch.randn(K, m, dtype=dt, device="cuda")
    for _ in range(3):
        C.addmm_(A, B, alpha=-1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        C.addmm_(A, B, alpha=-1.0)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"# C(m x m) -= A(m x {K}) B({K} x m); TFLOP/s (complex: 8 m^2 K flops)")
print(f"{'dtype':5s} {'m':>6s} {'k_gemm ms':>10s} {'TF/s':>7s} {'cuBLAS ms':>10s} {'TF/s':>7s} {'k_gemm/cuBLAS':>14s}")
for cx in (False, True):
    for m in [int(a) for a in sys.argv[1:]] or [1024, 2048, 3072, 4096]:
        prob = front(K, m, cx)
        Ap, nd, nd_loc, _ = hs.prepare(prob.A, prob.elim_tree)
        best = None
        for _ in range(3):
            st = factor_subtree(Ap, nd, nd_loc, cx)
            t, fl = st["ms_gemm_big"], st["gemm_flops_big"]
            best = t if best is None else min(best, t)
        tc = cublas(m, cx)
        flc = (8.0 if cx else 2.0) * m * m * K
        print(f"{'c64' if cx else 'f64':5s} {m:6d} {best:10.3f} {fl / best / 1e9:7.2f} {tc:10.3f} {flc / tc / 1e9:7.2f} {tc / best:14.2f}")
