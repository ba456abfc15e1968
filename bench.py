#!/usr/bin/env python
"""Benchmark of the hot path: multifrontal nested-dissection factor + preconditioned GMRES solve.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port on the host cores

A "step" is one pass of the hot path over the workload: numeric factorization of A (factor, factorization.jl:5-11)
followed by the driver's solve, GMRES(30) to reltol 1e-9 with the factorization as right preconditioner
(test/rungmres.jl:47).  Default workload = BASELINE.json configs[3], the largest configuration that fits one B200:
synthetic 2D 5-point Laplacian on a 2048×2048 grid, geometric nested dissection with nmax = 100, float64.
`--workload 3d --grid3 G` switches to configs[4], the 3D 7-point complex Helmholtz problem on a G³ grid.

Every number printed is MEASURED on the grid the line names — nothing is extrapolated:
  * our arm: `value` (inputs resident in HBM), `e2e` (host CSC in, host x out), `roofline`, `roofline_hbm`,
    `c64` (the complex Helmholtz 2048² workload, secondary), `cpu_baseline` and `like_for_like` (the oracle port AND this
    library on the same bounded sample grid, in the same run);
  * the reference arm (`--impl reference`): the oracle port of the reference algorithm (the reference itself is Julia
    plus un-vendored packages and cannot run here) on a bounded SAMPLE grid that `config.workload` names; it never
    loads the product library.

Prints ONE JSON line (contract in the task statement).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "factor+solve time"
UNIT = "s"
GMRES_NOTE = ("GMRES solution update: x += y1*z1 with z1 = Pr^-1 v1 kept from the Arnoldi step (one preconditioner application "
              "per iteration when it converges in one step); IterativeSolvers applies Pr^-1 to V*y once more - same x by "
              "linearity, about half the solve time")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="2d", choices=["2d", "3d"])
    ap.add_argument("--grid", type=int, default=2048, help="grid side of the synthetic 2D problem")
    ap.add_argument("--grid3", type=int, default=64, help="grid side of the synthetic 3D problem (--workload 3d)")
    ap.add_argument("--kind", default=None, choices=["poisson", "helmholtz"], help="default: poisson (2d), helmholtz (3d)")
    ap.add_argument("--nmax", type=int, default=100)
    ap.add_argument("--cpu-grid", type=int, default=0, help="side of the bounded CPU sample grid (0 = 256 in 2D, 20 in 3D)")
    ap.add_argument("--cpu-extra", default="auto", help="reference arm: extra single measured runs on larger grids, comma-separated "
                                                        "sides, 'auto' (512 in 2D / 28 in 3D when time allows) or 'none'")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-compressed", action="store_true", help="skip the informational compressed-fronts variant")
    ap.add_argument("--no-c64", action="store_true", help="skip the secondary complex Helmholtz 2048^2 block")
    args = ap.parse_args()
    if args.kind is None:
        args.kind = "helmholtz" if args.workload == "3d" else "poisson"
    if args.cpu_grid <= 0:
        args.cpu_grid = 20 if args.workload == "3d" else 256
    return args


def shape_of(args, side=None):
    if args.workload == "3d":
        s = side or args.grid3
        return (s, s, s)
    s = side or args.grid
    return (s, s)


def workload_name(args, shape, kind, extra=""):
    cx = kind == "helmholtz"
    dims = "x".join(str(s) for s in shape)
    op = ("7-point" if len(shape) == 3 else "5-point") + (" Helmholtz (complex, 10 ppw)" if cx else " Laplacian")
    return (f"synthetic {len(shape)}D {op} {dims}, geometric nested dissection nmax={args.nmax}, uncompressed (swlevel=0), "
            f"GMRES(30) reltol 1e-9{extra}")


def front_flops(ni, nb, complex_):
    """Σ ⅔ni³ + 2ni²nb + 2ni·nb² over the fronts, ×4 for complex."""
    ni = np.asarray(ni, dtype=np.float64)
    nb = np.asarray(nb, dtype=np.float64)
    f = (2.0 / 3.0) * ni ** 3 + 2.0 * ni ** 2 * nb + 2.0 * ni * nb ** 2
    return float(f.sum()) * (4.0 if complex_ else 1.0)


def factor_flops(et, complex_):
    return front_flops(et.ninter(), et.nbound(), complex_)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference algorithm; the reference itself is Julia and cannot run here).
# Uses ONLY oracle/ and the pure-Python problem generator — the product package (and its .so) is never imported.
# ------------------------------------------------------------------------------------------------
def load_problem_generator():
    """hierarchicalsolvers.jl_b200/problems.py as a stand-alone module: host-side integer/sparse bookkeeping with no
    dependency on the CUDA library (its only use of it, `nested_dissection`, imports it lazily)."""
    name = "hs_problems_standalone"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "hierarchicalsolvers.jl_b200", "problems.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuOracle:
    """The oracle port on one problem: factor (factorization.jl:5-75) + GMRES with ldiv! as Pr (rungmres.jl:47)."""

    def __init__(self, shape, kind, nmax):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import hs_oracle as orc
        P = load_problem_generator()
        self.orc = orc
        self.shape, self.kind = tuple(shape), kind
        self.prob = P.grid_problem(self.shape, kind, nmax=nmax)
        self.Ao, self.nd, self.nd_loc, _ = orc.prepare(self.prob.A, self.prob.elim_tree)
        self.flops = factor_flops(self.prob.elim_tree, kind == "helmholtz")
        self.n = self.Ao.shape[0]

    def step(self, threads):
        from threadpoolctl import threadpool_limits
        orc = self.orc
        with threadpool_limits(limits=threads):
            t0 = time.perf_counter()
            F = orc.factor(self.Ao, self.nd, self.nd_loc)
            t1 = time.perf_counter()
            x, res, conv = orc.gmres(self.Ao, self.prob.b, Pr=lambda v: orc.ldiv(F, v), reltol=1e-9, restart=30, maxiter=30)
            t2 = time.perf_counter()
        resid = float(np.linalg.norm(self.Ao @ x - self.prob.b) / np.linalg.norm(self.prob.b))
        return {"seconds": t2 - t0, "factor_s": t1 - t0, "solve_s": t2 - t1, "gmres_iters": len(res), "residual": resid}

    def splu(self):
        """Second CPU yardstick: SuperLU on the same (nested-dissection permuted) matrix."""
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        t0 = time.perf_counter()
        lu = spla.splu(sp.csc_matrix(self.Ao), permc_spec="NATURAL")   # the matrix already is in ND post-order
        t1 = time.perf_counter()
        x = lu.solve(np.asarray(self.prob.b, dtype=self.Ao.dtype))
        t2 = time.perf_counter()
        return {"factor_s": t1 - t0, "solve_s": t2 - t1, "seconds": t2 - t0,
                "residual": float(np.linalg.norm(self.Ao @ x - self.prob.b) / np.linalg.norm(self.prob.b))}


def pick_threads(cpu: CpuOracle):
    """LAPACK thread count by measurement: the port is bound by per-front Python/sparse-slicing overhead on small grids,
    where one thread beats an oversubscribed pool.  Returns (threads, {threads: seconds})."""
    cores = host_cores()
    trials = {}
    for t in sorted({1, cores}):
        trials[t] = cpu.step(t)["seconds"]
    best = min(trials, key=trials.get)
    return best, trials


def cpu_sample(args, steps, warmup, with_splu=True):
    """Measured oracle-port timings on the bounded sample grid."""
    shape = shape_of(args, args.cpu_grid if args.workload == "3d" else min(args.cpu_grid, args.grid))
    cpu = CpuOracle(shape, args.kind, args.nmax)
    threads, trials = pick_threads(cpu)          # also serves as warm-up (two full steps)
    for _ in range(max(0, warmup - len(trials))):
        cpu.step(threads)
    runs = [cpu.step(threads) for _ in range(steps)]
    out = {"shape": list(shape), "n": cpu.n, "flops": cpu.flops, "threads": threads, "cores": host_cores(),
           "thread_trials_s": {str(k): v for k, v in trials.items()},
           "seconds": float(np.mean([r["seconds"] for r in runs])), "factor_s": float(np.mean([r["factor_s"] for r in runs])),
           "solve_s": float(np.mean([r["solve_s"] for r in runs])), "gmres_iters": runs[-1]["gmres_iters"],
           "residual": runs[-1]["residual"], "steps": steps}
    if with_splu:
        try:
            out["splu"] = cpu.splu()
        except Exception as exc:
            out["splu"] = {"error": repr(exc)}
    return out, cpu


def reference_arm(args):
    """`--impl reference`: the oracle port on the host cores.  value = measured seconds per step on the SAMPLE grid named
    in config.workload (no scaling to the full workload)."""
    t_start = time.perf_counter()
    r, cpu = cpu_sample(args, steps=max(1, args.steps), warmup=max(0, args.warmup))
    shape = tuple(r["shape"])
    extra = []
    if args.cpu_extra != "none":
        sides = ([512] if args.workload == "2d" else [28]) if args.cpu_extra == "auto" else [int(s) for s in args.cpu_extra.split(",") if s]
        for s in sides:
            if args.cpu_extra == "auto" and time.perf_counter() - t_start > 150:
                break       # keep the whole run within a few minutes
            if s <= shape[0]:
                continue
            c2 = CpuOracle(shape_of(args, s), args.kind, args.nmax)
            m = c2.step(r["threads"])
            m.update({"shape": list(c2.shape), "n": c2.n, "flops": c2.flops, "threads": r["threads"]})
            extra.append(m)
    sample = (f"oracle port (NumPy/SciPy restatement of the reference - the Julia reference cannot run here), LAPACK threads={r['threads']} "
              f"of {r['cores']} cores chosen by measurement {r['thread_trials_s']}; {r['steps']} measured steps on the sample grid "
              f"{'x'.join(map(str, shape))} (N={r['n']}, {r['flops'] / 1e9:.2f} GFLOP): {r['seconds']:.3f} s/step = {r['factor_s']:.3f} factor + "
              f"{r['solve_s']:.3f} GMRES ({r['gmres_iters']} it).  NOT scaled to the full workload.")
    full_shape = shape_of(args)
    config = {"workload": workload_name(args, shape, args.kind, extra=" - bounded CPU sample of the " + "x".join(map(str, full_shape)) + " workload"),
              "n": r["n"], "sample_of": workload_name(args, full_shape, args.kind)}
    line = {"impl": "reference", "metric": METRIC, "value": r["seconds"], "unit": UNIT, "residual": r["residual"],
            "gmres_iters": r["gmres_iters"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["seconds"] * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "c64" if args.kind == "helmholtz" else "f64", "data": "synthetic", "config": config,
            "factor_s": r["factor_s"], "solve_s": r["solve_s"],
            "cpu_baseline": {"value": r["seconds"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample,
                             "host_cores": r["cores"]},
            "e2e": {"value": r["seconds"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "splu_same_grid": r.get("splu"), "larger_grids_single_run": extra,
            "note": "value is the measured time of the grid named in config.workload; compare with the `like_for_like` block of "
                    "the CUDA arm (same grid, same run), not with its 2048x2048 value"}
    print(json.dumps(line), flush=True)


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak(complex_):
    """FP64 tensor denominator.  MEASURED_PEAKS.json carries no FP64 figure, so the denominator is cuBLAS
    DGEMM/ZGEMM 8192³ measured on this pool's B200 by tools/fp64_peak.py (profiles/fp64_peaks.json)."""
    key = "zgemm_8192_tflops" if complex_ else "dgemm_8192_tflops"
    for p in (os.path.join(ROOT, "profiles", "fp64_peaks.json"), os.path.join(ROOT, "gpurun_out", "fp64_peaks.json")):
        if os.path.exists(p):
            return float(json.load(open(p))[key]), f"measured cuBLAS {'ZGEMM' if complex_ else 'DGEMM'} 8192^3 ({os.path.relpath(p, ROOT)})"
    return (37.0 if complex_ else 35.5), "fallback: earlier cuBLAS measurement on this pool"


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6456.2, "fallback 6456.2 GB/s (earlier MEASURED_PEAKS.json of this pool)"


TRAFFIC_FILES = ("r02_traffic_2048.json", "r01_traffic_2048.json")


def ncu_traffic(kernel, per_launch=True):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu pass over the default workload
    (profiles/r0N_traffic_2048.json, tools/profile_run.py 2048); None for any other workload."""
    for fn in TRAFFIC_FILES:
        try:
            k = json.load(open(os.path.join(ROOT, "profiles", fn)))["kernels"][kernel]
            return k["dram_bytes_per_launch"] if per_launch else k["dram_bytes_total"]
        except Exception:
            continue
    return None


DEFAULT_WORKLOAD = True


def gemm_roofline(stp, peak, peak_src):
    gemm_tf = stp["gemm_flops"] / (stp["ms_gemm"] * 1e-3) / 1e12 if stp["ms_gemm"] > 0 else 0.0
    split = None
    if stp.get("ms_gemm_big", 0) > 0 and stp["ms_gemm"] > stp["ms_gemm_big"]:
        big_tf = stp["gemm_flops_big"] / (stp["ms_gemm_big"] * 1e-3) / 1e12
        sm_tf = (stp["gemm_flops"] - stp["gemm_flops_big"]) / ((stp["ms_gemm"] - stp["ms_gemm_big"]) * 1e-3) / 1e12
        split = {"trailing_updates_K_outer_block": {"tflops": big_tf, "frac": big_tf / peak, "ms": stp["ms_gemm_big"],
                                                    "flops_share": stp["gemm_flops_big"] / max(stp["gemm_flops"], 1.0)},
                 "in_block_updates_K_panel_width": {"tflops": sm_tf, "frac": sm_tf / peak, "ms": stp["ms_gemm"] - stp["ms_gemm_big"],
                                                    "note": "launches of a few CTAs, bounded by pipeline fill + C-tile epilogue; they overlap the trailing "
                                                            "updates on the look-ahead stream outside profile mode"}}
    return {"split": split,"kernel": "k_gemm (FP64 DMMA m8n8k4 Schur/trailing update)", "bound": "tensor", "achieved": gemm_tf,
            "peak": peak, "unit": "TFLOP/s", "frac": gemm_tf / peak,
            "traffic": ncu_traffic("k_gemm") if DEFAULT_WORKLOAD else None, "traffic_unit": "bytes per launch (ncu, all launches averaged)",
            "peak_source": peak_src,
            "launches": stp["gemm_launches"], "flops_per_launch": stp["gemm_flops"] / max(stp["gemm_launches"], 1),
            "ms_per_launch": stp["ms_gemm"] / max(stp["gemm_launches"], 1),
            "phase_ms": {k: stp[k] for k in ("ms_assemble", "ms_small", "ms_panel", "ms_trsm", "ms_gemm", "ms_solve_prep")}}


def hbm_rooflines(stp, solve_ms):
    """The two HBM-bound phases north_star names: extend-add (2·esz·Σnb² bytes) and the tree solve
    (esz·Σ(ni²+2·ni·nb) bytes per right-hand side), against the measured copy bandwidth of MEASURED_PEAKS.json."""
    peak, src = hbm_peak()
    out = []
    if stp.get("ms_extend_add", 0) > 0:
        a = stp["extadd_bytes"] / (stp["ms_extend_add"] * 1e-3) / 1e9
        out.append({"kernel": "k_extend_add", "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                    "traffic": ncu_traffic("k_extend_add", False) if DEFAULT_WORKLOAD else None, "peak_source": src, "bytes": stp["extadd_bytes"], "ms": stp["ms_extend_add"]})
    if solve_ms and solve_ms > 0:
        a = stp["solve_bytes"] / (solve_ms * 1e-3) / 1e9
        names = ("k_sv_small_fwd", "k_sv_small_bwd", "k_sv_big_fwd", "k_sv_big_bwd", "k_gemv_rect", "k_sv_tri_fwd", "k_sv_tri_bwd")
        out.append({"kernel": "tree solve (k_sv_small_*, k_sv_tri_*, k_gemv_rect), one right-hand side", "bound": "hbm", "achieved": a,
                    "peak": peak, "unit": "GB/s", "frac": a / peak,
                    "traffic": (sum(ncu_traffic(k, False) or 0 for k in names) or None) if DEFAULT_WORKLOAD else None,
                    "peak_source": src, "bytes": stp["solve_bytes"], "ms": solve_ms})
    return out


def h2d_bytes(Ap, nd, nd_loc, b):
    tree_bytes = 8 * (2 * nd.nnodes + 4 * (nd.nnodes + 1) + len(nd.int_idx) + len(nd.bnd_idx) + len(nd_loc.iloc_idx) + len(nd_loc.bloc_idx))
    a_bytes = Ap.indptr.size * Ap.indptr.itemsize + Ap.indices.size * Ap.indices.itemsize + Ap.data.nbytes
    return int(a_bytes + tree_bytes + b.nbytes)


def _ff(nd):
    return front_flops(np.diff(nd.int_ptr), np.diff(nd.bnd_ptr), False)


def run_distributed(args, hs, torch, world, rank, local_rank, Ap, nd, nd_loc, b, flops, tdt):
    """N > 1: one disjoint bottom subtree per GPU; the fronts above the cut are owned by the rank of their left child and
    receive the other child's Schur block point-to-point over NCCL (hierarchicalsolvers.jl_b200/parallel.py).  Strong scaling."""
    import torch.distributed as dist
    from hsolve_b200.parallel import CudaEngine, DistributedFactor, gmres_replicated
    lib = hs._lib.lib
    eng = CudaEngine(local_rank)
    t0 = time.perf_counter()
    DF = DistributedFactor(Ap, nd, nd_loc, engine=eng, swlevel=0)
    t_first = time.perf_counter() - t0
    A_t = eng.matvec(DF.h_sub)   # v -> A·v with the matrix this rank's subtree factorization holds in HBM (hs_spmv)
    b_dev = eng.to_device(b)
    out = {}

    def step():
        DF.refactor()
        x, res, conv = gmres_replicated(A_t, b_dev, DF.ldiv_device, 1e-9, 30, 30)
        out["x"], out["iters"] = x, len(res)

    for _ in range(args.warmup):
        step()
    lc0 = C.c_int64(); lib.hs_launch_count(eng.ctx, C.byref(lc0))
    sampler = ClockSampler(local_rank)
    sampler.start()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fac_ms = []
    for _ in range(args.steps):
        tf0 = time.perf_counter()
        DF.refactor()
        torch.cuda.synchronize()
        fac_ms.append((time.perf_counter() - tf0) * 1e3)
        x, res, conv = gmres_replicated(A_t, b_dev, DF.ldiv_device, 1e-9, 30, 30)
        out["x"], out["iters"] = x, len(res)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    clocks = sampler.stop()
    lc1 = C.c_int64(); lib.hs_launch_count(eng.ctx, C.byref(lc1))
    t = torch.tensor([e0.elapsed_time(e1) / args.steps, float(np.mean(fac_ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    xh = out["x"].cpu().numpy()
    resid = float(np.linalg.norm(Ap @ xh - b) / np.linalg.norm(b))
    # roofline: the DMMA update of this rank's subtree and of the fronts above the cut it owns, event-timed
    hs._lib.check(lib.hs_set_profile(eng.ctx, 1))
    DF.refactor()
    hs._lib.check(lib.hs_set_profile(eng.ctx, 0))
    sts = [eng.stats(h) for h in DF.local_handles()]
    stp = dict(sts[0])
    for k in ("gemm_flops", "ms_gemm", "gemm_launches", "ms_assemble", "ms_small", "ms_panel", "ms_trsm", "ms_solve_prep",
              "solve_bytes", "front_bytes", "gemm_flops_big", "ms_gemm_big"):
        stp[k] = sum(st[k] for st in sts)
    peak, peak_src = fp64_peak(b.dtype == np.complex128)
    e2e = None
    par = {"scheme": "subtree-per-GPU; fronts above the cut owned by the left child's rank, Schur blocks sent point-to-point (NCCL); GMRES replicated", "top_mode": DF.mode,
           "cut_nodes": [int(c) for c in DF.part.cut], "schur_bytes_exchanged": DF.schur_bytes,
           "top_flops_share": float(1.0 - sum(_ff(sn) for sn in DF.part.sub_nd) / float(np.sum(DF.part.work)))}
    DF.free()                      # the end-to-end run builds its own factorization: release this one first
    del out["x"]
    torch.cuda.empty_cache()
    if not args.no_e2e:
        dist.barrier(); torch.cuda.synchronize()
        passes = []
        for rep in range(2):          # one untimed warm-up pass (as at N = 1), one timed
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            DF2 = DistributedFactor(Ap, nd, nd_loc, engine=eng, swlevel=0)
            t1 = time.perf_counter()
            bh = eng.to_device(b)
            x2, _, _ = gmres_replicated(eng.matvec(DF2.h_sub), bh, DF2.ldiv_device, 1e-9, 30, 30)
            _ = x2.cpu()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            passes.append((t2 - t0, t1 - t0, t2 - t1))
            DF2.free(); del DF2, x2, bh
            torch.cuda.empty_cache()
        te = torch.tensor(list(passes[-1]), device="cuda", dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": float(te[0].item()), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(Ap, nd, nd_loc, b), "d2h_bytes_per_step": int(b.nbytes),
               "construct_s": float(te[1].item()), "gmres_s": float(te[2].item()), "warmup_pass_s": float(passes[0][0]),
               "note": "DistributedFactor(host CSC: partition, plan, upload, factor) + replicated GMRES, per rank; max over ranks; one untimed warm-up pass before"}
    return {"ms_step": float(t[0].item()), "fac_ms": float(t[1].item()), "iters": out["iters"], "resid": resid,
            "roofline": gemm_roofline(stp, peak, peak_src), "e2e": e2e, "clocks": clocks,
            "launches": int(lc1.value - lc0.value) // max(args.steps, 1), "stats": stp, "t_first": t_first,
            "parallelism": par}


def resident_single(args, hs, torch, ctx, stream, local_rank, Ap, nd, nd_loc, b, cx, steps, warmup, with_e2e, with_profile=True):
    """One GPU.  Returns a dict with the resident-input step time (`value`), the event-timed kernel phases, the tree
    solve roofline and (optionally) the end-to-end time through the public API with host buffers."""
    lib = hs._lib.lib
    dtype = np.complex128 if cx else np.float64
    tdt = torch.complex128 if cx else torch.float64
    n = Ap.shape[0]
    t0 = time.perf_counter()
    F = hs.factor(Ap, nd, nd_loc, swlevel=0, device=local_rank)      # builds the plan, leaves A resident in HBM
    t_first = time.perf_counter() - t0
    h = F._hd.h
    nz_dev = torch.from_numpy(np.ascontiguousarray(Ap.data, dtype=dtype)).to("cuda")
    b_dev = torch.from_numpy(b).to("cuda")
    x_dev = torch.zeros(n, dtype=tdt, device="cuda")
    res = np.zeros(30, dtype=np.float64)
    nit, conv = C.c_int64(), C.c_int32()

    def step_resident():
        hs._lib.check(lib.hs_refactor(h, C.c_void_p(nz_dev.data_ptr()), 1))
        hs._lib.check(lib.hs_gmres(ctx, hs._lib.HS_C64 if cx else hs._lib.HS_F64, n, None, None, None, 0, h,
                                   C.c_void_p(b_dev.data_ptr()), C.c_void_p(x_dev.data_ptr()), 1e-9, 30, 30,
                                   res.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nit), C.byref(conv), 1))

    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step_resident()
        lc0 = C.c_int64(); lib.hs_launch_count(ctx, C.byref(lc0))
        sampler = ClockSampler(local_rank)
        sampler.start()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fac_ms = []
        for _ in range(steps):
            step_resident()
            fac_ms.append(F.stats()["ms_factor_total"])
        e1.record(stream)
        torch.cuda.synchronize()
        clocks = sampler.stop()
        lc1 = C.c_int64(); lib.hs_launch_count(ctx, C.byref(lc1))
        ms_step = e0.elapsed_time(e1) / steps
    out = {"ms_step": ms_step, "fac_ms": float(np.mean(fac_ms)), "iters": int(nit.value), "clocks": clocks,
           "launches": int(lc1.value - lc0.value) // max(steps, 1), "t_first": t_first}
    xh = x_dev.cpu().numpy()
    out["resid"] = float(np.linalg.norm(Ap @ xh - b) / np.linalg.norm(b))
    peak, peak_src = fp64_peak(cx)
    if with_profile:
        # roofline of the dominant kernel (DMMA Schur/trailing update), event-timed per launch
        hs._lib.check(lib.hs_set_profile(ctx, 1))
        with torch.cuda.stream(stream):
            hs._lib.check(lib.hs_refactor(h, C.c_void_p(nz_dev.data_ptr()), 1))
        hs._lib.check(lib.hs_set_profile(ctx, 0))
        stp = F.stats()
        out["roofline"] = gemm_roofline(stp, peak, peak_src)
        hs.ldiv(F, b)          # one preconditioner application, event-timed inside the library
        out["roofline_hbm"] = hbm_rooflines(stp, F.stats()["ms_solve_total"])
        out["stats"] = stp
    else:
        out["stats"] = F.stats()
    del F, nz_dev, b_dev, x_dev
    torch.cuda.empty_cache()
    if with_e2e:
        ts, tf_, tg_ = [], [], []
        st2 = None
        for rep in range(1 + max(3, min(steps, 5))):     # one untimed warm-up pass (first-touch of host staging pages), then 3 to 5 timed
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            F2 = hs.factor(Ap, nd, nd_loc, swlevel=0, device=local_rank)
            t1 = time.perf_counter()
            x2, hist = hs.gmres(Ap, b, Pr=F2, reltol=1e-9, restart=30, maxiter=30, log=True, A_is_factored=True)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if rep > 0:
                ts.append(t2 - t0); tf_.append(t1 - t0); tg_.append(t2 - t1)
            st2 = F2.stats()
            del F2
        # median, every pass listed: a pass now and then catches a host hiccup (a 15 GB cudaFree/cudaMalloc cycle, page
        # reclaim) that triples its plan build — the mean of three would report that, not the path
        mid = int(np.argsort(ts)[len(ts) // 2])
        out["e2e"] = {"value": float(ts[mid]), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(Ap, nd, nd_loc, b),
                      "d2h_bytes_per_step": int(b.nbytes + 8 * 31), "note": "hs.factor(host CSC) + hs.gmres(host b) incl. plan build; median of the timed passes (all listed)",
                      "passes_s": [float(v) for v in ts],
                      "factor_call_s": float(tf_[mid]), "gmres_call_s": float(tg_[mid]),
                      "inside_factor_ms": {k: st2[k] for k in ("ms_analyze", "ms_h2d", "ms_factor_total")}}
    return out


def main():
    args = parse_args()
    global DEFAULT_WORKLOAD
    DEFAULT_WORKLOAD = args.workload == "2d" and args.grid == 2048 and args.kind == "poisson" and args.nmax == 100
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cx = args.kind == "helmholtz"

    if args.impl == "reference":
        # rank 0 alone runs and prints; the product package is never imported in this arm
        if rank == 0:
            reference_arm(args)
        return

    import _pkg
    hs = _pkg.load()
    shape = shape_of(args)
    config = {"workload": workload_name(args, shape, args.kind), "n": int(np.prod(shape)),
              "l2": "inputs larger than L2 (fronts >> 126 MB), no flush", "gmres_update": GMRES_NOTE}

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = hs._lib.default_context(local_rank)
    stream = torch.cuda.Stream() if world == 1 else torch.cuda.current_stream()
    hs._lib.check(hs._lib.lib.hs_set_stream(ctx, C.c_void_p(stream.cuda_stream)))

    # ---- problem (untimed) ----------------------------------------------------------------------
    t0 = time.perf_counter()
    prob = hs.grid_problem(shape, args.kind, nmax=args.nmax)
    Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
    Ap.sort_indices()
    t_setup = time.perf_counter() - t0
    dtype = np.complex128 if cx else np.float64
    b = np.ascontiguousarray(prob.b, dtype=dtype)
    flops = factor_flops(prob.elim_tree, cx)
    tdt = torch.complex128 if cx else torch.float64
    peak, peak_src = fp64_peak(cx)
    roofline_hbm = None
    if world > 1:
        r = run_distributed(args, hs, torch, world, rank, local_rank, Ap, nd, nd_loc, b, flops, tdt)
        config["parallelism"] = r.pop("parallelism")
    else:
        r = resident_single(args, hs, torch, ctx, stream, local_rank, Ap, nd, nd_loc, b, cx, args.steps, args.warmup,
                            with_e2e=not args.no_e2e)
        roofline_hbm = r.get("roofline_hbm")
    ms_step, fac_ms_mean, gm_iters, resid = r["ms_step"], r["fac_ms"], r["iters"], r["resid"]
    stp = r["stats"]
    # a wrong answer is not a benchmark result: the run fails loudly (N = 1 and N > 1 alike)
    if not (resid <= 1e-8):
        print(json.dumps({"error": "residual check failed", "residual": resid, "gmres_iters": gm_iters, "n_gpus": world}), flush=True)
        raise SystemExit(f"bench.py: relative residual {resid:.3e} > 1e-8 at n_gpus={world}")

    # ---- informational: the same workload with compressed upper fronts (rungmres.jl:39-style options) ----
    compressed = None
    if world == 1 and not args.no_compressed and args.workload == "2d":
        compressed = {}
        for tag, copts in (("lowrank_dense_schur", dict(swlevel=-2, swsize=480, atol=1e-5, rtol=1e-5, hss=False)),
                           ("hss_schur", dict(swlevel=-2, swsize=128, atol=1e-5, rtol=1e-5, leafsize=32, hss=True))):
            try:
                Fc = hs.factor(Ap, nd, nd_loc, device=local_rank, **copts)
                Fc.refactor(Ap)   # steady state: side buffers sized
                xc, hc = hs.gmres(Ap, b, Pr=Fc, reltol=1e-9, restart=30, maxiter=30, log=True, A_is_factored=True)
                sc = Fc.stats()
                compressed[tag] = {"opts": copts, "factor_ms": sc["ms_factor_total"], "maxrank": int(hs.maxrank(Fc)),
                                   "gmres_iters": hc.iters, "converged": bool(hc.isconverged), "apply_ms": sc["ms_solve_total"],
                                   "front_bytes": sc["front_bytes"], "lowrank_bytes": sc["lowrank_bytes"],
                                   "residual": float(np.linalg.norm(Ap @ xc - b) / np.linalg.norm(b)),
                                   "note": "not part of `value` (DESIGN.md §4a/§4b)"}
                del Fc
            except Exception as exc:  # informational only
                compressed[tag] = {"error": repr(exc)}
            torch.cuda.empty_cache()

    # ---- secondary: the complex Helmholtz 2048² workload (c64 / ZGEMM numbers, driver-observed) ----
    c64 = None
    if world == 1 and DEFAULT_WORKLOAD and not args.no_c64:
        try:
            prob_c = hs.grid_problem(shape, "helmholtz", nmax=args.nmax)
            Ac, ndc, ndc_loc, _ = hs.prepare(prob_c.A, prob_c.elim_tree)
            Ac.sort_indices()
            bc = np.ascontiguousarray(prob_c.b, dtype=np.complex128)
            rc = resident_single(args, hs, torch, ctx, stream, local_rank, Ac, ndc, ndc_loc, bc, True,
                                 max(1, min(args.steps, 5)), max(3, min(args.warmup, 3)), with_e2e=False)
            fl_c = factor_flops(prob_c.elim_tree, True)
            pk_c, _ = fp64_peak(True)
            c64 = {"workload": workload_name(args, shape, "helmholtz"), "value": rc["ms_step"] * 1e-3, "unit": UNIT,
                   "factor_ms": rc["fac_ms"], "solve_ms": rc["ms_step"] - rc["fac_ms"], "gmres_iters": rc["iters"], "residual": rc["resid"],
                   "factor_tflops": fl_c / (rc["fac_ms"] * 1e-3) / 1e12, "factor_frac_of_zgemm_peak": fl_c / (rc["fac_ms"] * 1e-3) / 1e12 / pk_c,
                   "roofline": rc.get("roofline"), "roofline_hbm": rc.get("roofline_hbm"), "clocks": rc["clocks"]}
            del prob_c, Ac
        except Exception as exc:
            c64 = {"error": repr(exc)}
        if not (c64.get("residual", 0.0) <= 1e-8):   # same gate as the main workload: a wrong answer is not a timing
            raise SystemExit(f"bench: complex Helmholtz block, relative residual {c64.get('residual')!r} > 1e-8")

    # ---- CPU baseline + like-for-like: the oracle port and this library on the SAME bounded sample grid ----
    cpu_baseline = like = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rcpu, cpu = cpu_sample(args, steps=3 if args.workload == "2d" else 1, warmup=0)
        sshape = tuple(rcpu["shape"])
        cpu_baseline = {"value": rcpu["seconds"], "unit": UNIT, "cores": rcpu["threads"], "kind": "port", "host_cores": rcpu["cores"],
                        "sample": (f"oracle port (NumPy/SciPy restatement of the reference), LAPACK threads={rcpu['threads']} of {rcpu['cores']} cores "
                                   f"(chosen by measurement: {rcpu['thread_trials_s']}); MEASURED on the sample grid {'x'.join(map(str, sshape))} "
                                   f"(N={rcpu['n']}, {rcpu['flops'] / 1e9:.2f} GFLOP): {rcpu['seconds']:.3f} s = {rcpu['factor_s']:.3f} factor + "
                                   f"{rcpu['solve_s']:.3f} GMRES; not scaled to the full workload - see like_for_like"),
                        "splu_same_grid": rcpu.get("splu")}
        try:
            ps = hs.grid_problem(sshape, args.kind, nmax=args.nmax)
            As, nds, nds_loc, _ = hs.prepare(ps.A, ps.elim_tree)
            As.sort_indices()
            bs = np.ascontiguousarray(ps.b, dtype=dtype)
            rs = resident_single(args, hs, torch, ctx, stream, local_rank, As, nds, nds_loc, bs, cx, 5, 3, with_e2e=True, with_profile=False)
            like = {"grid": list(sshape), "n": rcpu["n"], "cpu_port_s": rcpu["seconds"], "cpu_threads": rcpu["threads"],
                    "cpu_splu_s": (rcpu.get("splu") or {}).get("seconds"),
                    "gpu_value_s": rs["ms_step"] * 1e-3, "gpu_e2e_s": rs["e2e"]["value"], "gpu_residual": rs["resid"],
                    "cpu_residual": rcpu["residual"], "gmres_iters": {"cpu": rcpu["gmres_iters"], "gpu": rs["iters"]},
                    "note": "same grid, same run, both measured; at this size the GPU path is launch-latency-bound"}
        except Exception as exc:
            like = {"error": repr(exc)}
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": ms_step * 1e-3, "unit": UNIT, "residual": resid, "gmres_iters": gm_iters,
            "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "c64" if cx else "f64",
            "data": "synthetic", "config": config,
            "factor_ms": fac_ms_mean, "solve_ms": ms_step - fac_ms_mean,
            "factor_tflops": flops / (fac_ms_mean * 1e-3) / 1e12, "factor_flops": flops,
            "factor_frac_of_fp64_peak": flops / (fac_ms_mean * 1e-3) / 1e12 / peak,
            "solve_bytes_per_rhs": stp["solve_bytes"], "front_bytes": stp["front_bytes"],
            "setup_s": {"generate+symfact": t_setup, "first_factor_incl_plan": r["t_first"]},
            "e2e": r.get("e2e"), "gpu_launches": r["launches"], "roofline": r["roofline"], "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu_baseline, "like_for_like": like, "clocks": r["clocks"]}
    if c64 is not None:
        line["c64"] = c64
    if compressed is not None:
        line["compressed_variant"] = compressed
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
