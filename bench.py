#!/usr/bin/env python
"""Benchmark of the hot path: multifrontal nested-dissection factor + preconditioned GMRES solve.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port on the host cores

A "step" is one pass of the hot path over the workload: numeric factorization of A (factor, factorization.jl:5-11)
followed by the driver's solve, GMRES(30) to reltol 1e-9 with the factorization as right preconditioner
(test/rungmres.jl:47).  Workload = BASELINE.json configs[3], the largest configuration that fits one B200:
synthetic 2D 5-point Laplacian on a 2048×2048 grid, geometric nested dissection with nmax = 100, float64.

Prints ONE JSON line (contract in the task statement): metric/value/unit, e2e, roofline, cpu_baseline, clocks, …
"""
from __future__ import annotations

import argparse
import ctypes as C
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=2048, help="grid side of the synthetic 2D problem")
    ap.add_argument("--kind", default="poisson", choices=["poisson", "helmholtz"])
    ap.add_argument("--nmax", type=int, default=100)
    ap.add_argument("--cpu-grid", type=int, default=384, help="grid side of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-compressed", action="store_true", help="skip the informational compressed-fronts variant")
    return ap.parse_args()


def factor_flops(et, complex_):
    """Σ ⅔ni³ + 2ni²nb + 2ni·nb² over the fronts, ×4 for complex."""
    ni = et.ninter().astype(np.float64)
    nb = et.nbound().astype(np.float64)
    f = (2.0 / 3.0) * ni ** 3 + 2.0 * ni ** 2 * nb + 2.0 * ni * nb ** 2
    return float(f.sum()) * (4.0 if complex_ else 1.0)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference algorithm; the reference itself is Julia and cannot run here)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_run(hs, grid, kind, nmax, steps=1, warmup=0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hs_oracle as orc
    prob = hs.grid_problem((grid, grid), kind, nmax=nmax)
    Ao, nd, nd_loc, _ = orc.prepare(prob.A, prob.elim_tree)
    times, iters = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        F = orc.factor(Ao, nd, nd_loc)
        t1 = time.perf_counter()
        x, res, conv = orc.gmres(Ao, prob.b, Pr=lambda v: orc.ldiv(F, v), reltol=1e-9, restart=30, maxiter=30)
        t2 = time.perf_counter()
        if it >= warmup:
            times.append((t2 - t0, t1 - t0, t2 - t1))
        iters = len(res)
    tot = float(np.mean([t[0] for t in times]))
    return {"seconds": tot, "factor_s": float(np.mean([t[1] for t in times])), "solve_s": float(np.mean([t[2] for t in times])),
            "gmres_iters": iters, "flops": factor_flops(prob.elim_tree, kind == "helmholtz"), "n": Ao.shape[0]}


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


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


def ncu_traffic(kernel, per_launch=True):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu pass over the default workload
    (profiles/r01_traffic_2048.json, tools/profile_run.py 2048); None for any other workload."""
    try:
        k = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_2048.json")))["kernels"][kernel]
        return k["dram_bytes_per_launch"] if per_launch else k["dram_bytes_total"]
    except Exception:
        return None


DEFAULT_WORKLOAD = True


def gemm_roofline(stp, peak, peak_src):
    gemm_tf = stp["gemm_flops"] / (stp["ms_gemm"] * 1e-3) / 1e12 if stp["ms_gemm"] > 0 else 0.0
    return {"kernel": "k_gemm (FP64 DMMA m8n8k4 Schur/trailing update)", "bound": "tensor", "achieved": gemm_tf,
            "peak": peak, "unit": "TFLOP/s", "frac": gemm_tf / peak,
            "traffic": ncu_traffic("k_gemm") if DEFAULT_WORKLOAD else None, "traffic_unit": "bytes per launch (ncu, all launches averaged)",
            "peak_source": peak_src,
            "launches": stp["gemm_launches"], "flops_per_launch": stp["gemm_flops"] / max(stp["gemm_launches"], 1),
            "ms_per_launch": stp["ms_gemm"] / max(stp["gemm_launches"], 1),
            "phase_ms": {k: stp[k] for k in ("ms_assemble", "ms_small", "ms_panel", "ms_trsm", "ms_gemm", "ms_solve_prep")}}


def hbm_rooflines(stp, solve_ms):
    """The two HBM-bound phases north_star names: extend-add (2·esz·Σnb² bytes) and the tree solve
    (esz·Σ(ni²+2·ni·nb) bytes per right-hand side), against the measured copy bandwidth of MEASURED_PEAKS.json."""
    peak = 6456.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        src = "fallback 6456.2 GB/s (earlier MEASURED_PEAKS.json of this pool)"
    out = []
    if stp.get("ms_extend_add", 0) > 0:
        a = stp["extadd_bytes"] / (stp["ms_extend_add"] * 1e-3) / 1e9
        out.append({"kernel": "k_extend_add", "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                    "traffic": ncu_traffic("k_extend_add", False) if DEFAULT_WORKLOAD else None, "peak_source": src, "bytes": stp["extadd_bytes"], "ms": stp["ms_extend_add"]})
    if solve_ms and solve_ms > 0:
        a = stp["solve_bytes"] / (solve_ms * 1e-3) / 1e9
        out.append({"kernel": "tree solve (k_sv_small_*, k_sv_big_*, k_gemv_rect), one right-hand side", "bound": "hbm", "achieved": a,
                    "peak": peak, "unit": "GB/s", "frac": a / peak,
                    "traffic": (sum(ncu_traffic(k, False) or 0 for k in ("k_sv_small_fwd", "k_sv_small_bwd", "k_sv_big_fwd", "k_sv_big_bwd", "k_gemv_rect")) or None)
                    if DEFAULT_WORKLOAD else None, "peak_source": src, "bytes": stp["solve_bytes"],
                    "ms": solve_ms})
    return out


def h2d_bytes(Ap, nd, nd_loc, b):
    tree_bytes = 8 * (2 * nd.nnodes + 4 * (nd.nnodes + 1) + len(nd.int_idx) + len(nd.bnd_idx) + len(nd_loc.iloc_idx) + len(nd_loc.bloc_idx))
    a_bytes = Ap.indptr.size * 8 + Ap.indices.size * 8 + Ap.data.nbytes
    return int(2 * a_bytes + tree_bytes + b.nbytes)


def run_distributed(args, hs, torch, world, rank, local_rank, Ap, nd, nd_loc, b, flops, tdt):
    """N > 1: one disjoint bottom subtree per GPU, Schur blocks of the subtree roots all-gathered over NCCL, the fronts
    above the cut and the GMRES iteration replicated (hierarchicalsolvers.jl_b200/parallel.py).  Strong scaling."""
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
    # roofline: the DMMA update of this rank's subtree and of the replicated top, event-timed
    hs._lib.check(lib.hs_set_profile(eng.ctx, 1))
    DF.refactor()
    hs._lib.check(lib.hs_set_profile(eng.ctx, 0))
    sts = [eng.stats(h) for h in DF.local_handles()]
    stp = dict(sts[0])
    for k in ("gemm_flops", "ms_gemm", "gemm_launches", "ms_assemble", "ms_small", "ms_panel", "ms_trsm", "ms_solve_prep",
              "solve_bytes", "front_bytes"):
        stp[k] = sum(st[k] for st in sts)
    peak, peak_src = fp64_peak(b.dtype == np.complex128)
    e2e = None
    if not args.no_e2e:
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        DF2 = DistributedFactor(Ap, nd, nd_loc, engine=eng, swlevel=0)
        bh = eng.to_device(b)
        x2, _, _ = gmres_replicated(eng.matvec(DF2.h_sub), bh, DF2.ldiv_device, 1e-9, 30, 30)
        _ = x2.cpu()
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": float(te.item()), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(Ap, nd, nd_loc, b), "d2h_bytes_per_step": int(b.nbytes),
               "note": "DistributedFactor(host CSC) + replicated GMRES, per rank; max over ranks"}
    return {"ms_step": float(t[0].item()), "fac_ms": float(t[1].item()), "iters": out["iters"], "resid": resid,
            "roofline": gemm_roofline(stp, peak, peak_src), "e2e": e2e, "clocks": clocks,
            "launches": int(lc1.value - lc0.value) // max(args.steps, 1), "stats": stp, "t_first": t_first,
            "parallelism": {"scheme": "subtree-per-GPU; fronts above the cut owned by the left child's rank, Schur blocks sent point-to-point (NCCL); GMRES replicated", "top_mode": DF.mode,
                            "cut_nodes": [int(c) for c in DF.part.cut], "schur_bytes_allgathered": DF.schur_bytes,
                            "top_flops_share": float(1.0 - sum(float(np.sum(_ff(sn))) for sn in DF.part.sub_nd) / float(np.sum(DF.part.work)))}}


def _ff(nd):
    ni = np.diff(nd.int_ptr).astype(np.float64)
    nb = np.diff(nd.bnd_ptr).astype(np.float64)
    return (2.0 / 3.0) * ni ** 3 + 2.0 * ni ** 2 * nb + 2.0 * ni * nb ** 2


def main():
    args = parse_args()
    global DEFAULT_WORKLOAD
    DEFAULT_WORKLOAD = args.grid == 2048 and args.kind == "poisson" and args.nmax == 100
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import _pkg
    hs = _pkg.load()
    cx = args.kind == "helmholtz"
    workload = (f"synthetic 2D 5-point {'Helmholtz (complex, 10 ppw)' if cx else 'Laplacian'} {args.grid}x{args.grid}, "
                f"geometric nested dissection nmax={args.nmax}, uncompressed (swlevel=0), GMRES(30) reltol 1e-9")
    config = {"workload": workload, "n": args.grid * args.grid, "l2": "inputs larger than L2 (fronts ≫ 126 MB), no flush"}

    if args.impl == "reference":
        # the reference is Julia + un-vendored packages and cannot run here: this arm times the oracle port of its
        # algorithm on the host cores, on a bounded sample of the workload
        if rank != 0:
            return
        g = min(args.cpu_grid, args.grid)
        r = cpu_oracle_run(hs, g, args.kind, args.nmax, steps=max(1, min(args.steps, 2)), warmup=0)
        cores = host_threads()
        full = factor_flops(hs.grid_elimtree((args.grid, args.grid), args.nmax), cx) if g != args.grid else r["flops"]
        scale = full / r["flops"]
        sample = (f"oracle port (NumPy/SciPy restatement of the reference — the Julia reference cannot run here; LAPACK threads={cores}) "
                  f"on a {g}x{g} grid of the same generator: {r['seconds']:.2f} s/step measured ({r['factor_s']:.2f} factor + "
                  f"{r['solve_s']:.2f} GMRES, {r['flops'] / 1e9:.2f} GFLOP); value = measured × flop ratio {scale:.1f} to the "
                  f"{args.grid}x{args.grid} workload")
        val = r["seconds"] * scale
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": val * 1e3, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "c64" if cx else "f64", "data": "synthetic",
                "config": config, "gmres_iters": r["gmres_iters"], "measured_sample_s": r["seconds"],
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

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
    prob = hs.grid_problem((args.grid, args.grid), args.kind, nmax=args.nmax)
    Ap, nd, nd_loc, perm = hs.prepare(prob.A, prob.elim_tree)
    Ap.sort_indices()
    t_setup = time.perf_counter() - t0
    dtype = np.complex128 if cx else np.float64
    n = Ap.shape[0]
    b = np.ascontiguousarray(prob.b, dtype=dtype)
    flops = factor_flops(prob.elim_tree, cx)

    lib = hs._lib.lib
    tdt = torch.complex128 if cx else torch.float64
    peak, peak_src = fp64_peak(cx)
    roofline_hbm = None
    if world > 1:
        line_extra = run_distributed(args, hs, torch, world, rank, local_rank, Ap, nd, nd_loc, b, flops, tdt)
        ms_step, fac_ms_mean, gm_iters, resid = (line_extra.pop(k) for k in ("ms_step", "fac_ms", "iters", "resid"))
        roofline = line_extra.pop("roofline")
        e2e = line_extra.pop("e2e")
        clocks = line_extra.pop("clocks")
        launches = line_extra.pop("launches")
        stp = line_extra.pop("stats")
        t_first = line_extra.pop("t_first")
        config["parallelism"] = line_extra.pop("parallelism")
    else:
        # first factorization builds the plan and leaves A resident in HBM
        t0 = time.perf_counter()
        F = hs.factor(Ap, nd, nd_loc, swlevel=0, device=local_rank)
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
            for _ in range(args.warmup):
                step_resident()
            lc0 = C.c_int64(); lib.hs_launch_count(ctx, C.byref(lc0))
            sampler = ClockSampler(local_rank)
            sampler.start()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fac_ms = []
            for _ in range(args.steps):
                step_resident()
                fac_ms.append(F.stats()["ms_factor_total"])
            e1.record(stream)
            torch.cuda.synchronize()
            clocks = sampler.stop()
            lc1 = C.c_int64(); lib.hs_launch_count(ctx, C.byref(lc1))
            ms_step = e0.elapsed_time(e1) / args.steps
        launches = int(lc1.value - lc0.value) // max(args.steps, 1)
        gm_iters = int(nit.value)
        xh = x_dev.cpu().numpy()
        resid = float(np.linalg.norm(Ap @ xh - b) / np.linalg.norm(b))
        fac_ms_mean = float(np.mean(fac_ms))
        # ---- roofline of the dominant kernel (DMMA Schur/trailing update), event-timed per launch ----
        hs._lib.check(lib.hs_set_profile(ctx, 1))
        with torch.cuda.stream(stream):
            hs._lib.check(lib.hs_refactor(h, C.c_void_p(nz_dev.data_ptr()), 1))
        hs._lib.check(lib.hs_set_profile(ctx, 0))
        stp = F.stats()
        roofline = gemm_roofline(stp, peak, peak_src)
        xs = hs.ldiv(F, b)          # one preconditioner application, event-timed inside the library
        roofline_hbm = hbm_rooflines(stp, F.stats()["ms_solve_total"])
        # ---- end to end through the public API with host buffers ------------------------------------
        e2e = None
        if not args.no_e2e:
            del F
            ts, tf_, tg_ = [], [], []
            for _ in range(max(1, min(args.steps, 2))):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                F2 = hs.factor(Ap, nd, nd_loc, swlevel=0, device=local_rank)
                t1 = time.perf_counter()
                x2, hist = hs.gmres(Ap, b, Pr=F2, reltol=1e-9, restart=30, maxiter=30, log=True)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                ts.append(t2 - t0); tf_.append(t1 - t0); tg_.append(t2 - t1)
                st2 = F2.stats()
                del F2
            e2e = {"value": float(np.mean(ts)), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(Ap, nd, nd_loc, b),
                   "d2h_bytes_per_step": int(b.nbytes + 8 * 31), "note": "hs.factor(host CSC) + hs.gmres(host b) incl. plan build",
                   "factor_call_s": float(np.mean(tf_)), "gmres_call_s": float(np.mean(tg_)),
                   "inside_factor_ms": {k: st2[k] for k in ("ms_analyze", "ms_h2d", "ms_factor_total")}}

    # ---- informational: the same workload with compressed upper fronts (rungmres.jl:39: swlevel=-2, swsize=480) ----
    compressed = None
    if world == 1 and not args.no_compressed:
        try:
            copts = dict(swlevel=-2, swsize=480, atol=1e-5, rtol=1e-5)
            Fc = hs.factor(Ap, nd, nd_loc, device=local_rank, **copts)
            Fc.refactor(Ap)   # steady state: side buffers sized
            xc, hc = hs.gmres(Ap, b, Pr=Fc, reltol=1e-9, restart=30, maxiter=30, log=True)
            sc = Fc.stats()
            compressed = {"opts": copts, "factor_ms": sc["ms_factor_total"], "maxrank": int(hs.maxrank(Fc)),
                          "gmres_iters": hc.iters, "converged": bool(hc.isconverged), "apply_ms": sc["ms_solve_total"],
                          "residual": float(np.linalg.norm(Ap @ xc - b) / np.linalg.norm(b)),
                          "note": "not part of `value`; low-rank Gauss transforms, dense Schur complements (DESIGN.md §4a)"}
            del Fc
        except Exception as exc:  # informational only
            compressed = {"error": repr(exc)}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        g = min(args.cpu_grid, args.grid)
        r = cpu_oracle_run(hs, g, args.kind, args.nmax)
        cores = host_threads()
        cpu_baseline = {"value": r["seconds"] * flops / r["flops"], "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": (f"oracle port (NumPy/SciPy restatement of the reference, LAPACK threads={cores}) on a {g}x{g} grid: "
                                   f"{r['seconds']:.2f} s measured ({r['factor_s']:.2f} factor + {r['solve_s']:.2f} GMRES, {r['flops'] / 1e9:.2f} GFLOP); "
                                   f"value = measured × flop ratio {flops / r['flops']:.1f} to the {args.grid}x{args.grid} workload"),
                        "measured_sample_s": r["seconds"]}
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": ms_step * 1e-3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "c64" if cx else "f64",
            "data": "synthetic", "config": config,
            "factor_ms": fac_ms_mean, "solve_ms": ms_step - fac_ms_mean, "gmres_iters": gm_iters, "residual": resid,
            "factor_tflops": flops / (fac_ms_mean * 1e-3) / 1e12, "factor_flops": flops,
            "factor_frac_of_fp64_peak": flops / (fac_ms_mean * 1e-3) / 1e12 / peak,
            "solve_bytes_per_rhs": stp["solve_bytes"], "front_bytes": stp["front_bytes"],
            "setup_s": {"generate+symfact": t_setup, "first_factor_incl_plan": t_first},
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu_baseline,
            "clocks": clocks}
    if compressed is not None:
        line["compressed_variant"] = compressed
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
