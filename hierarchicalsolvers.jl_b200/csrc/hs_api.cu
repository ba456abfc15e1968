// libhsolve_cuda: C ABI, level-batched plan construction and kernel orchestration.
// Reference call sites each entry point replaces are listed in include/hsolve_cuda.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <exception>
#include <mutex>
#include <numeric>
#include <thread>

#include "hs_fac.cuh"
#include "hs_internal.h"
#include "hs_kernels.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

int32_t hs_fail(int32_t code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

extern "C" const char* hs_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t hs_version(void) { return HS_VERSION; }

extern "C" int32_t hs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int32_t hs_create(hs_ctx** out, int32_t device) {
  HS_TRY_BEGIN
  if (!out) return hs_fail(HS_EARG, "hs_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return hs_fail(HS_ECUDA, "hs_create: no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return hs_fail(HS_EARG, "hs_create: device out of range");
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return hs_fail(HS_ECUDA, std::string("hs_create: built for sm_100a, found ") + prop.name);
  auto* c = new hs_ctx();
  c->device = device;
  CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  {
    int lo = 0, hi = 0;  // numerically lower = higher priority
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_OK(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, hi));
    if (!getenv("HS_NO_PREP_OVERLAP")) CUDA_OK(cudaStreamCreateWithPriority(&c->prep_stream, cudaStreamNonBlocking, lo));
    if (!getenv("HS_NO_BELOW_STREAM")) CUDA_OK(cudaStreamCreateWithPriority(&c->below_stream, cudaStreamNonBlocking, lo));
  }
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_pan, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_urow, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_below, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_p0, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_p1, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_c2, cudaEventDisableTiming));
  if (getenv("HS_LOOKAHEAD")) c->lookahead_max_fronts = atoi(getenv("HS_LOOKAHEAD"));
  c->profile = getenv("HS_PROFILE") != nullptr;
  // opt in to 16-CTA clusters for the tall panels, and to >48 KB dynamic shared memory for the DMMA tiles
  hs_panel_setup_f64();
  hs_panel_setup_c64();
  hs_solve_setup();
  hs_comp_setup();
  hs_hss_setup();
  c->max_cluster = getenv("HS_MAX_CLUSTER") ? atoi(getenv("HS_MAX_CLUSTER")) : 16;
  if (getenv("HS_OUTER_BLOCK")) c->outer_block = std::max(1, atoi(getenv("HS_OUTER_BLOCK")));
  CUDA_OK(cudaFuncSetAttribute(k_gemm<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<double>()));
  CUDA_OK(cudaFuncSetAttribute(k_gemm<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<cplx>()));
  *out = c;
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_set_stream(hs_ctx* ctx, void* s) {
  if (!ctx) return hs_fail(HS_EARG, "hs_set_stream: null context");
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return HS_OK;
}

extern "C" int32_t hs_set_profile(hs_ctx* ctx, int32_t on) {
  if (!ctx) return hs_fail(HS_EARG, "hs_set_profile: null context");
  ctx->profile = on != 0;
  return HS_OK;
}

extern "C" int32_t hs_launch_count(hs_ctx* ctx, int64_t* count) {
  if (!ctx || !count) return hs_fail(HS_EARG, "hs_launch_count: null argument");
  *count = ctx->launches;
  return HS_OK;
}

extern "C" int32_t hs_destroy(hs_ctx* ctx) {
  if (!ctx) return HS_OK;
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaFree(ctx->gm_buf);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->prep_stream) cudaStreamDestroy(ctx->prep_stream);
  if (ctx->below_stream) cudaStreamDestroy(ctx->below_stream);
  if (ctx->ev_pan) cudaEventDestroy(ctx->ev_pan);
  if (ctx->ev_urow) cudaEventDestroy(ctx->ev_urow);
  if (ctx->ev_below) cudaEventDestroy(ctx->ev_below);
  if (ctx->ev_p0) cudaEventDestroy(ctx->ev_p0);
  if (ctx->ev_p1) cudaEventDestroy(ctx->ev_p1);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->ev_c2) cudaEventDestroy(ctx->ev_c2);
  delete ctx;
  return HS_OK;
}

// accumulate the time of a phase only when HS_PROFILE is set (it serialises the stream)
struct PhaseTimer {
  hs_fac* f;
  double* acc;
  cudaEvent_t a = nullptr, b = nullptr;
  PhaseTimer(hs_fac* f_, double* acc_) : f(f_), acc(acc_) {
    if (f->ctx->profile) {
      cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a, f->ctx->stream);
    }
  }
  ~PhaseTimer() {
    if (f->ctx->profile) {
      cudaEventRecord(b, f->ctx->stream);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      *acc += ms;
      cudaEventDestroy(a); cudaEventDestroy(b);
    }
  }
};

// solve preparation (in-place inversion of the diagonal blocks of L11/U11): nothing in the rest of the factorization
// reads a finished pivot block, so it runs on a low-priority side stream; numeric() joins it at the end.  Thin fronts
// need it at once (hs_comp_schur back-substitutes with the inverted blocks).
static void solve_prep(hs_fac* f, const Level& L) {
  hs_ctx* c = f->ctx;
  if (c->profile || L.thin >= 0 || !c->prep_stream) {
    PhaseTimer t(f, &f->stats.ms_solve_prep);
    hs_solve_prep(f, L, c->stream);
    return;
  }
  CUDA_OK(cudaEventRecord(c->ev_p0, c->stream));
  CUDA_OK(cudaStreamWaitEvent(c->prep_stream, c->ev_p0, 0));
  hs_solve_prep(f, L, c->prep_stream);
  c->prep_pending = true;
}

template <typename T> struct PanelW;  // widest register tile per scalar type
template <> struct PanelW<double> { static constexpr int W0 = 64; };
template <> struct PanelW<cplx> { static constexpr int W0 = 32; };

template <typename T, int W> static void launch_trsm_w(hs_fac* f, int f0, int nact, int J0, int j0, int NB, int cmode, int max_cols, cudaStream_t st) {
  if constexpr (W == 32 || W == 64) {
    // few columns in all (the panel chain of the upper levels): one warp per column
    static const long long wmax = getenv("HS_TRSM_WARP_COLS") ? atoll(getenv("HS_TRSM_WARP_COLS")) : 148 * 8 * 4;
    if ((long long)nact * max_cols <= wmax) {
      dim3 gw(nact, (max_cols + 7) / 8);
      k_swap_trsm_warp<T, W><<<gw, 256, 0, st>>>(f->d_fronts, (T*)f->pool, f->d_ipiv, f0, J0, j0, NB, cmode);
      CUDA_OK(cudaGetLastError());
      return;
    }
  }
  dim3 grid(nact, (max_cols + 127) / 128);
  k_swap_trsm<T, W><<<grid, 128, 0, st>>>(f->d_fronts, (T*)f->pool, f->d_ipiv, f0, J0, j0, NB, cmode);
  CUDA_OK(cudaGetLastError());
}

template <typename T> static void trsm_dispatch(hs_fac* f, int W, int f0, int nact, int J0, int j0, int NB, int cmode, int max_cols, cudaStream_t st) {
  constexpr int W0 = PanelW<T>::W0;
  if (max_cols <= 0) return;
  if (W == W0) launch_trsm_w<T, W0>(f, f0, nact, J0, j0, NB, cmode, max_cols, st);
  else if (W == W0 / 2) launch_trsm_w<T, W0 / 2>(f, f0, nact, J0, j0, NB, cmode, max_cols, st);
  else if (W == W0 / 4) launch_trsm_w<T, W0 / 4>(f, f0, nact, J0, j0, NB, cmode, max_cols, st);
  else if constexpr (W0 / 8 >= 8) launch_trsm_w<T, W0 / 8>(f, f0, nact, J0, j0, NB, cmode, max_cols, st);
  ++f->stats.launches_factor;
}

// partial LU of all fronts of one level (pivot block + Schur update), batched over the level.
// Two-level blocking: outer blocks of NB pivot columns, inner register-resident panels of width W.
template <typename T> static void factor_level(hs_fac* f, const Level& L) {
  cudaStream_t st = f->ctx->stream;
  if (L.max_n <= hs_small_max_n(f->dtype) && L.max_ni > 0) {
    // small fronts: one fused register-resident kernel does pivot block, both panels and the Schur update
    {
      PhaseTimer t(f, &f->stats.ms_small);
      hs_small_factor(f, L);
    }
    solve_prep(f, L);
    return;
  }
  // the panel cluster spans the rows pivots are taken from (the pivot block, or one half of a split pivot block)
  int max_prow = 0;
  bool has_split = false;
  for (int i = L.f0; i < L.f1; ++i) {
    const Front& fr = f->fronts[i];
    max_prow = std::max(max_prow, fr.split > 0 ? std::max(fr.split, fr.ni - fr.split) : fr.ni);
    has_split = has_split || fr.split > 0;
  }
  const int W = hs_panel_width(f, max_prow, L.f1 - L.f0);
  if (W < 0) throw hs_error(HS_ESIZE, "pivot block with " + std::to_string(max_prow) + " rows exceeds the panel kernels");
  if (L.max_ni > 0) {   // pivot order: the identity now, permuted by every panel's row moves
    dim3 g(L.f1 - L.f0, (L.max_ni + 255) / 256);
    k_rperm_init<<<g, 256, 0, st>>>(f->d_fronts, f->d_rperm, L.f0);
    CUDA_OK(cudaGetLastError());
    f->stats.launches_factor += 1;
  }
  auto pivot_rows = [&](int j0, int nact) {   // tallest panel among the active fronts
    if (!has_split) return L.max_ni - j0;
    int m = 0;
    for (int i = L.f0; i < L.f0 + nact; ++i) m = std::max(m, hs_plim(f->fronts[i], j0) - j0);
    return m;
  };
  const int NB = std::max(W, f->ctx->outer_block / W * W);
  constexpr int smem_gemm = gemm_smem_bytes<T>();
  using Cfg = GemmCfg<T>;
  auto nactive = [&](int j) {
    // fronts are sorted by ni descending: the active ones (ni > j) are a prefix
    return (int)(std::partition_point(L.ni_sorted.begin(), L.ni_sorted.end(), [&](int v) { return v > j; }) - L.ni_sorted.begin());
  };
  auto gemm = [&](int nact, int J0, int j0, int mode, int mrows, int mcols, cudaStream_t stream) {
    if (nact <= 0 || mrows <= 0 || mcols <= 0) return;
    const bool big = mode == 1 || mode == 3 || mode == 4;
    if (big) {   // flops of this trailing update, for the big / in-block split of the GEMM roofline
      const double cxf = f->dtype == HS_C64 ? 4.0 : 1.0;
      for (int i = L.f0; i < L.f0 + nact; ++i) {
        const Front& fr = f->fronts[i];
        if (fr.ni <= J0) continue;
        const int BE = std::min(J0 + NB, fr.ni), BE2 = std::min(BE + NB, fr.ni);
        const double rows = fr.n - BE, cols = mode == 1 ? fr.n - BE : (mode == 3 ? BE2 - BE : fr.n - BE2);
        f->stats.gemm_flops_big += cxf * 2.0 * rows * cols * (BE - J0);
      }
    }
    double dt = 0;
    {
      PhaseTimer t(f, &dt);
      dim3 grid(nact, (mrows + Cfg::TM - 1) / Cfg::TM + 1, (mcols + Cfg::TN - 1) / Cfg::TN);
      k_gemm<T><<<grid, gemm_threads<T>(), smem_gemm, stream>>>(f->d_fronts, (T*)f->pool, L.f0, J0, j0, NB, W, mode, nullptr);
      CUDA_OK(cudaGetLastError());
    }
    f->stats.ms_gemm += dt;
    if (big) f->stats.ms_gemm_big += dt;
    ++f->stats.gemm_launches;
    ++f->stats.launches_factor;
  };
  // look-ahead (few, large fronts): while the big trailing update of block k runs on the caller's stream, the panels
  // of block k+1 — which only need the update of their own columns and occupy one cluster per front — run on a
  // high-priority second stream.
  const bool lookahead = !f->ctx->profile && (L.f1 - L.f0) <= f->ctx->lookahead_max_fronts && L.max_ni > NB;
  cudaStream_t hi = f->ctx->aux_stream;
  // Rows below the pivot rows (boundary rows) never enter a pivot search: their triangular solves and their share of the
  // in-block updates run on the below-rows stream, one step behind the panel chain (few large fronts only — with many
  // fronts per level the launches are throughput-bound and the extra events cost more than they hide).
  cudaStream_t sb = f->ctx->below_stream;
  const bool below2 = sb && !f->ctx->profile && !has_split && L.max_nb > 0 && (L.f1 - L.f0) <= f->ctx->lookahead_max_fronts;
  bool below_pending = false;
  auto join_below = [&](cudaStream_t s1, cudaStream_t s2) {   // these streams are about to read the boundary rows of the block's columns
    if (!below_pending) return;
    CUDA_OK(cudaEventRecord(f->ctx->ev_below, sb));
    CUDA_OK(cudaStreamWaitEvent(s1, f->ctx->ev_below, 0));
    if (s2) CUDA_OK(cudaStreamWaitEvent(s2, f->ctx->ev_below, 0));
    below_pending = false;
  };
  auto phaseA = [&](int J0, int JE, cudaStream_t s_) {
    // factor the block's columns; interchanges, solves and updates stay inside the block
    for (int j0 = J0; j0 < JE; j0 += W) {
      const int nact = nactive(j0);
      if (nact == 0) break;
      const int m = L.max_n - j0;
      {
        PhaseTimer t(f, &f->stats.ms_panel);
        hs_panel_launch(f, W, L.f0, nact, j0, pivot_rows(j0, nact), s_);
        ++f->stats.panel_launches;
        ++f->stats.launches_factor;
      }
      if (below2) {
        CUDA_OK(cudaEventRecord(f->ctx->ev_pan, s_));
        trsm_dispatch<T>(f, W, L.f0, nact, J0, j0, NB, 0, JE - J0, s_);
        CUDA_OK(cudaEventRecord(f->ctx->ev_urow, s_));
        gemm(nact, J0, j0, 5, L.max_ni - j0, JE - j0, s_);
        CUDA_OK(cudaStreamWaitEvent(sb, f->ctx->ev_pan, 0));
        hs_trsm_rows(f, W, L.f0, nact, j0, L.max_nb, sb);
        ++f->stats.launches_factor;
        CUDA_OK(cudaStreamWaitEvent(sb, f->ctx->ev_urow, 0));
        gemm(nact, J0, j0, 6, L.max_nb, JE - j0, sb);
        below_pending = true;
        continue;
      }
      {
        PhaseTimer t(f, &f->stats.ms_trsm);
        // rows below the pivot rows (boundary rows, second half of a split pivot block): L21 = A_bi·U_pp⁻¹, one thread per row
        const int below = has_split ? L.max_n - j0 - 1 : L.max_nb;
        if (below > 0) { hs_trsm_rows(f, W, L.f0, nact, j0, below, s_); ++f->stats.launches_factor; }
        trsm_dispatch<T>(f, W, L.f0, nact, J0, j0, NB, 0, JE - J0, s_);
      }
      gemm(nact, J0, j0, 0, m, JE - j0, s_);
    }
  };
  bool a_on_hi = false;  // phase A of the current block was issued on the look-ahead stream
  for (int J0 = 0; J0 < L.max_ni; J0 += NB) {
    const int JE = std::min(J0 + NB, L.max_ni);
    if (a_on_hi) { CUDA_OK(cudaStreamWaitEvent(st, f->ctx->ev_c2, 0)); a_on_hi = false; }
    else phaseA(J0, JE, st);
    // phase B: the columns outside the block: all interchanges first, then solve / update sub-block by sub-block
    const int nactB = nactive(J0);
    const int mB = L.max_n - J0 - 1;  // at least one pivot column is gone
    if (nactB > 0 && L.max_n > 1) {
      {
        PhaseTimer t(f, &f->stats.ms_trsm);
        dim3 grid(nactB, (L.max_n + 127) / 128);
        k_laswp<T><<<grid, 128, 0, st>>>(f->d_fronts, (T*)f->pool, f->d_ipiv, L.f0, J0, NB);
        CUDA_OK(cudaGetLastError());
        ++f->stats.launches_factor;
      }
      for (int j0 = J0; j0 < JE; j0 += W) {
        const int nact = nactive(j0);
        if (nact == 0) break;
        {
          PhaseTimer t(f, &f->stats.ms_trsm);
          trsm_dispatch<T>(f, W, L.f0, nact, J0, j0, NB, 1, mB, st);
        }
        gemm(nact, J0, j0, 2, JE - j0, mB, st);
      }
      // phase C: the big update with K = BE − J0
      if (lookahead && JE < L.max_ni) {
        CUDA_OK(cudaEventRecord(f->ctx->ev_b, st));
        CUDA_OK(cudaStreamWaitEvent(hi, f->ctx->ev_b, 0));
        join_below(hi, st);   // before the next block queues its own boundary-row work
        gemm(nactB, J0, J0, 3, mB, NB, hi);                     // the next block's own columns …
        phaseA(JE, std::min(JE + NB, L.max_ni), hi);            // … and its panels, on the look-ahead stream
        CUDA_OK(cudaEventRecord(f->ctx->ev_c2, hi));
        a_on_hi = true;
        gemm(nactB, J0, J0, 4, mB, mB, st);                     // everything right of the next block
      } else {
        join_below(st, nullptr);
        gemm(nactB, J0, J0, 1, mB, mB, st);
      }
    }
  }
  if (a_on_hi) CUDA_OK(cudaStreamWaitEvent(st, f->ctx->ev_c2, 0));
  join_below(st, nullptr);
  {  // flops the GEMM launches of this level issue: Σ_steps 2·(n−j0−wc)²·wc per front
    const double cx = f->dtype == HS_C64 ? 4.0 : 1.0;
    for (int i = L.f0; i < L.f1; ++i) {
      const Front& fr = f->fronts[i];
      for (int j0 = 0; j0 < fr.ni; j0 += W) {
        const double wc = std::min(W, fr.ni - j0), mt = fr.n - j0 - wc;
        f->stats.gemm_flops += cx * 2.0 * mt * mt * wc;  // same total however the update is blocked
      }
    }
  }
  // solve preparation: invert the diagonal blocks of L11/U11 in place (see hs_solve.cu)
  if (L.max_ni > 0) solve_prep(f, L);
}

template <typename T> static void numeric(hs_fac* f) {
  cudaStream_t st = f->ctx->stream;
  T* pool = (T*)f->pool;
  hs_stats_t& s = f->stats;
  s.ms_assemble = s.ms_panel = s.ms_trsm = s.ms_gemm = s.ms_solve_prep = s.ms_small = s.ms_extend_add = s.ms_compress = 0;
  s.maxrank = 0;
  s.ms_hss = 0; s.hss_bytes = 0; s.hss_maxrank = 0; s.hss_rounds = 0; s.hss_nodes = 0; s.sketch_flops = 0;
  s.launches_factor = 0;
  s.gemm_launches = s.panel_launches = 0;
  s.gemm_flops = 0;
  s.gemm_flops_big = 0; s.ms_gemm_big = 0;
  CUDA_OK(cudaEventRecord(f->ev0, st));
  CUDA_OK(cudaMemsetAsync(f->d_info, 0, 4 * sizeof(int), st));
  for (size_t li = 0; li < f->levels.size(); ++li) {
    const Level& L = f->levels[li];
    const int nf = L.f1 - L.f0;
    if (!L.pseudo) {
      PhaseTimer t(f, &s.ms_assemble);
      CUDA_OK(cudaMemsetAsync(pool + L.poff0, 0, (size_t)(L.poff1 - L.poff0) * sizeof(T), st));
      if (L.tpoff1 > L.tpoff0) CUDA_OK(cudaMemsetAsync(pool + L.tpoff0, 0, (size_t)(L.tpoff1 - L.tpoff0) * sizeof(T), st));
      dim3 grid(nf, (L.max_n + 255) / 256);
      k_fill_owner<<<grid, 256, 0, st>>>(f->d_fronts, f->d_gidx, f->d_own, f->d_pos, L.f0);
      k_scatter_A<T><<<grid, 256, 0, st>>>(f->d_fronts, pool, f->d_gidx, f->d_own, f->d_pos, f->d_colptr,
                                            f->d_rowval, (const T*)f->d_nzval, L.f0);
      s.launches_factor += 3;
      for (int i = L.f0; i < L.f1; ++i) {  // external leaves: the front IS the imported Schur block
        if (!f->ext_src[i]) continue;
        const Front& fr = f->fronts[i];
        CUDA_OK(cudaMemcpy2DAsync(pool + fr.off, (size_t)fr.ld * sizeof(T), f->ext_src[i], (size_t)f->ext_ld[i] * sizeof(T),
                                  (size_t)fr.n * sizeof(T), fr.n, cudaMemcpyDeviceToDevice, st));
      }
      if (li > 0 && !f->levels[li - 1].pseudo) {
        const Level& Lc = f->levels[li - 1];
        if (Lc.max_nb > 0) {
          // a CTA covers ~8192 elements of a child's Schur block: whole small blocks, column groups of large ones;
          // lanes per column sized so that one or two passes cover the rows of the level's fronts
          constexpr int RPL = sizeof(T) == 8 ? 2 : 1;
          const int cols_per_cta = std::max(32, (8192 / Lc.max_nb + 31) / 32 * 32);
          dim3 g2(Lc.f1 - Lc.f0, (Lc.max_nb + cols_per_cta - 1) / cols_per_cta);
          PhaseTimer t2(f, &s.ms_extend_add);
          if (Lc.max_nb <= 8 * RPL * 4) k_extend_add<T, 2, 8><<<g2, 256, 0, st>>>(f->d_fronts, pool, f->d_cmap, Lc.f0, cols_per_cta);
          else if (Lc.max_nb <= 16 * RPL * 4) k_extend_add<T, 2, 16><<<g2, 256, 0, st>>>(f->d_fronts, pool, f->d_cmap, Lc.f0, cols_per_cta);
          else k_extend_add<T, 4, 32><<<g2, 256, 0, st>>>(f->d_fronts, pool, f->d_cmap, Lc.f0, cols_per_cta);
          s.launches_factor += 1;
        }
      }
      CUDA_OK(cudaGetLastError());
    }
    // the level's dense fronts, then its compressed ones: low-rank Gauss transforms, LU of the thin bordered fronts,
    // Schur complement back into the dense slots (hs_compress.cu)
    for (Level& FL : f->flevels) {
      if (FL.thin < 0) { if (FL.f0 == L.f0 && FL.pseudo == L.pseudo) factor_level<T>(f, FL); continue; }
      CompLevel& C = f->clevels[FL.thin];
      if (C.li != (int)li) continue;
      {
        PhaseTimer t(f, &s.ms_compress);
        hs_comp_prepare(f, C);
      }
      factor_level<T>(f, FL);
      {
        PhaseTimer t(f, &s.ms_compress);
        hs_comp_schur(f, C);
      }
      // HSS storage of the Schur complements (`randcompress_adaptive`, factorization.jl:102-110): matrix-free sketches,
      // interpolative decompositions per HSS level, then the represented matrix goes back into the dense slots
      if (!f->hss.empty()) {
        PhaseTimer t(f, &s.ms_hss);
        hs_hss_build(f, C);
      }
    }
  }
  s.lowrank_bytes = 0;
  for (const CompLevel& C : f->clevels) s.lowrank_bytes += (double)C.side_bytes;
  if (f->ctx->prep_pending) {
    CUDA_OK(cudaEventRecord(f->ctx->ev_p1, f->ctx->prep_stream));
    CUDA_OK(cudaStreamWaitEvent(st, f->ctx->ev_p1, 0));
    f->ctx->prep_pending = false;
  }
  CUDA_OK(cudaEventRecord(f->ev1, st));
  f->ctx->launches += s.launches_factor;
  int info[4];
  CUDA_OK(cudaMemcpyAsync(info, f->d_info, sizeof(info), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, f->ev0, f->ev1));
  s.ms_factor_total = ms;
  s.singular_front = -1;
  s.singular_col = -1;
  if (info[0]) {
    s.singular_front = f->fronts[info[1]].flags & 1 ? f->nnodes - 1 : -1;
    for (int64_t k = 0; k < f->nnodes; ++k)
      if (f->node2front[k] == info[1]) s.singular_front = k;
    s.singular_col = info[2];
  }
}

// ------------------------------------------------------------------------------------------------
// hs_factor
// ------------------------------------------------------------------------------------------------
static void check_opts(const hs_opts& o) {  // chkopts! HierarchicalSolvers.jl:73-79
  if (o.swsize < 1) throw hs_error(HS_EARG, "swsize");
  if (!(o.atol >= 0.)) throw hs_error(HS_EARG, "atol");
  if (!(o.rtol >= 0.)) throw hs_error(HS_EARG, "rtol");
  if (!(0. < o.c_tol && o.c_tol <= 1.)) throw hs_error(HS_EARG, "c_tol");
  if (o.leafsize < 1) throw hs_error(HS_EARG, "leafsize");
}

template <typename V> static void dev_upload(V** dst, const V* src, size_t count, cudaStream_t st) {
  CUDA_OK(cudaMalloc((void**)dst, std::max<size_t>(count, 1) * sizeof(V)));
  if (count) CUDA_OK(cudaMemcpyAsync(*dst, src, count * sizeof(V), cudaMemcpyHostToDevice, st));
}

// host-side loops over millions of index entries run on a few threads (the plan build is inside the end-to-end time)
template <typename Fn> static void parallel_for(int64_t n, Fn&& fn) {
  static const unsigned nt_max = [] {
    const char* e = getenv("HS_HOST_THREADS");
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    return e ? (unsigned)std::max(1, atoi(e)) : std::min(hw, 16u);
  }();
  const unsigned nt = (unsigned)std::min<int64_t>(nt_max, (n + 16383) / 16384);
  if (nt <= 1) { fn((int64_t)0, n); return; }
  std::vector<std::thread> th;
  std::exception_ptr ep;
  std::mutex mu;
  const int64_t chunk = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; ++t) {
    const int64_t lo = t * chunk, hi = std::min<int64_t>(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([&, lo, hi] {
      try { fn(lo, hi); } catch (...) { std::lock_guard<std::mutex> g(mu); if (!ep) ep = std::current_exception(); }
    });
  }
  for (auto& x : th) x.join();
  if (ep) std::rethrow_exception(ep);
}

struct PlanClock {
  bool on = getenv("HS_PLAN_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void tick(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[plan] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

static void build_plan(hs_fac* f, const hs_tree* t) {
  PlanClock clk;
  const int64_t nn = t->nnodes, base = t->index_base;
  if (nn <= 0) throw hs_error(HS_EARG, "hs_factor: empty tree");
  f->nnodes = nn;
  f->left.assign(t->left, t->left + nn);
  f->right.assign(t->right, t->right + nn);
  f->parent.assign(nn, -1);
  for (int64_t k = 0; k < nn; ++k) {
    for (int side = 0; side < 2; ++side) {
      int64_t& c = side ? f->right[k] : f->left[k];
      if (c == -1) continue;
      c -= base;
      if (c < 0 || c >= nn || c == k) throw hs_error(HS_EARG, "hs_factor: child id out of range");
      if (f->parent[c] != -1) throw hs_error(HS_EARG, "hs_factor: node has two parents");
      f->parent[c] = k;
    }
    // factorization.jl:24-26
    if ((f->left[k] == -1) != (f->right[k] == -1))
      throw hs_error(HS_ETREE, "Expected nested dissection to be a binary tree. Found a node with only one child.");
  }
  int64_t root = -1;
  for (int64_t k = 0; k < nn; ++k)
    if (f->parent[k] == -1) {
      if (root != -1) throw hs_error(HS_EARG, "hs_factor: more than one root");
      root = k;
    }
  if (root < 0) throw hs_error(HS_EARG, "hs_factor: no root");
  // levels: root = 1 as `_factor(..., 1)` (factorization.jl:9)
  f->level.assign(nn, 0);
  {
    std::vector<int64_t> st{root};
    f->level[root] = 1;
    int64_t seen = 0;
    while (!st.empty()) {
      int64_t k = st.back(); st.pop_back(); ++seen;
      for (int64_t c : {f->left[k], f->right[k]})
        if (c >= 0) { f->level[c] = f->level[k] + 1; st.push_back(c); }
    }
    if (seen != nn) throw hs_error(HS_EARG, "hs_factor: tree is not connected");
  }
  const int64_t maxlev = *std::max_element(f->level.begin(), f->level.end());
  f->depth = maxlev;
  clk.tick("tree links + levels");
  // int / bnd (global DOFs) are only read while the plan is built: used in place, `base` applied on the fly
  auto copy_ptr = [&](const int64_t* ptr, std::vector<int64_t>& p) {
    p.assign(ptr, ptr + nn + 1);
    if (p[0] != 0) throw hs_error(HS_EARG, "hs_factor: malformed index pointer array");
    for (int64_t k = 0; k < nn; ++k)
      if (p[k + 1] < p[k]) throw hs_error(HS_EARG, "hs_factor: malformed index pointer array");
  };
  std::vector<int64_t> int_ptr, bnd_ptr;
  copy_ptr(t->int_ptr, int_ptr);
  copy_ptr(t->bnd_ptr, bnd_ptr);
  const int64_t *int_idx = t->int_idx, *bnd_idx = t->bnd_idx;
  auto copy_loc = [&](const int64_t* ptr, const int64_t* idx, std::vector<int64_t>& p, IntBuf& v) {
    copy_ptr(ptr, p);
    v.reserve(p[nn]);
    v.set_size(p[nn]);
    int* dst = v.data();
    parallel_for(p[nn], [&](int64_t lo, int64_t hi) {
      for (int64_t q = lo; q < hi; ++q) {
        const int64_t a = idx[q] - base;
        if (a < 0 || a >= (1ll << 30)) throw hs_error(HS_EARG, "hs_factor: nd_loc position out of range");
        dst[q] = (int)a;
      }
    });
  };
  copy_loc(t->iloc_ptr, t->iloc_idx, f->iloc_ptr, f->iloc_idx);
  copy_loc(t->bloc_ptr, t->bloc_idx, f->bloc_ptr, f->bloc_idx);
  f->node_ni.resize(nn);
  f->node_nb.resize(nn);
  for (int64_t k = 0; k < nn; ++k) {
    f->node_ni[k] = (int)(int_ptr[k + 1] - int_ptr[k]);
    f->node_nb[k] = (int)(bnd_ptr[k + 1] - bnd_ptr[k]);
    if ((int64_t)f->node_ni[k] + f->node_nb[k] > (1 << 30)) throw hs_error(HS_ESIZE, "front too large");
  }
  clk.tick("copy index sets");
  parallel_for(int_ptr[nn], [&](int64_t lo, int64_t hi) {
    for (int64_t q = lo; q < hi; ++q)
      if (int_idx[q] < base || int_idx[q] - base >= f->n) throw hs_error(HS_EARG, "hs_factor: DOF index out of range in int");
  });
  parallel_for(bnd_ptr[nn], [&](int64_t lo, int64_t hi) {
    for (int64_t q = lo; q < hi; ++q)
      if (bnd_idx[q] < base || bnd_idx[q] - base >= f->n) throw hs_error(HS_EARG, "hs_factor: DOF index out of range in bnd");
  });
  // consistency of (nd, nd_loc): a branch's sets are the concatenation of its children's selected boundary rows
  // (nesteddissection.jl:64-65, factorization.jl:63-64)
  auto nloc = [&](const std::vector<int64_t>& p, int64_t k) { return p[k + 1] - p[k]; };
  parallel_for(nn, [&](int64_t klo, int64_t khi) {
  for (int64_t k = klo; k < khi; ++k) {
    if (f->left[k] < 0) continue;
    const int64_t l = f->left[k], r = f->right[k];
    const int64_t nil = nloc(f->iloc_ptr, l), nir = nloc(f->iloc_ptr, r), nbl = nloc(f->bloc_ptr, l), nbr = nloc(f->bloc_ptr, r);
    if (nil + nir != f->node_ni[k] || nbl + nbr != f->node_nb[k])
      throw hs_error(HS_EDIM, "hs_factor: node " + std::to_string(k) + " int/bnd sizes do not match its children's nd_loc");
    auto chk = [&](int64_t c, const std::vector<int64_t>& lp, IntBuf& li, const int64_t* dst) {
      for (int64_t q = lp[c]; q < lp[c + 1]; ++q) {
        const int64_t a = li[q];
        if (a < 0 || a >= f->node_nb[c]) throw hs_error(HS_EARG, "hs_factor: nd_loc position out of range");
        if (bnd_idx[bnd_ptr[c] + a] != dst[q - lp[c]])
          throw hs_error(HS_EARG, "hs_factor: (nd, nd_loc) inconsistent — was the tree produced by symfact!?");
      }
    };
    chk(l, f->iloc_ptr, f->iloc_idx, int_idx + int_ptr[k]);
    chk(r, f->iloc_ptr, f->iloc_idx, int_idx + int_ptr[k] + nil);
    chk(l, f->bloc_ptr, f->bloc_idx, bnd_idx + bnd_ptr[k]);
    chk(r, f->bloc_ptr, f->bloc_idx, bnd_idx + bnd_ptr[k] + nbl);
  }
  });
  clk.tick("validate");
  // swlevel < 0 is relative to the tree depth (factorization.jl:8);
  // compression_flag = (level ≤ swlevel) && (|bnd| ≥ swsize) (factorization.jl:15).  Compressed leaves keep dense
  // L and R in the reference too (:45-59) and S stays dense here, so only branches change.
  f->swlevel_resolved = f->opts.swlevel < 0 ? std::max<int64_t>(f->depth + f->opts.swlevel, 0) : f->opts.swlevel;
  std::vector<char> cflag(nn, 0);
  for (int64_t k = 0; k < nn; ++k)   // subtree mode included: the root's dense slot is persistent, so its S can be exported
      cflag[k] = f->left[k] >= 0 && f->level[k] <= f->swlevel_resolved && f->node_nb[k] >= f->opts.swsize &&
                 f->node_ni[k] > 0 && f->node_nb[k] > 0;
  // which compressed nodes will hold an HSS Schur complement (cluster root not a single leaf, hs_hss.cu) and which of them
  // have the first cluster split forced at the int/bnd boundary (0 < |int_loc| < |perm|, factorization.jl:109): a
  // compressed node whose two children both do takes the HSS-children methods (:86-91 by dispatch)
  std::vector<char> hss_forced(nn, 0);
  if (f->opts.hss)
    for (int64_t k = 0; k < nn; ++k) {
      const int64_t n1 = nloc(f->iloc_ptr, k), m = n1 + nloc(f->bloc_ptr, k);
      hss_forced[k] = cflag[k] && n1 > 0 && n1 < m;
    }
  // front order: deepest level first; inside a level the dense fronts before the compressed ones, each group by ni
  // descending (active panels form a prefix)
  std::vector<int64_t> order(nn);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    if (f->level[a] != f->level[b]) return f->level[a] > f->level[b];
    if (cflag[a] != cflag[b]) return cflag[a] < cflag[b];
    return f->node_ni[a] > f->node_ni[b];
  });
  const bool pseudo = f->node_nb[root] > 0 && !f->opts.subtree;
  const int nfr = (int)nn + (pseudo ? 1 : 0);
  f->fronts.assign(nfr, Front{});
  f->node2front.assign(nn, -1);
  for (int i = 0; i < (int)nn; ++i) f->node2front[order[i]] = i;
  f->root_front = f->node2front[root];
  // rows one panel cluster covers with its narrowest register tile (hs_panel.cuh: 256·R rows per CTA, R ≤ 8 (f64) / 4 (c64))
  int prow_cap = 256 * (f->dtype == HS_F64 ? 8 : 4) * std::max(1, f->ctx->max_cluster);
  if (getenv("HS_PROW_CAP")) prow_cap = std::min(prow_cap, std::max(64, atoi(getenv("HS_PROW_CAP"))));  // tests: force the split path
  IntBuf &gidx = f->ctx->sc_gidx, &cmap = f->ctx->sc_cmap;  // grow-only scratch: no page faults after the first call
  long long poff = 0, ioff = 0;
  const long long align = 32;
  // Dense slots of compressed fronts are transient: they are needed while their level is processed and when the
  // parent level is assembled, so they live in two arenas used alternately by level parity (HS_KEEP_SCHUR=1 keeps
  // them, e.g. to read F.S of a compressed node back).  The root stays persistent (its S may be factored in place).
  f->transient_schur = !(getenv("HS_KEEP_SCHUR") && atoi(getenv("HS_KEEP_SCHUR")) != 0);
  std::vector<long long> toff(nn, -1);
  long long arena[2] = {0, 0}, tcur = 0;
  f->levels.clear();
  double flops = 0, sbytes = 0, ebytes = 0;
  int64_t max_ni = 0, max_nb = 0;
  for (int i = 0; i < (int)nn; ++i) {
    const int64_t k = order[i];
    Front& fr = f->fronts[i];
    fr.ni = f->node_ni[k];
    fr.n = fr.ni + f->node_nb[k];
    fr.ld = (fr.n + 1) & ~1;
    if (fr.ld == 0) fr.ld = 2;
    fr.off = poff;
    fr.ioff = ioff;
    fr.parent = f->parent[k] >= 0 ? f->node2front[f->parent[k]] : -1;
    fr.flags = 0;
    if (f->left[k] < 0) { fr.ni_l = -1; fr.nb_l = 0; }
    else { fr.ni_l = (int)nloc(f->iloc_ptr, f->left[k]); fr.nb_l = (int)nloc(f->bloc_ptr, f->left[k]); }
    fr.split = 0;
    if (fr.ni > prow_cap) {
      // pivot block taller than one panel cluster: eliminate it as the reference's blockfactor does (blockmatrix.jl:115-120),
      // [A11 A12; A21 A22] with pivoting inside A11 and inside S22 — split at the children's boundary, moved down to a
      // multiple of the outer block so that no panel straddles it
      const int NB0 = std::max(64, f->ctx->outer_block / 64 * 64);
      int sp = fr.ni_l > 0 ? fr.ni_l / NB0 * NB0 : 0;
      if (sp <= 0 || sp > prow_cap || fr.ni - sp > prow_cap)
        throw hs_error(HS_ESIZE, "pivot block with " + std::to_string(fr.ni) + " rows exceeds the panel kernels (" + std::to_string(prow_cap) +
                                     " rows per diagonal block)");
      fr.split = sp;
    }
    if (f->levels.empty() || f->level[k] != f->level[order[f->levels.back().f0]]) {
      Level L; L.f0 = i; L.f1 = i; L.fm = i; L.ioff0 = ioff; L.poff0 = poff;
      f->levels.push_back(L);
      tcur = 0;
    }
    Level& L = f->levels.back();
    L.f1 = i + 1;
    if (!cflag[k]) L.fm = i + 1;  // dense fronts come first: fm ends up one past the last of them
    else if (L.fm < L.f0) L.fm = L.f0;
    L.max_n = std::max(L.max_n, fr.n);
    L.max_ni = std::max(L.max_ni, fr.ni);
    L.max_nb = std::max(L.max_nb, fr.n - fr.ni);
    L.ni_sorted.push_back(fr.ni);
    const long long slot = ((long long)fr.ld * fr.n + align - 1) / align * align;
    if (cflag[k] && f->transient_schur && k != root) {
      const int par = (int)((f->levels.size() - 1) & 1);
      toff[i] = tcur; tcur += slot;
      arena[par] = std::max(arena[par], tcur);
      L.tpoff1 = tcur;  // relative to the arena until the arenas are placed (below)
    } else {
      poff += slot;
    }
    ioff += fr.n;
    L.ioff1 = ioff; L.poff1 = poff;
    const double ni = fr.ni, nb = fr.n - fr.ni;
    flops += 2.0 / 3.0 * ni * ni * ni + 2.0 * ni * ni * nb + 2.0 * ni * nb * nb;
    sbytes += ni * ni + 2.0 * ni * nb;
    ebytes += 2.0 * nb * nb;
    max_ni = std::max<int64_t>(max_ni, fr.ni);
    max_nb = std::max<int64_t>(max_nb, fr.n - fr.ni);
  }
  {  // place the two arenas behind the persistent fronts
    const long long abase[2] = {poff, poff + arena[0]};
    for (size_t li = 0; li < f->levels.size(); ++li) {
      Level& L = f->levels[li];
      if (L.tpoff1 == 0) continue;
      L.tpoff0 = abase[li & 1];
      L.tpoff1 += abase[li & 1];
      for (int i = L.fm; i < L.f1; ++i)
        if (toff[i] >= 0) f->fronts[i].off = abase[li & 1] + toff[i];
    }
    poff += arena[0] + arena[1];
  }
  clk.tick("front descriptors");
  // per-row tables: global DOF of every front row
  {
    size_t cap = (size_t)ioff + (pseudo ? f->node_nb[root] : 0);
    for (int64_t k = 0; k < nn; ++k)
      if (cflag[k]) cap += 2 * ((size_t)f->node_ni[k] + f->node_nb[k]);  // thin descriptors (border ≤ ni + nb rows)
    gidx.reserve(cap); cmap.reserve(cap);
    gidx.set_size(ioff); cmap.set_size(ioff);
  }
  parallel_for(nn, [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t k = order[i];
      int* g = gidx.data() + f->fronts[i].ioff;
      int* cm = cmap.data() + f->fronts[i].ioff;
      for (int q = 0; q < f->fronts[i].n; ++q) cm[q] = -1;
      for (int64_t q = int_ptr[k]; q < int_ptr[k + 1]; ++q) *g++ = (int)(int_idx[q] - base);
      for (int64_t q = bnd_ptr[k]; q < bnd_ptr[k + 1]; ++q) *g++ = (int)(bnd_idx[q] - base);
    }
  });
  clk.tick("gidx");
  // child → parent row maps: the children's S[perm,perm] lands as diagonal blocks of the parent front
  // (factorization.jl:41,74 and :118-121)
  parallel_for(nn, [&](int64_t klo, int64_t khi) {
  for (int64_t k = klo; k < khi; ++k) {
    if (f->left[k] < 0) continue;
    const int64_t l = f->left[k], r = f->right[k];
    const int nil = (int)nloc(f->iloc_ptr, l), nbl = (int)nloc(f->bloc_ptr, l), nip = f->node_ni[k];
    auto fill = [&](int64_t c, int ioffset, int boffset) {
      const Front& fc = f->fronts[f->node2front[c]];
      int* m = cmap.data() + fc.ioff + fc.ni;
      for (int64_t q = f->iloc_ptr[c]; q < f->iloc_ptr[c + 1]; ++q) m[f->iloc_idx[q]] = ioffset + (int)(q - f->iloc_ptr[c]);
      for (int64_t q = f->bloc_ptr[c]; q < f->bloc_ptr[c + 1]; ++q) m[f->bloc_idx[q]] = nip + boffset + (int)(q - f->bloc_ptr[c]);
    };
    fill(l, 0, 0);
    fill(r, nil, nbl);
  }
  });
  clk.tick("cmap");
  if (pseudo) {  // root boundary solve `F.S \ C[F.bnd,:]` (factornode.jl:72) as one more dense front
    const Front& rf = f->fronts[f->root_front];
    Front& pf = f->fronts[nn];
    pf.n = pf.ni = rf.n - rf.ni;
    pf.ld = rf.ld;
    pf.off = rf.off + (long long)rf.ni * rf.ld + rf.ni;
    pf.ioff = ioff;
    pf.parent = -1;
    pf.ni_l = -1; pf.nb_l = 0;
    pf.flags = 1;
    Level L; L.f0 = (int)nn; L.f1 = (int)nn + 1; L.fm = L.f1; L.max_n = L.max_ni = pf.n; L.max_nb = 0;
    L.ioff0 = ioff; L.ioff1 = ioff + pf.n; L.poff0 = L.poff1 = poff; L.pseudo = true;
    L.ni_sorted.push_back(pf.ni);
    f->levels.push_back(L);
    f->pseudo_front = (int)nn;
    for (int64_t q = bnd_ptr[root]; q < bnd_ptr[root + 1]; ++q) gidx.push_back((int)(bnd_idx[q] - base));
    ioff += pf.n;
    cmap.resize(ioff, -1);
    const double ni = pf.ni;
    flops += 2.0 / 3.0 * ni * ni * ni;
    sbytes += ni * ni;
  }
  f->ext_src.assign(f->fronts.size(), nullptr);
  f->ext_ld.assign(f->fronts.size(), 0);
  f->pool_elems = poff;
  f->idx_total = ioff;
  // factor / solve levels.  A compressed front gets a second ("thin") descriptor at nfr + its id, rows
  // [int; r border rows]; the border rows live in virtual slots n + voff + k of the internal solution vector.
  f->nfr = nfr;
  f->fronts.resize(2 * (size_t)nfr, Front{});
  f->flevels.clear(); f->comp.clear(); f->clevels.clear();
  f->nvirt = 0;
  std::vector<int> front2comp(nfr, -1);
  for (size_t li = 0; li < f->levels.size(); ++li) {
    const Level& L = f->levels[li];
    if (L.fm > L.f0 || L.f1 == L.f0) {
      Level D = L;
      D.f1 = L.fm; D.thin = -1;
      if (L.fm < L.f1) {  // only the dense part
        D.max_n = D.max_ni = D.max_nb = 0;
        D.ni_sorted.assign(L.ni_sorted.begin(), L.ni_sorted.begin() + (L.fm - L.f0));
        for (int i = L.f0; i < L.fm; ++i) {
          D.max_n = std::max(D.max_n, f->fronts[i].n); D.max_ni = std::max(D.max_ni, f->fronts[i].ni);
          D.max_nb = std::max(D.max_nb, f->fronts[i].n - f->fronts[i].ni);
        }
        D.ioff0 = f->fronts[L.f0].ioff; D.ioff1 = f->fronts[L.fm - 1].ioff + f->fronts[L.fm - 1].n;
      }
      f->flevels.push_back(D);
    }
    if (L.fm < L.f1) {
      CompLevel C; C.li = (int)li; C.c0 = (int)f->comp.size();
      Level Tl; Tl.f0 = nfr + L.fm; Tl.f1 = nfr + L.f1; Tl.fm = Tl.f1; Tl.thin = (int)f->clevels.size();
      Tl.ioff0 = ioff;
      for (int i = L.fm; i < L.f1; ++i) {
        const Front& fd = f->fronts[i];
        CompFront cf; cf.fi = i; cf.ni = fd.ni; cf.nb = fd.n - fd.ni; cf.rcap = std::min(cf.ni, cf.nb);
        {
          const int64_t k = order[i], l = f->left[k], r = f->right[k];
          cf.hchild = l >= 0 && hss_forced[l] && hss_forced[r];
          if (cf.hchild) { cf.cl = front2comp[f->node2front[l]]; cf.cr = front2comp[f->node2front[r]]; }
          cf.ni_l = fd.ni_l; cf.nb_l = fd.nb_l;
          front2comp[i] = (int)f->comp.size();
        }
        cf.vcap = cf.hchild ? cf.ni + cf.nb : cf.rcap;
        cf.voff = f->nvirt; f->nvirt += cf.vcap;
        f->comp.push_back(cf);
        Front& th = f->fronts[nfr + i];
        th = fd;
        th.ioff = ioff; th.parent = -1; th.ni_l = -1; th.nb_l = 0; th.flags = 0;
        th.n = fd.ni; th.ld = (fd.ni + 1) & ~1;  // set per factorization once the ranks are known
        for (int q = 0; q < fd.ni; ++q) gidx.push_back(gidx[fd.ioff + q]);
        for (int q = 0; q < cf.vcap; ++q) gidx.push_back((int)(f->n + cf.voff + q));
        ioff += fd.ni + cf.vcap;
        Tl.ni_sorted.push_back(fd.ni);
        Tl.max_ni = std::max(Tl.max_ni, fd.ni);
      }
      Tl.max_n = Tl.max_ni; Tl.max_nb = 0;
      Tl.ioff1 = ioff;
      C.c1 = (int)f->comp.size();
      C.flevel = (int)f->flevels.size();
      f->flevels.push_back(Tl);
      f->clevels.push_back(C);
    }
  }
  if ((long long)f->n + f->nvirt >= (1ll << 31)) throw hs_error(HS_ESIZE, "solution vector too long for 32-bit row tables");
  f->xld = f->n + f->nvirt;
  cmap.resize(ioff, -1);
  f->max_level_idx = 0;
  for (auto& L : f->flevels) f->max_level_idx = std::max(f->max_level_idx, L.ioff1 - L.ioff0 + 0);
  for (auto& C : f->clevels) {  // the border of a thin front may grow to rcap rows
    long long span = 0;
    for (int c = C.c0; c < C.c1; ++c) span += f->comp[c].ni + f->comp[c].vcap;
    f->max_level_idx = std::max(f->max_level_idx, span);
  }
  const double cx = f->dtype == HS_C64 ? 4.0 : 1.0;
  hs_stats_t& s = f->stats;
  s.nnodes = nn; s.nlevels = maxlev; s.n = f->n; s.max_ni = max_ni; s.max_nb = max_nb;
  s.factor_flops = flops * cx;
  s.solve_bytes = sbytes * f->esz;
  s.extadd_bytes = ebytes * f->esz;
  s.front_bytes = (double)poff * f->esz;
  s.singular_front = s.singular_col = -1;
  s.maxrank = 0;
  // upload
  clk.tick("levels + stats");
  cudaStream_t st = f->ctx->stream;
  CUDA_OK(cudaMalloc(&f->pool, std::max<size_t>((size_t)poff, 1) * f->esz));
  clk.tick("cudaMalloc pool");
  dev_upload(&f->d_fronts, f->fronts.data(), f->fronts.size(), st);
  dev_upload(&f->d_gidx, gidx.data(), gidx.size(), st);
  dev_upload(&f->d_cmap, cmap.data(), cmap.size(), st);
  CUDA_OK(cudaMalloc((void**)&f->d_ipiv, std::max<size_t>(ioff, 1) * sizeof(int)));
  CUDA_OK(cudaMalloc((void**)&f->d_rperm, std::max<size_t>(ioff, 1) * sizeof(int)));
  CUDA_OK(cudaMalloc((void**)&f->d_own, std::max<size_t>(f->n, 1) * sizeof(int)));
  CUDA_OK(cudaMalloc((void**)&f->d_pos, std::max<size_t>(f->n, 1) * sizeof(int)));
  CUDA_OK(cudaMalloc((void**)&f->d_info, 4 * sizeof(int)));
  CUDA_OK(cudaMemsetAsync(f->d_own, 0xff, std::max<size_t>(f->n, 1) * sizeof(int), st));
  CUDA_OK(cudaMemsetAsync(f->d_ipiv, 0, std::max<size_t>(ioff, 1) * sizeof(int), st));
  CUDA_OK(cudaMemsetAsync(f->d_rperm, 0, std::max<size_t>(ioff, 1) * sizeof(int), st));
  CUDA_OK(cudaStreamSynchronize(st));  // host vectors go out of scope
  clk.tick("upload tables");
  hs_comp_plan(f);
  hs_hss_plan(f);
}

__global__ void k_widen_index(long long* dst, const int* src, long long n, long long base) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (long long)src[i] - base;
}

__global__ void k_shift_index(long long* a, long long n, long long base) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] -= base;
}

static int32_t factor_impl(hs_ctx* ctx, hs_dtype dtype, int64_t n, const int64_t* colptr, const int64_t* rowval,
                           const void* nzval, const hs_tree* tree, const hs_opts* opts, int32_t flags,
                           hs_fac** out, bool do_numeric) {
  HS_TRY_BEGIN
  const bool on_device = flags & HS_ON_DEVICE;
  const bool idx32 = flags & HS_CSC_INT32;
  if (on_device && idx32) return hs_fail(HS_EARG, "hs_factor: device-resident CSC arrays must be int64");
  if (!ctx || !colptr || !rowval || !nzval || !tree || !out) return hs_fail(HS_EARG, "hs_factor: null argument");
  if (dtype != HS_F64 && dtype != HS_C64) return hs_fail(HS_EARG, "hs_factor: dtype must be HS_F64 or HS_C64");
  if (n <= 0) return hs_fail(HS_EARG, "hs_factor: n must be positive");
  CUDA_OK(cudaSetDevice(ctx->device));
  auto t_begin = std::chrono::steady_clock::now();
  std::unique_ptr<hs_fac> f(new hs_fac());
  f->ctx = ctx;
  f->dtype = dtype;
  f->esz = dtype == HS_C64 ? 16 : 8;
  f->n = n;
  if (opts) f->opts = *opts;
  else { f->opts = hs_opts{5, 1, 1e-6, 1e-6, 0.5, 32, -1, 10, 0, 0}; }
  check_opts(f->opts);
  CUDA_OK(cudaEventCreate(&f->ev0));
  CUDA_OK(cudaEventCreate(&f->ev1));
  // the matrix goes to the device on a second host thread (pageable copies block their caller) while this one builds
  // the plan; it uses the look-ahead stream, which is idle until the numeric phase
  const int64_t base = (flags & HS_CSC_ZERO_BASED) ? 0 : tree->index_base;
  hs_fac* fp = f.get();
  double ms_up = 0;
  auto upload = [&, fp]() {
    auto t0 = std::chrono::steady_clock::now();
    CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->aux_stream;
    CUDA_OK(cudaMalloc((void**)&fp->d_colptr, (size_t)(n + 1) * sizeof(long long)));
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int64_t last = 0;
    if (on_device) {
      CUDA_OK(cudaMemcpyAsync(&last, colptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
    } else last = idx32 ? (int64_t)((const int32_t*)colptr)[n] : colptr[n];
    fp->nnz = last - (on_device ? 0 : base);
    if (fp->nnz < 0) throw hs_error(HS_EARG, "hs_factor: negative nnz");
    CUDA_OK(cudaMalloc((void**)&fp->d_rowval, std::max<size_t>(fp->nnz, 1) * sizeof(long long)));
    CUDA_OK(cudaMalloc(&fp->d_nzval, std::max<size_t>(fp->nnz, 1) * fp->esz));
    if (idx32) {
      // int32 index arrays (SciPy): stage in a scratch buffer, widen to int64 on the device
      int* t32 = nullptr;
      CUDA_OK(cudaMalloc((void**)&t32, (size_t)(n + 1 + std::max<int64_t>(fp->nnz, 1)) * sizeof(int)));
      CUDA_OK(cudaMemcpyAsync(t32, colptr, (size_t)(n + 1) * sizeof(int), kind, st));
      CUDA_OK(cudaMemcpyAsync(t32 + n + 1, rowval, (size_t)fp->nnz * sizeof(int), kind, st));
      k_widen_index<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(fp->d_colptr, t32, n + 1, base);
      if (fp->nnz) k_widen_index<<<(unsigned)((fp->nnz + 255) / 256), 256, 0, st>>>(fp->d_rowval, t32 + n + 1, fp->nnz, base);
      CUDA_OK(cudaStreamSynchronize(st));
      cudaFree(t32);
    } else {
      CUDA_OK(cudaMemcpyAsync(fp->d_colptr, colptr, (size_t)(n + 1) * sizeof(long long), kind, st));
      CUDA_OK(cudaMemcpyAsync(fp->d_rowval, rowval, (size_t)fp->nnz * sizeof(long long), kind, st));
      if (!on_device && base != 0) {
        k_shift_index<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(fp->d_colptr, n + 1, base);
        if (fp->nnz) k_shift_index<<<(unsigned)((fp->nnz + 255) / 256), 256, 0, st>>>(fp->d_rowval, fp->nnz, base);
      }
    }
    CUDA_OK(cudaMemcpyAsync(fp->d_nzval, nzval, (size_t)fp->nnz * fp->esz, kind, st));
    CUDA_OK(cudaStreamSynchronize(st));
    ms_up = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  };
  // device-resident inputs may still be in flight on the caller's stream; the copies below run on another one
  if (on_device) CUDA_OK(cudaStreamSynchronize(ctx->stream));
  std::exception_ptr up_err;
  std::thread up_thread([&] { try { upload(); } catch (...) { up_err = std::current_exception(); } });
  try { build_plan(f.get(), tree); } catch (...) { up_thread.join(); throw; }
  auto t_plan = std::chrono::steady_clock::now();
  f->stats.ms_analyze = std::chrono::duration<double, std::milli>(t_plan - t_begin).count();
  up_thread.join();
  if (up_err) std::rethrow_exception(up_err);
  f->stats.ms_h2d = ms_up;  // overlapped with ms_analyze
  if (getenv("HS_PLAN_TIMING"))
    fprintf(stderr, "[factor] plan %.2f ms, upload %.2f ms (overlapped), joined at %.2f ms\n", f->stats.ms_analyze, ms_up,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  if (do_numeric) { if (dtype == HS_F64) numeric<double>(f.get()); else numeric<cplx>(f.get()); }
  if (getenv("HS_PLAN_TIMING"))
    fprintf(stderr, "[factor] numeric done at %.2f ms\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  *out = f.release();
  if ((*out)->stats.singular_front >= 0 || (*out)->stats.singular_col >= 0)
    return hs_fail(HS_ESINGULAR, "hs_factor: exactly singular pivot block in node " + std::to_string((*out)->stats.singular_front) +
                                     ", column " + std::to_string((*out)->stats.singular_col));
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_factor(hs_ctx* ctx, hs_dtype dtype, int64_t n, const int64_t* colptr, const int64_t* rowval,
                             const void* nzval, const hs_tree* tree, const hs_opts* opts, int32_t on_device,
                             hs_fac** out) {
  return factor_impl(ctx, dtype, n, colptr, rowval, nzval, tree, opts, on_device, out, true);
}

extern "C" int32_t hs_analyze(hs_ctx* ctx, hs_dtype dtype, int64_t n, const int64_t* colptr, const int64_t* rowval,
                              const void* nzval, const hs_tree* tree, const hs_opts* opts, int32_t on_device,
                              hs_fac** out) {
  return factor_impl(ctx, dtype, n, colptr, rowval, nzval, tree, opts, on_device, out, false);
}

extern "C" int32_t hs_schur_export(hs_fac* f, int64_t node, void* dst, int64_t ld) {
  HS_TRY_BEGIN
  if (!f || !dst) return hs_fail(HS_EARG, "hs_schur_export: null argument");
  if (node < 0 || node >= f->nnodes) return hs_fail(HS_EARG, "hs_schur_export: node out of range");
  const Front& fr = f->fronts[f->node2front[node]];
  const int nb = fr.n - fr.ni;
  if (ld < nb) return hs_fail(HS_EDIM, "hs_schur_export: leading dimension smaller than the boundary");
  if (f->pseudo_front >= 0 && f->node2front[node] == f->root_front)
    return hs_fail(HS_EARG, "hs_schur_export: the root boundary block was factored for the solve; factor with opts.subtree = 1");
  if (f->transient_schur && f->node2front[node] != f->root_front)
    for (const CompFront& cf : f->comp)
      if (cf.fi == f->node2front[node]) return hs_fail(HS_EARG, "hs_schur_export: Schur block of a compressed front is transient (HS_KEEP_SCHUR=1 keeps it)");
  CUDA_OK(cudaSetDevice(f->ctx->device));
  if (nb) {
    const char* src = (const char*)f->pool + ((size_t)fr.off + (size_t)fr.ni * fr.ld + fr.ni) * f->esz;
    CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)ld * f->esz, src, (size_t)fr.ld * f->esz, (size_t)nb * f->esz, nb,
                              cudaMemcpyDeviceToDevice, f->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(f->ctx->stream));
  }
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_schur_import(hs_fac* f, int64_t node, const void* src, int64_t ld) {
  HS_TRY_BEGIN
  if (!f || !src) return hs_fail(HS_EARG, "hs_schur_import: null argument");
  if (node < 0 || node >= f->nnodes) return hs_fail(HS_EARG, "hs_schur_import: node out of range");
  const int fi = f->node2front[node];
  Front& fr = f->fronts[fi];
  if (f->left[node] >= 0 || fr.ni != 0) return hs_fail(HS_EARG, "hs_schur_import: node must be a leaf with an empty int set");
  if (ld < fr.n) return hs_fail(HS_EDIM, "hs_schur_import: leading dimension smaller than the boundary");
  CUDA_OK(cudaSetDevice(f->ctx->device));
  f->ext_src[fi] = src;
  f->ext_ld[fi] = ld;
  if (!(fr.flags & 4)) {  // tell the assembly kernels not to gather A into this front
    fr.flags |= 4;
    CUDA_OK(cudaMemcpy(f->d_fronts + fi, &fr, sizeof(Front), cudaMemcpyHostToDevice));
  }
  return HS_OK;
  HS_TRY_END
}

template <typename T> static void sweep_impl(hs_fac* f, int64_t nrhs, void* xv, int which) {
  cudaStream_t st = f->ctx->stream;
  if (nrhs > f->rhs_cap) {
    cudaFree(f->d_x); cudaFree(f->d_work);
    f->d_x = f->d_work = nullptr;
    CUDA_OK(cudaMalloc(&f->d_x, (size_t)f->xld * nrhs * sizeof(T)));
    CUDA_OK(cudaMalloc(&f->d_work, std::max<size_t>((size_t)f->max_level_idx * nrhs, 1) * sizeof(T)));
    f->rhs_cap = nrhs;
  }
  f->stats.launches_solve = 0;
  if (!f->nvirt) {
    hs_solve_run(f, nrhs, xv, which);
  } else {
    // compressed fronts: the border rows of the thin fronts live in virtual slots behind the n entries of x, so the
    // sweep runs on the internal buffer (leading dimension n + nvirt).  The slots carry nothing from the forward to
    // the backward half (k_lr_bwd rewrites them), so the two halves may be separate calls.
    T* x = (T*)f->d_x;
    CUDA_OK(cudaMemcpy2DAsync(x, (size_t)f->xld * sizeof(T), xv, (size_t)f->n * sizeof(T), (size_t)f->n * sizeof(T), nrhs,
                              cudaMemcpyDeviceToDevice, st));
    CUDA_OK(cudaMemset2DAsync(x + f->n, (size_t)f->xld * sizeof(T), 0, (size_t)f->nvirt * sizeof(T), nrhs, st));
    hs_solve_run(f, nrhs, x, which);
    CUDA_OK(cudaMemcpy2DAsync(xv, (size_t)f->n * sizeof(T), x, (size_t)f->xld * sizeof(T), (size_t)f->n * sizeof(T), nrhs,
                              cudaMemcpyDeviceToDevice, st));
  }
  f->ctx->launches += f->stats.launches_solve;
  CUDA_OK(cudaStreamSynchronize(st));
}

extern "C" int32_t hs_solve_sweep(hs_fac* f, int64_t nrhs, void* x, int64_t ldx, int32_t which) {
  HS_TRY_BEGIN
  if (!f || !x) return hs_fail(HS_EARG, "hs_solve_sweep: null argument");
  if (nrhs <= 0 || !(which & 3)) return HS_OK;
  if (ldx != f->n) return hs_fail(HS_EDIM, "hs_solve_sweep: ldx must equal n");
  CUDA_OK(cudaSetDevice(f->ctx->device));
  if (f->dtype == HS_F64) sweep_impl<double>(f, nrhs, x, which); else sweep_impl<cplx>(f, nrhs, x, which);
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_refactor(hs_fac* f, const void* nzval, int32_t on_device) {
  HS_TRY_BEGIN
  if (!f) return hs_fail(HS_EARG, "hs_refactor: null argument");
  CUDA_OK(cudaSetDevice(f->ctx->device));
  if (nzval && nzval != f->d_nzval) {
    CUDA_OK(cudaMemcpyAsync(f->d_nzval, nzval, (size_t)f->nnz * f->esz, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                            f->ctx->stream));
    // the CSR image hs_gmres / hs_spmv build on first use holds the OLD values: mark it stale (its values are gathered
    // again on next use) so that the operator of the Krylov loop and the preconditioner keep describing the same matrix
    f->csr_stale = true;
  }
  if (f->dtype == HS_F64) numeric<double>(f); else numeric<cplx>(f);
  if (f->stats.singular_col >= 0) return hs_fail(HS_ESINGULAR, "hs_refactor: exactly singular pivot block");
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_factor_free(hs_fac* f) {
  if (f) { cudaSetDevice(f->ctx->device); delete f; }
  return HS_OK;
}

// ------------------------------------------------------------------------------------------------
// hs_solve
// ------------------------------------------------------------------------------------------------
template <typename T> static void solve_impl(hs_fac* f, int64_t nrhs, const void* B, int64_t ldb, void* X, int64_t ldx, int on_device) {
  cudaStream_t st = f->ctx->stream;
  if (nrhs > f->rhs_cap) {
    cudaFree(f->d_x); cudaFree(f->d_work);
    f->d_x = f->d_work = nullptr;
    CUDA_OK(cudaMalloc(&f->d_x, (size_t)f->xld * nrhs * sizeof(T)));
    CUDA_OK(cudaMalloc(&f->d_work, std::max<size_t>((size_t)f->max_level_idx * nrhs, 1) * sizeof(T)));
    f->rhs_cap = nrhs;
  }
  T* x = (T*)f->d_x;
  const cudaMemcpyKind kin = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind kout = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CUDA_OK(cudaMemcpy2DAsync(x, (size_t)f->xld * sizeof(T), B, (size_t)ldb * sizeof(T), (size_t)f->n * sizeof(T), nrhs, kin, st));
  if (f->nvirt)  // border rows of the thin fronts start from zero
    CUDA_OK(cudaMemset2DAsync(x + f->n, (size_t)f->xld * sizeof(T), 0, (size_t)f->nvirt * sizeof(T), nrhs, st));
  CUDA_OK(cudaEventRecord(f->ev0, st));
  hs_stats_t& s = f->stats;
  s.launches_solve = 0;
  hs_solve_run(f, nrhs, x, 3);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(f->ev1, st));
  f->ctx->launches += s.launches_solve;
  CUDA_OK(cudaMemcpy2DAsync(X, (size_t)ldx * sizeof(T), x, (size_t)f->xld * sizeof(T), (size_t)f->n * sizeof(T), nrhs, kout, st));
  CUDA_OK(cudaStreamSynchronize(st));
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, f->ev0, f->ev1));
  s.ms_solve_total = ms;
}

extern "C" int32_t hs_solve(hs_fac* f, int64_t nrhs, const void* B, int64_t ldb, void* X, int64_t ldx, int32_t on_device) {
  HS_TRY_BEGIN
  if (!f || !B || !X) return hs_fail(HS_EARG, "hs_solve: null argument");
  if (nrhs <= 0) return HS_OK;
  if (ldb < f->n || ldx < f->n) return hs_fail(HS_EDIM, "hs_solve: leading dimension smaller than n");
  CUDA_OK(cudaSetDevice(f->ctx->device));
  if (f->dtype == HS_F64) solve_impl<double>(f, nrhs, B, ldb, X, ldx, on_device);
  else solve_impl<cplx>(f, nrhs, B, ldb, X, ldx, on_device);
  return HS_OK;
  HS_TRY_END
}

// ------------------------------------------------------------------------------------------------
// introspection: FactorNode fields as the reference defines them (host-side reconstruction from the front)
// ------------------------------------------------------------------------------------------------
template <typename T> static void node_get_impl(hs_fac* f, int64_t node, hs_which which, void* out, int64_t* dims);

// compressed node: D, L = Qb·(Rb·Aii⁻¹), R = (Aii⁻¹·Qi)·Ri from the thin front (as an uncompressed front with r
// boundary rows), S from the dense slot
template <typename T> static void node_get_compressed(hs_fac* f, int64_t node, const CompFront& cf, hs_which which, void* out,
                                                      int64_t* dims) {
  const int fi = cf.fi, ni = cf.ni, nb = cf.nb;
  cudaStream_t st = f->ctx->stream;
  T* o = (T*)out;
  if (which == HS_GET_S && cf.hss >= 0) {
    // F.S is an HssMatrix (factorization.jl:110-111): the dense matrix it represents, S[perm,perm] order
    const HssFront& H = f->hss[cf.hss];
    if (dims) { dims[0] = H.m; dims[1] = H.m; }
    if (out) hs_hss_dense(f, cf.hss, out);
    return;
  }
  if (which == HS_GET_S && f->transient_schur && f->node2front[node] != f->root_front)
    throw hs_error(HS_EARG, "hs_node_get: the dense Schur block of a compressed front is transient (its slot is reused two levels up); "
                            "set HS_KEEP_SCHUR=1 before hs_factor to keep it");
  if (which == HS_GET_S || which == HS_GET_FRONT || which == HS_GET_PIV || which == HS_GET_D) {
    // S lives in the dense slot; D, the raw front and the pivots are those of the thin front
    const bool thin = which != HS_GET_S;
    std::swap(f->fronts[fi], f->fronts[f->nfr + fi]);  // node_get_impl reads fronts[node2front[node]]
    if (!thin) std::swap(f->fronts[fi], f->fronts[f->nfr + fi]);
    std::vector<CompFront> keep;
    keep.swap(f->comp);
    try { node_get_impl<T>(f, node, which, out, dims); } catch (...) {
      keep.swap(f->comp);
      if (thin) std::swap(f->fronts[fi], f->fronts[f->nfr + fi]);
      throw;
    }
    keep.swap(f->comp);
    if (thin) std::swap(f->fronts[fi], f->fronts[f->nfr + fi]);
    return;
  }
  if (which != HS_GET_L && which != HS_GET_R) throw hs_error(HS_EARG, "hs_node_get: unknown field");
  const bool isL = which == HS_GET_L;
  if (dims) { dims[0] = isL ? nb : ni; dims[1] = isL ? ni : nb; }
  if (!out) return;
  // thin-front quantity through the uncompressed code path: Lt = (Rb·U11⁻¹)·L11⁻¹·P (r×ni), Rt = U11⁻¹·(L11⁻¹PQi) (ni×r)
  const int r = cf.r;
  std::vector<T> thinq((size_t)std::max(r, 1) * ni);
  {
    std::swap(f->fronts[fi], f->fronts[f->nfr + fi]);
    std::vector<CompFront> keep;
    keep.swap(f->comp);
    int64_t d2[2];
    try { node_get_impl<T>(f, node, which, thinq.data(), d2); } catch (...) {
      keep.swap(f->comp);
      std::swap(f->fronts[fi], f->fronts[f->nfr + fi]);
      throw;
    }
    keep.swap(f->comp);
    std::swap(f->fronts[fi], f->fronts[f->nfr + fi]);
  }
  const T* pool = (const T*)f->pool;
  if (isL) {
    std::vector<T> Qb((size_t)nb * std::max(cf.r1, 1));
    if (cf.r1) CUDA_OK(cudaMemcpy2DAsync(Qb.data(), (size_t)nb * sizeof(T), pool + cf.qb, (size_t)cf.qb_ld * sizeof(T),
                                         (size_t)nb * sizeof(T), cf.r1, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    for (int j = 0; j < ni; ++j)
      for (int i = 0; i < nb; ++i) {
        T acc = hs_zero<T>();
        for (int k = 0; k < cf.r1; ++k) acc = hs_fma(acc, Qb[(size_t)k * nb + i], thinq[(size_t)j * r + k]);
        o[(size_t)j * nb + i] = acc;
      }
  } else {
    std::vector<T> Ri((size_t)std::max(cf.r2, 1) * nb);
    if (cf.r2) CUDA_OK(cudaMemcpy2DAsync(Ri.data(), (size_t)cf.r2 * sizeof(T), pool + cf.ri, (size_t)cf.ri_ld * sizeof(T),
                                         (size_t)cf.r2 * sizeof(T), nb, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    for (int j = 0; j < nb; ++j)
      for (int i = 0; i < ni; ++i) {
        T acc = hs_zero<T>();
        for (int k = 0; k < cf.r2; ++k) acc = hs_fma(acc, thinq[(size_t)k * ni + i], Ri[(size_t)j * cf.r2 + k]);
        o[(size_t)j * ni + i] = acc;
      }
  }
}

template <typename T> static void node_get_impl(hs_fac* f, int64_t node, hs_which which, void* out, int64_t* dims) {
  const Front& fd = f->fronts[f->node2front[node]];
  const CompFront* cf = nullptr;
  for (const CompFront& q : f->comp)
    if (q.fi == f->node2front[node]) cf = &q;
  if (cf) { node_get_compressed<T>(f, node, *cf, which, out, dims); return; }
  const Front& fr = fd;
  const int n = fr.n, ni = fr.ni, nb = n - ni;
  int64_t r = 0, c = 0;
  switch (which) {
    case HS_GET_D: r = ni; c = ni; break;
    case HS_GET_S: r = nb; c = nb; break;
    case HS_GET_L: r = nb; c = ni; break;
    case HS_GET_R: r = ni; c = nb; break;
    case HS_GET_FRONT: r = n; c = n; break;
    case HS_GET_PIV: r = ni; c = 1; break;
    default: throw hs_error(HS_EARG, "hs_node_get: unknown field");
  }
  if (dims) { dims[0] = r; dims[1] = c; }
  if (!out) return;
  cudaStream_t st = f->ctx->stream;
  std::vector<T> Fh((size_t)n * n);
  std::vector<int> piv(ni);
  if (n) CUDA_OK(cudaMemcpy2DAsync(Fh.data(), (size_t)n * sizeof(T), (T*)f->pool + fr.off, (size_t)fr.ld * sizeof(T), (size_t)n * sizeof(T), n,
                                   cudaMemcpyDeviceToHost, st));
  // The LAPACK interchange sequence is replayed from the pivot order (rperm[k] = original row that ends at row k):
  // every kernel family records rperm, the row-per-thread small-front kernel records nothing else.
  auto interchanges = [&](const Front& ff, int m, std::vector<int>& pv) {
    std::vector<int> rp(m), what(m), where(m);
    if (m) CUDA_OK(cudaMemcpyAsync(rp.data(), f->d_rperm + ff.ioff, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    pv.resize(m);
    for (int k = 0; k < m; ++k) what[k] = where[k] = k;
    for (int k = 0; k < m; ++k) {
      const int pr = rp[k], q = where[pr], other = what[k];   // row pr sits at position q ≥ k; swap positions k and q
      pv[k] = q;
      what[k] = pr; what[q] = other;
      where[pr] = k; where[other] = q;
    }
  };
  interchanges(fr, ni, piv);
  auto Fm = [&](int i, int j) -> T& { return Fh[(size_t)j * n + i]; };
  T* o = (T*)out;
  auto undo_prep = [&](int off, int m) {
    // the solve preparation replaced the DB×DB diagonal blocks of L11/U11 by their inverses: invert them back
    const int DB = hs_solve_block(f->dtype);
    for (int b0 = 0; b0 < m; b0 += DB) {
      const int db = std::min(DB, m - b0);
      std::vector<T> X((size_t)db * db), Y((size_t)db * db, hs_zero<T>());
      for (int j = 0; j < db; ++j) for (int i = 0; i < db; ++i) X[(size_t)j * db + i] = Fm(off + b0 + i, off + b0 + j);
      // L = (Linv)⁻¹, unit lower: forward substitution column by column
      for (int j = 0; j < db; ++j)
        for (int i = j + 1; i < db; ++i) {
          T sacc = X[(size_t)j * db + i];
          for (int k = j + 1; k < i; ++k) sacc = hs_fma(sacc, X[(size_t)k * db + i], Y[(size_t)j * db + k]);
          Y[(size_t)j * db + i] = hs_sub(hs_zero<T>(), sacc);
        }
      // U = (Uinv)⁻¹, upper: back substitution column by column
      for (int j = 0; j < db; ++j) {
        Y[(size_t)j * db + j] = hs_recip(X[(size_t)j * db + j]);
        for (int i = j - 1; i >= 0; --i) {
          T sacc = hs_zero<T>();
          for (int k = i + 1; k <= j; ++k) sacc = hs_fma(sacc, X[(size_t)k * db + i], Y[(size_t)j * db + k]);
          Y[(size_t)j * db + i] = hs_sub(hs_zero<T>(), hs_mul(sacc, hs_recip(X[(size_t)i * db + i])));
        }
      }
      for (int j = 0; j < db; ++j) for (int i = 0; i < db; ++i) Fm(off + b0 + i, off + b0 + j) = Y[(size_t)j * db + i];
    }
  };
  if (which != HS_GET_PIV) {
    undo_prep(0, ni);
    if (f->pseudo_front >= 0 && f->node2front[node] == f->root_front) undo_prep(ni, nb);
  }
  if (which == HS_GET_FRONT) { std::memcpy(out, Fh.data(), Fh.size() * sizeof(T)); return; }
  if (which == HS_GET_PIV) { int64_t* po = (int64_t*)out; for (int i = 0; i < ni; ++i) po[i] = piv[i]; return; }
  if (which == HS_GET_S) {
    // S[perm,perm], perm = [int_loc; bnd_loc] (factorization.jl:39-41,73-74); rows the parent drops are omitted
    std::vector<int> perm;
    for (int64_t q = f->iloc_ptr[node]; q < f->iloc_ptr[node + 1]; ++q) perm.push_back((int)f->iloc_idx[q]);
    for (int64_t q = f->bloc_ptr[node]; q < f->bloc_ptr[node + 1]; ++q) perm.push_back((int)f->bloc_idx[q]);
    const int np = (int)perm.size();
    if (dims) { dims[0] = np; dims[1] = np; }
    std::vector<T> Sd;
    if (f->pseudo_front >= 0 && f->node2front[node] == f->root_front && nb > 0) {
      // the root's Schur block was LU-factored in place for the boundary solve (factornode.jl:72): S = Pᵀ·L·U
      const Front& pf = f->fronts[f->pseudo_front];
      std::vector<int> pv;
      interchanges(pf, nb, pv);
      Sd.assign((size_t)nb * nb, hs_zero<T>());
      for (int j = 0; j < nb; ++j)
        for (int k = 0; k <= j; ++k) {
          const T u = Fm(ni + k, ni + j);
          Sd[(size_t)j * nb + k] = hs_add(Sd[(size_t)j * nb + k], u);
          for (int i = k + 1; i < nb; ++i) Sd[(size_t)j * nb + i] = hs_fma(Sd[(size_t)j * nb + i], Fm(ni + i, ni + k), u);
        }
      for (int k = nb - 1; k >= 0; --k)
        if (pv[k] != k)
          for (int j = 0; j < nb; ++j) std::swap(Sd[(size_t)j * nb + k], Sd[(size_t)j * nb + pv[k]]);
    }
    auto Sm = [&](int i, int j) -> T { return Sd.empty() ? Fm(ni + i, ni + j) : Sd[(size_t)j * nb + i]; };
    for (int j = 0; j < np; ++j)
      for (int i = 0; i < np; ++i) o[(size_t)j * np + i] = Sm(perm[i], perm[j]);
    return;
  }
  if (which == HS_GET_D) {
    // A_ii = Pᵀ·L11·U11
    std::vector<T> M((size_t)ni * ni, hs_zero<T>());
    for (int j = 0; j < ni; ++j)
      for (int k = 0; k <= j; ++k) {
        const T u = Fm(k, j);
        M[(size_t)j * ni + k] = hs_add(M[(size_t)j * ni + k], u);  // unit diagonal of L
        for (int i = k + 1; i < ni; ++i) M[(size_t)j * ni + i] = hs_fma(M[(size_t)j * ni + i], Fm(i, k), u);
      }
    for (int k = ni - 1; k >= 0; --k)
      if (piv[k] != k)
        for (int j = 0; j < ni; ++j) std::swap(M[(size_t)j * ni + k], M[(size_t)j * ni + piv[k]]);
    std::memcpy(out, M.data(), M.size() * sizeof(T));
    return;
  }
  if (which == HS_GET_R) {
    // R = A_ii⁻¹·A_ib = U11⁻¹·U12
    for (int j = 0; j < nb; ++j) {
      T* x = o + (size_t)j * ni;
      for (int i = 0; i < ni; ++i) x[i] = Fm(i, ni + j);
      for (int k = ni - 1; k >= 0; --k) {
        x[k] = hs_mul(x[k], hs_recip(Fm(k, k)));
        for (int i = 0; i < k; ++i) x[i] = hs_fnma(x[i], Fm(i, k), x[k]);
      }
    }
    return;
  }
  if (which == HS_GET_L) {
    // L = A_bi·A_ii⁻¹ = L21·L11⁻¹·P : solve X·L11 = L21 (unit lower), then undo the interchanges on the columns
    std::vector<T> X((size_t)nb * ni);
    for (int k = 0; k < ni; ++k)
      for (int i = 0; i < nb; ++i) X[(size_t)k * nb + i] = Fm(ni + i, k);
    for (int k = ni - 1; k >= 0; --k)
      for (int j = 0; j < k; ++j) {
        const T l = Fm(k, j);
        for (int i = 0; i < nb; ++i) X[(size_t)j * nb + i] = hs_fnma(X[(size_t)j * nb + i], X[(size_t)k * nb + i], l);
      }
    for (int k = ni - 1; k >= 0; --k)
      if (piv[k] != k)
        for (int i = 0; i < nb; ++i) std::swap(X[(size_t)k * nb + i], X[(size_t)piv[k] * nb + i]);
    std::memcpy(out, X.data(), X.size() * sizeof(T));
    return;
  }
}

extern "C" int32_t hs_node_get(hs_fac* f, int64_t node, hs_which which, void* out, int64_t* dims) {
  HS_TRY_BEGIN
  if (!f) return hs_fail(HS_EARG, "hs_node_get: null argument");
  if (node < 0 || node >= f->nnodes) return hs_fail(HS_EARG, "hs_node_get: node out of range");
  CUDA_OK(cudaSetDevice(f->ctx->device));
  if (f->dtype == HS_F64) node_get_impl<double>(f, node, which, out, dims);
  else node_get_impl<cplx>(f, node, which, out, dims);
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_maxrank(hs_fac* f, int64_t* rank) {
  if (!f || !rank) return hs_fail(HS_EARG, "hs_maxrank: null argument");
  *rank = f->stats.maxrank;  // 0 when nothing is compressed (factornode.jl:49-57)
  return HS_OK;
}

extern "C" int32_t hs_node_rank(hs_fac* f, int64_t node, int64_t* rank_l, int64_t* rank_r) {
  if (!f || !rank_l || !rank_r) return hs_fail(HS_EARG, "hs_node_rank: null argument");
  if (node < 0 || node >= f->nnodes) return hs_fail(HS_EARG, "hs_node_rank: node out of range");
  *rank_l = *rank_r = 0;
  const int fi = f->node2front[node];
  for (const CompFront& cf : f->comp)
    if (cf.fi == fi) { *rank_l = cf.r1; *rank_r = cf.r2; }
  return HS_OK;
}

static const CompFront* comp_of(hs_fac* f, int64_t node) {
  const int fi = f->node2front[node];
  for (const CompFront& cf : f->comp)
    if (cf.fi == fi) return &cf;
  return nullptr;
}

extern "C" int32_t hs_hss_info(hs_fac* f, int64_t node, int64_t* nhss, int64_t* info) {
  HS_TRY_BEGIN
  if (!f || !nhss) return hs_fail(HS_EARG, "hs_hss_info: null argument");
  if (node < 0 || node >= f->nnodes) return hs_fail(HS_EARG, "hs_hss_info: node out of range");
  *nhss = 0;
  const CompFront* cf = comp_of(f, node);
  if (!cf || cf->hss < 0) return HS_OK;
  const HssFront& H = f->hss[cf->hss];
  *nhss = (int64_t)H.tree.size();
  if (info)
    for (size_t t = 0; t < H.tree.size(); ++t) {
      const HssTreeNode& tn = H.tree[t];
      int64_t* o = info + 8 * t;
      o[0] = tn.lo; o[1] = tn.hi; o[2] = tn.left; o[3] = tn.right; o[4] = H.st[t].r0; o[5] = H.st[t].r1; o[6] = tn.parent; o[7] = tn.left < 0;
    }
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_hss_get(hs_fac* f, int64_t node, int64_t hnode, hs_hss_which which, void* out, int64_t* dims) {
  HS_TRY_BEGIN
  if (!f || !dims) return hs_fail(HS_EARG, "hs_hss_get: null argument");
  if (node < 0 || node >= f->nnodes) return hs_fail(HS_EARG, "hs_hss_get: node out of range");
  const CompFront* cf = comp_of(f, node);
  if (!cf || cf->hss < 0) return hs_fail(HS_EARG, "hs_hss_get: the Schur complement of this node is not an HSS matrix");
  const HssFront& H = f->hss[cf->hss];
  if (hnode < 0 || hnode >= (int64_t)H.tree.size()) return hs_fail(HS_EARG, "hs_hss_get: HSS node out of range");
  const HssTreeNode& tn = H.tree[hnode];
  const HssStored& S = H.st[hnode];
  const bool leaf = tn.left < 0;
  const int m = tn.hi - tn.lo;
  long long src = 0; int ld = 0, rows = 0, cols = 0; bool ct = false, have = false;   // ct: stored conjugate-transposed
  int ra0 = 0, ra1 = 0, rb0 = 0, rb1 = 0;
  if (!leaf) { ra0 = H.st[tn.left].r0; ra1 = H.st[tn.left].r1; rb0 = H.st[tn.right].r0; rb1 = H.st[tn.right].r1; }
  switch (which) {
    case HS_HSS_D: if (leaf) { have = true; src = S.D; ld = S.ldD; rows = cols = m; } break;
    case HS_HSS_U: if (leaf) { have = true; src = S.U; ld = S.ldU; rows = m; cols = S.r0; } break;
    case HS_HSS_V: if (leaf) { have = true; src = S.VH; ld = S.ldVH; rows = m; cols = S.r1; ct = true; } break;
    case HS_HSS_B12: if (!leaf) { have = true; src = S.B12; ld = S.ldB12; rows = ra0; cols = rb1; } break;
    case HS_HSS_B21: if (!leaf) { have = true; src = S.B21; ld = S.ldB21; rows = rb0; cols = ra1; } break;
    case HS_HSS_R: if (!leaf && tn.parent >= 0) { have = true; src = S.R; ld = S.ldR; rows = ra0 + rb0; cols = S.r0; } break;
    case HS_HSS_W: if (!leaf && tn.parent >= 0) { have = true; src = S.WH; ld = S.ldWH; rows = ra1 + rb1; cols = S.r1; ct = true; } break;
    default: return hs_fail(HS_EARG, "hs_hss_get: unknown generator");
  }
  if (!have) return hs_fail(HS_EARG, "hs_hss_get: this HSS node has no such generator");
  src += H.sbase;
  dims[0] = rows; dims[1] = cols;
  if (!out || rows == 0 || cols == 0) return HS_OK;
  CUDA_OK(cudaSetDevice(f->ctx->device));
  const size_t esz = f->esz;
  const int srows = ct ? cols : rows, scols = ct ? rows : cols;
  std::vector<char> tmp((size_t)srows * scols * esz);
  CUDA_OK(cudaMemcpy2D(tmp.data(), (size_t)srows * esz, (const char*)f->pool + src * (long long)esz, (size_t)ld * esz, (size_t)srows * esz, scols,
                       cudaMemcpyDeviceToHost));
  if (!ct) { std::memcpy(out, tmp.data(), tmp.size()); return HS_OK; }
  const int nw = (int)(esz / 8);
  const double* in = (const double*)tmp.data();
  double* o = (double*)out;
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i) {   // out[i, j] = conj(stored[j, i])
      const double* e = in + ((size_t)i * srows + j) * nw;
      double* d = o + ((size_t)j * rows + i) * nw;
      d[0] = e[0];
      if (nw == 2) d[1] = -e[1];
    }
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_stats(hs_fac* f, hs_stats_t* out) {
  if (!f || !out) return hs_fail(HS_EARG, "hs_stats: null argument");
  *out = f->stats;
  return HS_OK;
}

extern "C" int32_t hs_matrix_device(hs_fac* f, const int64_t** colptr, const int64_t** rowval, const void** nzval, int64_t* nnz) {
  if (!f || !colptr || !rowval || !nzval || !nnz) return hs_fail(HS_EARG, "hs_matrix_device: null argument");
  *colptr = (const int64_t*)f->d_colptr;
  *rowval = (const int64_t*)f->d_rowval;
  *nzval = f->d_nzval;
  *nnz = f->nnz;
  return HS_OK;
}

extern "C" int32_t hs_resolved_swlevel(hs_fac* f, int64_t* sw) {
  if (!f || !sw) return hs_fail(HS_EARG, "hs_resolved_swlevel: null argument");
  *sw = f->swlevel_resolved;
  return HS_OK;
}
