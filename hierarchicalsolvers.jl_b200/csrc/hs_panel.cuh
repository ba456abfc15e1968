// LU panel kernel and its launcher; instantiated per scalar type in hs_panel_f64.cu / hs_panel_c64.cu.
#pragma once

#include <cooperative_groups.h>

#include <cstdlib>

#include "hs_fac.cuh"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------
// LU panel with partial pivoting restricted to the pivot block rows (rows < ni)
// ------------------------------------------------------------------------------------------------
// Panel = front rows [j0, plim) × columns [j0, j0+wc), plim = end of the diagonal block pivots are taken from (ni, or
// the split of a pivot block eliminated as 2×2 blocks).  The rows below — the rest of the pivot block behind a split and
// all boundary rows — take no part in the pivot search: they are X·U_pp = B triangular solves done by k_trsm_rows with
// one thread per row, so the latency-bound cluster spans the pivot rows only (half the CTAs or less at the upper levels).
// The 256 threads of a CTA form a 32×8 grid; thread (tr, tc)
// keeps rows tr + 32·i (i < 8·R) and columns tc + 8·k (k < W/8) of the CTA's row block in registers, a cluster of C
// CTAs covers 256·R·C rows.  Per pivot column:
//   (a) the 32 threads that own the column publish it and |.| of the eligible rows to shared memory      → barrier
//   (b) every warp finds the CTA's best row redundantly; the 8 threads owning that row publish it
//   (c) clusters only: every CTA pushes its candidate and candidate row into all CTAs' mailboxes (st.async counting
//       bytes on the receiver's mbarrier); each CTA waits for its own mailbox, then every warp finds the winner in
//       local shared memory — no cluster barrier and no fence inside the column loop
//   (d) rank-1 update from registers; the pivot row is written out by its owner and frozen.
// Pivoting is implicit: rows stay where they are until CTA 0 moves them to their LAPACK positions at the end (the
// interchange sequence `ipiv` is replayed from the pivot order).  Column indices are static in the unrolled outer
// loop (k = j / 8), so the body is small and stays in the instruction cache.
template <typename T> struct PanelW;  // widest register tile per scalar type
template <> struct PanelW<double> { static constexpr int W0 = 64; };
template <> struct PanelW<cplx> { static constexpr int W0 = 32; };

struct PanelCand {
  double val;
  int row;
  int cta;
};

__device__ __forceinline__ unsigned hs_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ unsigned hs_mapa(unsigned a, int rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
// 8-byte remote store that also counts 8 bytes on the destination CTA's mbarrier (no fence, no cluster barrier)
__device__ __forceinline__ void hs_st_async64(unsigned dst, unsigned long long v, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst), "l"(v), "r"(mbar) : "memory");
}

template <typename T, int W, int R, bool CL>
__global__ void __launch_bounds__(256, 1) k_panel(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                   int* __restrict__ ipiv, int* __restrict__ rperm, int f0, int j0,
                                                   int* __restrict__ info) {
  constexpr int NT = 256, TR = 32, TC = 8;
  constexpr int RPT = 8 * R, CPT = W / TC, ROWS = TR * RPT;
  static_assert(W % TC == 0 && RPT <= 64, "unsupported panel shape");
  cg::cluster_group cluster = cg::this_cluster();
  const int C = CL ? (int)cluster.num_blocks() : 1;
  const int crank = CL ? (int)cluster.block_rank() : 0;
  const int fi = f0 + (CL ? blockIdx.x / C : blockIdx.x);
  const Front fr = fronts[fi];
  if (fr.ni <= j0) return;  // uniform across the cluster
  const int plim = hs_plim(fr, j0);
  const int wc = min(W, plim - j0);
  const int m = plim - j0;
  const int tid = threadIdx.x, lane = tid & 31;
  const int tr = tid % TR, tc = tid / TR;
  const int rbase = crank * ROWS;
  T* F = pool + fr.off;
  const long long ld = fr.ld;

  extern __shared__ __align__(16) unsigned char smem_dyn[];
  T* s_col = reinterpret_cast<T*>(smem_dyn);                 // [2][ROWS]   column j of this CTA's rows
  double* s_abs = reinterpret_cast<double*>(s_col + 2 * ROWS);  // [2][ROWS] |.| of eligible rows, -1 otherwise
  T* stage = s_col;                                          // reused after the loop: staging of the row moves
  constexpr int CMAX = CL ? 16 : 1;
  __shared__ T s_crow[2][W];
  // clusters: every CTA PUSHES its candidate and candidate row into these mailboxes of all CTAs before the cluster
  // barrier, so that after the barrier the winner is found from local shared memory (no DSMEM round trip left on
  // the per-column critical path)
  __shared__ __align__(16) PanelCand s_allc[2][CMAX];
  __shared__ __align__(16) T s_allrow[2][CMAX][W];
  __shared__ __align__(8) unsigned long long s_mbar[2];  // one mailbox barrier per column parity
  if (CL) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(hs_smem_u32(&s_mbar[0])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(hs_smem_u32(&s_mbar[1])));
      asm volatile("fence.mbarrier_init.release.cluster;");
    }
    cluster.sync();  // every CTA's barriers exist before anyone stores into them
  }
  __shared__ int s_pivrow[W];
  __shared__ int s_what[W], s_where[W], s_isp[W];
  __shared__ int s_mdst[2 * W], s_msrc[2 * W], s_nmv;

  T a[RPT][CPT];
  unsigned long long done = 0;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = rbase + tr + TR * i;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tc + TC * k;
      a[i][k] = (r < m && c < wc) ? F[(long long)(j0 + c) * ld + (j0 + r)] : hs_zero<T>();
    }
  }

#pragma unroll
  for (int kj = 0; kj < CPT; ++kj) {
    for (int jt = 0; jt < TC; ++jt) {
      const int j = kj * TC + jt;
      if (j >= wc) break;
      const int par = j & 1;
      T* colp = s_col + par * ROWS;
      double* absp = s_abs + par * ROWS;
      // (a) the owners of column j publish it
      if (tc == jt) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int rl = tr + TR * i, r = rbase + rl;
          colp[rl] = a[i][kj];
          absp[rl] = (!((done >> i) & 1ull) && r < m) ? hs_abs1(a[i][kj]) : -1.0;
        }
      }
      __syncthreads();
      // (b) every warp finds the CTA's candidate (largest |.|, smallest row on ties)
      double cb = -1.0;
      int crl = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < ROWS / 32; ++q) {
        const int rl = lane + 32 * q;
        const double v = absp[rl];
        if (v > cb) { cb = v; crl = rl; }
      }
      warp_argmax(cb, crl);
      if (cb >= 0.0 && tr == crl % TR) {  // the 8 threads that own the candidate row publish it
        const int ip = crl / TR;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          if (i == ip) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) s_crow[par][tc + TC * k] = a[i][k];
          }
        }
      }
      double gb;
      int p, gc;
      const T* urow;
      if (CL) {
        __syncthreads();  // s_crow[par] is complete
        // (c) push: every CTA stores its candidate row and candidate into all CTAs' mailboxes with st.async, which
        // counts the bytes on the destination's mbarrier; a CTA goes on as soon as ITS mailbox is full.  Unlike
        // barrier.cluster this needs no release fence — which would wait for the global stores of step (d) — and
        // costs ~360 instead of ~1000 cycles for 16 CTAs (tools/microbench/cluster_sync.cu).
        {
          constexpr int NWORD = W * (int)sizeof(T) / 8;
          const unsigned long long* srcw = reinterpret_cast<const unsigned long long*>(&s_crow[par][0]);
          const unsigned row0 = hs_smem_u32(&s_allrow[par][crank][0]), mb0 = hs_smem_u32(&s_mbar[par]);
          for (int e = tid; e < C * NWORD; e += NT) {
            const int dst = e / NWORD, k = e % NWORD;
            hs_st_async64(hs_mapa(row0 + 8u * k, dst), srcw[k], hs_mapa(mb0, dst));
          }
          if (tid < C) {
            const unsigned c0 = hs_mapa(hs_smem_u32(&s_allc[par][crank]), tid), mbr = hs_mapa(mb0, tid);
            const int row = cb >= 0.0 ? rbase + crl : 0x7fffffff;
            hs_st_async64(c0, (unsigned long long)__double_as_longlong(cb), mbr);
            hs_st_async64(c0 + 8u, (unsigned long long)(unsigned)row, mbr);
          }
          if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb0), "r"(C * (NWORD * 8 + 16)) : "memory");
          unsigned ok = 0;
          const unsigned phase = (unsigned)(j >> 1) & 1u;
          while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(mb0), "r"(phase) : "memory");
        }
        // every warp picks the winner from its own CTA's mailboxes
        double wb = -1.0;
        int wr = 0x7fffffff;
        if (lane < C) { wb = s_allc[par][lane].val; wr = s_allc[par][lane].row; }
        const int myrow = wr;  // rows are disjoint between CTAs, so the winning row identifies its CTA
        warp_argmax(wb, wr);
        int wcid = (int)__reduce_min_sync(0xffffffffu, (wb >= 0.0 && myrow == wr) ? (unsigned)lane : 0xffffffffu);
        if (wcid < 0 || wcid >= C) wcid = 0;  // no candidate anywhere (cannot happen while j < wc): stay in range
        gb = wb; p = wr; gc = wcid;
        urow = &s_allrow[par][wcid][0];
      } else {
        __syncthreads();
        gb = cb; p = crl; gc = 0;
        urow = &s_crow[par][0];
      }
      // (d) the pivot row is final: its CTA writes it out (physical position; moved at the end) and freezes it
      if (gc == crank && gb >= 0.0) {
        if (tid < wc) F[(long long)(j0 + tid) * ld + (j0 + p)] = urow[tid];
        const int pl = p - rbase;
        if (tr == pl % TR) done |= 1ull << (pl / TR);
      }
      if (crank == 0 && tid == 0) {
        s_pivrow[j] = p;
        if (!(gb > 0.0) && atomicCAS(&info[0], 0, 1) == 0) { info[1] = fi; info[2] = j0 + j; }
      }
      if (gb > 0.0) {  // an exactly singular column is recorded and skipped, as LAPACK getf2 does
        // No per-row predicates here: frozen pivot rows are never searched or stored again, so letting the update
        // run over their registers is harmless, and rows beyond the front hold zeros.
        const T inv = hs_recip_pivot(urow[j]);
        T l[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) l[i] = hs_mul(colp[tr + TR * i], inv);
        if (tc == jt) {  // warp-uniform: this warp owns column j and keeps the multipliers
#pragma unroll
          for (int i = 0; i < RPT; ++i) a[i][kj] = l[i];
        }
#pragma unroll
        for (int k = kj; k < CPT; ++k) {
          if (k == kj && tc <= jt) continue;
          const T u = urow[tc + TC * k];
#pragma unroll
          for (int i = 0; i < RPT; ++i) a[i][k] = hs_fnma(a[i][k], l[i], u);
        }
      }
    }
  }
  // live rows go back to global memory (frozen pivot rows were written when they were chosen)
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = rbase + tr + TR * i;
    if (((done >> i) & 1ull) || r >= m) continue;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tc + TC * k;
      if (c < wc) F[(long long)(j0 + c) * ld + (j0 + r)] = a[i][k];
    }
  }
  // every row of the panel is in global memory at its physical position; CTA 0 moves rows to their LAPACK places
  if (CL) cluster.sync(); else __syncthreads();
  if (crank != 0) return;
  if (tid == 0) {
    // replay the interchanges: at step k the row chosen as pivot sits at position q → swap positions k and q
    for (int k = 0; k < wc; ++k) { s_what[k] = k; s_where[k] = k; s_isp[k] = 0; }
    for (int k = 0; k < wc; ++k) {
      const int pr = s_pivrow[k];
      const int q = pr < wc ? s_where[pr] : pr;  // rows from below the top block have not moved before
      ipiv[fr.ioff + j0 + k] = j0 + q;
      const int other = s_what[k];
      s_what[k] = pr;
      if (q < wc) s_what[q] = other;
      if (pr < wc) { s_where[pr] = k; s_isp[pr] = 1; }
      if (other < wc) s_where[other] = q;
    }
    int nm = 0;
    for (int k = 0; k < wc; ++k)
      if (s_pivrow[k] != k) { s_mdst[nm] = k; s_msrc[nm] = s_pivrow[k]; ++nm; }
    for (int r = 0; r < wc; ++r)
      if (!s_isp[r] && s_where[r] != r) { s_mdst[nm] = s_where[r]; s_msrc[nm] = r; ++nm; }
    s_nmv = nm;
  }
  __syncthreads();
  const int nm = s_nmv;
  if (nm == 0) return;
  // rperm (original row of every position, identity before the first panel) takes the same moves as the rows: the
  // pivot order is complete when the last panel is, no serial replay of ipiv afterwards
  int rsave = 0;
  if (tid < nm) rsave = rperm[fr.ioff + j0 + s_msrc[tid]];
  for (int e = tid; e < nm * wc; e += NT) {
    const int i = e % nm, c = e / nm;
    stage[e] = F[(long long)(j0 + c) * ld + (j0 + s_msrc[i])];
  }
  __syncthreads();
  if (tid < nm) rperm[fr.ioff + j0 + s_mdst[tid]] = rsave;
  for (int e = tid; e < nm * wc; e += NT) {
    const int i = e % nm, c = e / nm;
    F[(long long)(j0 + c) * ld + (j0 + s_mdst[i])] = stage[e];
  }
}

template <typename T, int W, int R> constexpr size_t panel_smem() {
  // max(column/abs double buffers, staging of 2W moved rows × W columns)
  constexpr size_t colabs = 2 * (size_t)(32 * 8 * R) * (sizeof(T) + sizeof(double));
  constexpr size_t stg = 2 * (size_t)W * W * sizeof(T);
  return colabs > stg ? colabs : stg;
}

// ------------------------------------------------------------------------------------------------
// panel launch: pick the register tile (W, R) and the cluster size for the tallest active panel
// ------------------------------------------------------------------------------------------------
template <typename T, int W, int R>
static void launch_panel(hs_fac* f, int f0, int nact, int j0, int C, cudaStream_t st) {
  T* pool = (T*)f->pool;
  if (C == 1) {
    k_panel<T, W, R, false><<<nact, 256, panel_smem<T, W, R>(), st>>>(f->d_fronts, pool, f->d_ipiv, f->d_rperm, f0, j0, f->d_info);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nact * C));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = panel_smem<T, W, R>();
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const Front* fr = f->d_fronts;
    int* ipiv = f->d_ipiv;
    int* rperm = f->d_rperm;
    int* info = f->d_info;
    CUDA_OK(cudaLaunchKernelEx(&cfg, k_panel<T, W, R, true>, fr, pool, ipiv, rperm, f0, j0, info));
  }
  CUDA_OK(cudaGetLastError());
}

static int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// rows [plim, n) of panel [j0, j0+wc):  X·U_pp = B with U_pp the upper triangle of the factored wc×wc pivot block; one
// thread per row, the row in registers (coalesced over consecutive rows of a column)
template <typename T, int W>
__global__ void __launch_bounds__(128) k_trsm_rows(const Front* __restrict__ fronts, T* __restrict__ pool, int f0, int j0) {
  const Front fr = fronts[f0 + blockIdx.x];
  if (fr.ni <= j0) return;
  const int plim = hs_plim(fr, j0);
  const int wc = min(W, plim - j0);
  const int nrows = fr.n - plim;
  if ((int)(blockIdx.y * blockDim.x) >= nrows) return;
  T* F = pool + fr.off;
  __shared__ T sU[W * W];   // sU[k·W + c] = U[k, c] for k < c, 1/U[c, c] on the diagonal
  {
    // consecutive threads read consecutive rows of a column (coalesced), every load issued before the first use
    constexpr int NT = 128, NE = (W * W + NT - 1) / NT;
    T tmp[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = threadIdx.x + q * NT;
      const int k = e % W, c = e / W;
      tmp[q] = (e < W * W && k < wc && c < wc && k <= c) ? F[(long long)(j0 + c) * fr.ld + (j0 + k)] : hs_zero<T>();
    }
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = threadIdx.x + q * NT;
      const int k = e % W, c = e / W;
      if (e < W * W) sU[k * W + c] = (k == c && k < wc) ? hs_recip_pivot(tmp[q]) : tmp[q];
    }
  }
  __syncthreads();
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  T* row = F + plim + r;
  T x[W];
#pragma unroll
  for (int c = 0; c < W; ++c) x[c] = c < wc ? row[(long long)(j0 + c) * fr.ld] : hs_zero<T>();
  // right-looking: once x[c] is final it is eliminated from all later entries — W independent FMAs per step instead of a
  // dependent chain of c FMAs per entry
#pragma unroll
  for (int c = 0; c < W; ++c) {
    if (c < wc) {
      const T xc = hs_mul(x[c], sU[c * W + c]);
      x[c] = xc;
#pragma unroll
      for (int j = c + 1; j < W; ++j) x[j] = hs_fnma(x[j], xc, sU[c * W + j]);
    }
  }
#pragma unroll
  for (int c = 0; c < W; ++c)
    if (c < wc) row[(long long)(j0 + c) * fr.ld] = x[c];
}

template <typename T, int W> static void launch_trsm_rows(hs_fac* f, int f0, int nact, int j0, int max_rows, cudaStream_t st) {
  if (max_rows <= 0) return;
  dim3 grid(nact, (max_rows + 127) / 128);
  k_trsm_rows<T, W><<<grid, 128, 0, st>>>(f->d_fronts, (T*)f->pool, f0, j0);
  CUDA_OK(cudaGetLastError());
}
template <typename T> static void trsm_rows_dispatch(hs_fac* f, int W, int f0, int nact, int j0, int max_rows, cudaStream_t st) {
  constexpr int W0 = PanelW<T>::W0;
  if (W == W0) launch_trsm_rows<T, W0>(f, f0, nact, j0, max_rows, st);
  else if (W == W0 / 2) launch_trsm_rows<T, W0 / 2>(f, f0, nact, j0, max_rows, st);
  else if (W == W0 / 4) launch_trsm_rows<T, W0 / 4>(f, f0, nact, j0, max_rows, st);
  else if constexpr (W0 / 8 >= 8) launch_trsm_rows<T, W0 / 8>(f, f0, nact, j0, max_rows, st);
}

// panel width used for a level whose tallest panel has max_n PIVOT rows
template <typename T> static int choose_width(const hs_fac* f, int max_n, int nfronts) {
  const int W0 = PanelW<T>::W0;
  const int maxC = f->ctx->max_cluster;
  // many fronts per level (more CTAs than SMs): throughput matters, not the latency of one front.  A CTA that holds
  // more rows of a narrower panel avoids the cluster barrier altogether.  Opt-in (HS_PANEL_TALL=1): it shortens the
  // panel phase by ~8 ms at 2048² but the K = 16/32 in-block updates it implies cost the DMMA phase ~6 ms.
  static const bool tall = getenv("HS_PANEL_TALL") && atoi(getenv("HS_PANEL_TALL")) != 0;
  if (tall && (long long)nfronts * ((max_n + 255) / 256) > 148 && max_n > 256) {
    for (int W = W0, R = 1; R <= 4 && W >= 8; W >>= 1, R <<= 1)
      if (256 * R >= max_n) return W;
  }
  for (int W = W0, R = 1; R <= 8 && W >= 8; W >>= 1, R <<= 1)
    if ((long long)256 * R * maxC >= max_n) return W;
  return -1;
}

template <typename T>
static void panel_dispatch(hs_fac* f, int W, int f0, int nact, int j0, int m, cudaStream_t st) {
  constexpr int W0 = PanelW<T>::W0;
  const int R = W0 / W;
  const int C = pow2_ceil((m + 256 * R - 1) / (256 * R));
  if (W == W0) launch_panel<T, W0, 1>(f, f0, nact, j0, C, st);
  else if (W == W0 / 2) launch_panel<T, W0 / 2, 2>(f, f0, nact, j0, C, st);
  else if (W == W0 / 4) launch_panel<T, W0 / 4, 4>(f, f0, nact, j0, C, st);
  else if constexpr (W0 / 8 >= 8) launch_panel<T, W0 / 8, 8>(f, f0, nact, j0, C, st);
}


template <typename T, int W, int R> static void set_panel_attrs() {
  CUDA_OK(cudaFuncSetAttribute(k_panel<T, W, R, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CUDA_OK(cudaFuncSetAttribute(k_panel<T, W, R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem<T, W, R>()));
  CUDA_OK(cudaFuncSetAttribute(k_panel<T, W, R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem<T, W, R>()));
}
template <typename T> static void panel_setup() {
  constexpr int W0 = PanelW<T>::W0;
  set_panel_attrs<T, W0, 1>(); set_panel_attrs<T, W0 / 2, 2>(); set_panel_attrs<T, W0 / 4, 4>();
  if constexpr (W0 / 8 >= 8) set_panel_attrs<T, W0 / 8, 8>();
}
