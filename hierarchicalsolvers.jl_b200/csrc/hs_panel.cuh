// LU panel kernel and its launcher; instantiated per scalar type in hs_panel_f64.cu / hs_panel_c64.cu.
#pragma once

#include <cooperative_groups.h>

#include "hs_fac.cuh"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------
// LU panel with partial pivoting restricted to the pivot block rows (rows < ni)
// ------------------------------------------------------------------------------------------------
// Panel = front rows [j0, n) × columns [j0, j0+wc).  Every thread keeps R rows × W columns in registers; a cluster
// of C CTAs (256 threads each) covers 256·R·C rows.
//
// Per column: local |a| arg-max → the CTA's candidate row is parked in shared memory → ONE cluster barrier → warp 0
// of every CTA pulls the C candidates and the winning row through distributed shared memory → rank-1 update from
// registers.  Pivoting is IMPLICIT: a row chosen as pivot is frozen where it is (its owner writes it out and stops
// updating it); rows are moved to their LAPACK positions once, at the end, by CTA 0.  The column loop is rolled:
// CH columns are processed with static register indices, then every row is rotated left by CH registers, so the
// loop body stays ~20 KB of code (a fully unrolled panel was 1.5 MB and ran at instruction-fetch speed).
struct PanelCand {
  double val;
  int row;
  int pad;
};

template <typename T, int W, int R, bool CL>
__global__ void __launch_bounds__(256, 1) k_panel(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                   int* __restrict__ ipiv, int f0, int j0, int* __restrict__ info) {
  constexpr int NT = 256;
  constexpr int CH = 4;
  static_assert(W % CH == 0, "panel width must be a multiple of the chunk");
  cg::cluster_group cluster = cg::this_cluster();
  const int C = CL ? (int)cluster.num_blocks() : 1;
  const int crank = CL ? (int)cluster.block_rank() : 0;
  const int fi = f0 + (CL ? blockIdx.x / C : blockIdx.x);
  const Front fr = fronts[fi];
  if (fr.ni <= j0) return;  // uniform across the cluster
  const int wc = min(W, fr.ni - j0);
  const int m = fr.n - j0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T* F = pool + fr.off;
  const long long ld = fr.ld;

  __shared__ double s_wval[NT / 32];
  __shared__ int s_wrow[NT / 32];
  __shared__ PanelCand s_cand[2];
  __shared__ T s_crow[2][W];
  __shared__ T s_u[W];
  __shared__ PanelCand s_win;
  __shared__ int s_pivrow[W];
  __shared__ int s_what[W], s_where[W], s_isp[W];
  __shared__ int s_mdst[2 * W], s_msrc[2 * W], s_nmv;
  extern __shared__ __align__(16) unsigned char smem_dyn[];  // CTA 0: staging of the final row permutation

  T a[R][W];
  int rows[R];
  unsigned done = 0;
#pragma unroll
  for (int s = 0; s < R; ++s) {
    rows[s] = crank * (NT * R) + s * NT + tid;
    const bool ok = rows[s] < m;
#pragma unroll
    for (int k = 0; k < W; ++k) a[s][k] = (ok && k < wc) ? F[(long long)(j0 + k) * ld + (j0 + rows[s])] : hs_zero<T>();
  }

  for (int jc = 0; jc < wc; jc += CH) {
#pragma unroll
    for (int t = 0; t < CH; ++t) {
      const int j = jc + t;
      if (j < wc) {
        const int par = j & 1;
        // 1. thread-local then warp-level arg-max of |a(:,j)| over the rows still eligible (ties → smallest row)
        double best = -1.0;
        int brow = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < R; ++s) {
          if (!((done >> s) & 1u) && j0 + rows[s] < fr.ni) {
            const double v = hs_abs1(a[s][t]);
            if (v > best || (v == best && rows[s] < brow)) { best = v; brow = rows[s]; }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
          if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
        }
        if (lane == 0) { s_wval[warp] = best; s_wrow[warp] = brow; }
        __syncthreads();
        // 2. CTA-level candidate; its owner parks the whole register row (L entries, pivot, U entries)
        double cb = -1.0;
        int cr = 0x7fffffff;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
          const double ov = s_wval[w];
          const int orow = s_wrow[w];
          if (ov > cb || (ov == cb && orow < cr)) { cb = ov; cr = orow; }
        }
#pragma unroll
        for (int s = 0; s < R; ++s) {
          if (rows[s] == cr && cb >= 0.0) {
#pragma unroll
            for (int k = 0; k < W; ++k) s_crow[par][k] = a[s][k];
          }
        }
        if (tid == 0) { s_cand[par].val = cb; s_cand[par].row = cr; }
        // 3. one barrier per column
        if (CL) cluster.sync(); else __syncthreads();
        // 4. warp 0 pulls the candidates and the winning row
        if (warp == 0) {
          double gb = -1.0;
          int gr = 0x7fffffff, gc = 0;
          if (lane < C) {
            const PanelCand* rc = CL ? cluster.map_shared_rank(&s_cand[par], lane) : &s_cand[par];
            gb = rc->val; gr = rc->row; gc = lane;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, gb, o);
            const int orow = __shfl_xor_sync(0xffffffffu, gr, o);
            const int oc = __shfl_xor_sync(0xffffffffu, gc, o);
            if (ov > gb || (ov == gb && orow < gr)) { gb = ov; gr = orow; gc = oc; }
          }
          const T* src = CL ? cluster.map_shared_rank(&s_crow[par][0], gc) : &s_crow[par][0];
          if (gb >= 0.0)
            for (int k = lane; k < W; k += 32) s_u[k] = src[k];
          if (lane == 0) { s_win.val = gb; s_win.row = gr; s_win.pad = gc; }
        }
        __syncthreads();
        // 5. the pivot row is final: the CTA that owns it writes it out (physical position; moved at the end)
        const double gb = s_win.val;
        const int p = s_win.row;
        if (s_win.pad == crank && gb >= 0.0) {
          for (int k = tid; k < W; k += NT)
            if (jc + k < wc) F[(long long)(j0 + jc + k) * ld + (j0 + p)] = s_u[k];
        }
        if (crank == 0 && tid == 0) {
          s_pivrow[j] = p;
          if (!(gb > 0.0) && atomicCAS(&info[0], 0, 1) == 0) { info[1] = fi; info[2] = j0 + j; }
        }
#pragma unroll
        for (int s = 0; s < R; ++s)
          if (rows[s] == p) done |= 1u << s;
        // 6. elimination (an exactly singular column is recorded and skipped, as LAPACK getf2 does)
        if (gb > 0.0) {
          const T inv = hs_recip(s_u[t]);
#pragma unroll
          for (int s = 0; s < R; ++s) {
            if (!((done >> s) & 1u) && rows[s] < m) {
              const T l = hs_mul(a[s][t], inv);
              a[s][t] = l;
#pragma unroll
              for (int k = t + 1; k < W; ++k) a[s][k] = hs_fnma(a[s][k], l, s_u[k]);
            }
          }
        }
      }
    }
    // end of chunk: the first CH registers of every live row are final multipliers → store, then rotate the row
#pragma unroll
    for (int s = 0; s < R; ++s) {
      if (!((done >> s) & 1u) && rows[s] < m) {
#pragma unroll
        for (int t = 0; t < CH; ++t)
          if (jc + t < wc) F[(long long)(j0 + jc + t) * ld + (j0 + rows[s])] = a[s][t];
      }
#pragma unroll
      for (int k = 0; k < W - CH; ++k) a[s][k] = a[s][k + CH];
#pragma unroll
      for (int k = W - CH; k < W; ++k) a[s][k] = hs_zero<T>();
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
  T* stage = reinterpret_cast<T*>(smem_dyn);  // nm × wc
  for (int e = tid; e < nm * wc; e += NT) {
    const int i = e % nm, c = e / nm;
    stage[e] = F[(long long)(j0 + c) * ld + (j0 + s_msrc[i])];
  }
  __syncthreads();
  for (int e = tid; e < nm * wc; e += NT) {
    const int i = e % nm, c = e / nm;
    F[(long long)(j0 + c) * ld + (j0 + s_mdst[i])] = stage[e];
  }
}

// ------------------------------------------------------------------------------------------------
// panel launch: pick the register tile (W, R) and the cluster size for the tallest active panel
// ------------------------------------------------------------------------------------------------
template <typename T, int W, int R>
static void launch_panel(hs_fac* f, int f0, int nact, int j0, int C) {
  cudaStream_t st = f->ctx->stream;
  T* pool = (T*)f->pool;
  if (C == 1) {
    k_panel<T, W, R, false><<<nact, 256, 2 * W * W * sizeof(T), st>>>(f->d_fronts, pool, f->d_ipiv, f0, j0, f->d_info);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nact * C));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 2 * W * W * sizeof(T);
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
    int* info = f->d_info;
    CUDA_OK(cudaLaunchKernelEx(&cfg, k_panel<T, W, R, true>, fr, pool, ipiv, f0, j0, info));
  }
  CUDA_OK(cudaGetLastError());
}

template <typename T> struct PanelW;  // widest register tile per scalar type
template <> struct PanelW<double> { static constexpr int W0 = 64; };
template <> struct PanelW<cplx> { static constexpr int W0 = 32; };

static int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// panel width used for a level whose tallest front has n rows
template <typename T> static int choose_width(const hs_fac* f, int max_n) {
  const int W0 = PanelW<T>::W0;
  const int maxC = f->ctx->max_cluster;
  for (int W = W0, R = 1; R <= 8; W >>= 1, R <<= 1)
    if ((long long)256 * R * maxC >= max_n) return W;
  return -1;
}

template <typename T>
static void panel_dispatch(hs_fac* f, int W, int f0, int nact, int j0, int m) {
  constexpr int W0 = PanelW<T>::W0;
  const int R = W0 / W;
  const int C = pow2_ceil((m + 256 * R - 1) / (256 * R));
  if (W == W0) launch_panel<T, W0, 1>(f, f0, nact, j0, C);
  else if (W == W0 / 2) launch_panel<T, W0 / 2, 2>(f, f0, nact, j0, C);
  else if (W == W0 / 4) launch_panel<T, W0 / 4, 4>(f, f0, nact, j0, C);
  else launch_panel<T, W0 / 8, 8>(f, f0, nact, j0, C);
}


template <typename T, int W, int R> static void set_panel_attrs() {
  CUDA_OK(cudaFuncSetAttribute(k_panel<T, W, R, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CUDA_OK(cudaFuncSetAttribute(k_panel<T, W, R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * W * W * (int)sizeof(T)));
  CUDA_OK(cudaFuncSetAttribute(k_panel<T, W, R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * W * W * (int)sizeof(T)));
}
template <typename T> static void panel_setup() {
  constexpr int W0 = PanelW<T>::W0;
  set_panel_attrs<T, W0, 1>(); set_panel_attrs<T, W0 / 2, 2>(); set_panel_attrs<T, W0 / 4, 4>(); set_panel_attrs<T, W0 / 8, 8>();
}
