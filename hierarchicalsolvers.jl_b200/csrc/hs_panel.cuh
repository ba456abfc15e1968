// LU panel kernel and its launcher; instantiated per scalar type in hs_panel_f64.cu / hs_panel_c64.cu.
#pragma once

#include <cooperative_groups.h>

#include "hs_fac.cuh"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------
// LU panel with partial pivoting restricted to the pivot block rows (rows < ni)
// ------------------------------------------------------------------------------------------------
// Panel = front rows [j0, n) × columns [j0, j0+wc).  Every thread keeps R rows × W columns in registers; a
// cluster of C CTAs (256 threads each) covers 256·R·C rows.  Per column: local |a| arg-max → candidate row
// parked in shared memory → ONE cluster barrier → warp 0 of every CTA pulls the C candidates and the winning
// row through distributed shared memory → rank-1 update from registers.
struct PanelCand {
  double val;
  int row;
  int pad;
};

template <typename T, int W, int R, bool CL>
__global__ void __launch_bounds__(256, 1) k_panel(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                   int* __restrict__ ipiv, int f0, int j0, int* __restrict__ info) {
  constexpr int NT = 256;
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

  __shared__ double s_wval[NT / 32];
  __shared__ int s_wrow[NT / 32];
  __shared__ PanelCand s_cand[2];
  __shared__ T s_crow[2][W];
  __shared__ T s_rowj[2][W];
  __shared__ T s_u[W];
  __shared__ PanelCand s_win;

  T a[R][W];
  int rows[R];
#pragma unroll
  for (int s = 0; s < R; ++s) {
    rows[s] = crank * (NT * R) + s * NT + tid;
    const bool ok = rows[s] < m;
#pragma unroll
    for (int k = 0; k < W; ++k)
      a[s][k] = (ok && k < wc) ? F[(long long)(j0 + k) * fr.ld + (j0 + rows[s])] : hs_zero<T>();
  }

#pragma unroll
  for (int j = 0; j < W; ++j) {
    if (j < wc) {
      const int par = j & 1;
      // 1. thread-local then warp-level arg-max of |a(:,j)| over candidate rows (ties → smallest row)
      double best = -1.0;
      int brow = 0x7fffffff;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        if (rows[s] >= j && j0 + rows[s] < fr.ni) {
          const double v = hs_abs1(a[s][j]);
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
      // 2. CTA-level candidate; its owner parks the row
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
      if (crank == 0 && tid == j) {  // the thread that owns panel row j (slot 0 of CTA 0)
#pragma unroll
        for (int k = 0; k < W; ++k) s_rowj[par][k] = a[0][k];
      }
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
        if (lane == 0) { s_win.val = gb; s_win.row = gr; }
      }
      __syncthreads();
      // 5. interchange + elimination
      const double gb = s_win.val;
      const int p = s_win.row;
      if (gb > 0.0) {
        if (p != j) {
#pragma unroll
          for (int s = 0; s < R; ++s) {
            if (rows[s] == p) {  // I held the pivot row: take over the old row j
              const T* rj = CL ? cluster.map_shared_rank(&s_rowj[par][0], 0) : &s_rowj[par][0];
#pragma unroll
              for (int k = 0; k < W; ++k) a[s][k] = rj[k];
            }
          }
          if (crank == 0 && tid == j) {
#pragma unroll
            for (int k = 0; k < W; ++k) a[0][k] = s_u[k];
          }
        }
        const T inv = hs_recip(s_u[j]);
#pragma unroll
        for (int s = 0; s < R; ++s) {
          if (rows[s] > j && rows[s] < m) {
            const T l = hs_mul(a[s][j], inv);
            a[s][j] = l;
#pragma unroll
            for (int k = j + 1; k < W; ++k) a[s][k] = hs_fnma(a[s][k], l, s_u[k]);
          }
        }
        if (crank == 0 && tid == 0) ipiv[fr.ioff + j0 + j] = j0 + p;
      } else {
        // exactly singular column: LAPACK getf2 records info and moves on without interchange
        if (crank == 0 && tid == 0) {
          ipiv[fr.ioff + j0 + j] = j0 + j;
          if (atomicCAS(&info[0], 0, 1) == 0) { info[1] = fi; info[2] = j0 + j; }
        }
      }
    }
  }
  // the last column's DSMEM reads must finish before any CTA of the cluster may exit
  if (CL) cluster.sync();
#pragma unroll
  for (int s = 0; s < R; ++s) {
    if (rows[s] < m) {
#pragma unroll
      for (int k = 0; k < W; ++k)
        if (k < wc) F[(long long)(j0 + k) * fr.ld + (j0 + rows[s])] = a[s][k];
    }
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
    k_panel<T, W, R, false><<<nact, 256, 0, st>>>(f->d_fronts, pool, f->d_ipiv, f0, j0, f->d_info);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nact * C));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
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
}
template <typename T> static void panel_setup() {
  constexpr int W0 = PanelW<T>::W0;
  set_panel_attrs<T, W0, 1>(); set_panel_attrs<T, W0 / 2, 2>(); set_panel_attrs<T, W0 / 4, 4>(); set_panel_attrs<T, W0 / 8, 8>();
}
