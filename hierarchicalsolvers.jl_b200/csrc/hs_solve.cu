// Tree solve (preconditioner application) — replaces ldiv! and its three sweeps, factornode.jl:62-99.
//
// Storage after hs_solve_prep(): inside every pivot block LU(A_ii) the diagonal DB×DB blocks of L11 (unit lower)
// and U11 (upper) are replaced IN PLACE by their inverses.  With that, every triangular solve becomes a sequence
// of dense matrix-vector products (no substitution chain inside a block):
//
//   forward  (post-order, _lsolve! + L half of _dsolve!):  v = [P·x[int]; x[bnd]]
//            v_0 = L00⁻¹ v_0;  for b = 0..B-1:  v[b1:n] -= F[b1:n, b]·v_b ;  v_{b+1} = L_{b+1,b+1}⁻¹ v_{b+1}
//   backward (pre-order,  U half of _dsolve! + _rsolve!):  t -= U12·x[bnd];  t_{B-1} = U⁻¹ t_{B-1};
//            for b = B-1..1:  t[0:b0] -= F[0:b0, b]·t_b ;  t_{b-1} = U_{b-1,b-1}⁻¹ t_{b-1}
//
// All of it streams each factor entry exactly once per right-hand side: the roofline is HBM bandwidth
// (esz·Σ(ni² + 2·ni·nb) bytes per RHS).  Fronts with ni ≤ DB run in one fused CTA per front; larger fronts run one
// launch per block step with one CTA per 32-row tile.
#include <cuda_runtime.h>

#include <cooperative_groups.h>

#include <algorithm>

#include "hs_fac.cuh"

namespace {

namespace cg = cooperative_groups;

constexpr int DB = 64;  // diagonal blocks of L11/U11 that are inverted in place (both scalar types)
template <typename T> struct SolveCfg { static constexpr int DB = ::DB; };

constexpr int NW = 8;          // warps per CTA
constexpr int NTH = NW * 32;

// ------------------------------------------------------------------------------------------------
// in-place inversion of the DB×DB diagonal blocks of L11 (unit lower) and U11 (upper); one CTA per (front, block).
// Thread j < db builds column j of L_bb⁻¹ by forward substitution, thread 64 + j column j of U_bb⁻¹ by back
// substitution; columns are independent, so there is no barrier inside the loops.  The factor is read from global
// memory (every lane reads the same entry → one broadcast transaction), the inverse grows in shared memory.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_trtri_diag(const Front* __restrict__ fronts, T* __restrict__ pool, int f0, int lds) {
  const Front fr = fronts[f0 + blockIdx.x];
  const int b0 = blockIdx.y * DB;
  if (b0 >= fr.ni) return;
  const int db = min(DB, fr.ni - b0);
  const int LDS = lds;
  extern __shared__ __align__(16) unsigned char smem_x[];
  T* S = reinterpret_cast<T*>(smem_x);  // the factor block (both triangles), column-major db × LDS
  T* X = S + (size_t)lds * lds;         // its inverse
  T* G = pool + fr.off + (long long)b0 * fr.ld + b0;
  const long long ld = fr.ld;
  const int tid = threadIdx.x;
  for (int e = tid; e < db * db; e += blockDim.x) {
    const int i = e % db, j = e / db;
    S[j * LDS + i] = G[(long long)j * ld + i];
  }
  __syncthreads();
  // all lanes walk the same (i, k) pairs: S[k,i] is a broadcast read, X[.,j] is private to the lane
  if (tid < 64) {
    const int j = tid;  // column j of L⁻¹:  x_ij = −(L_ij + Σ_{j<k<i} L_ik·x_kj)
    const bool act = j < db;
    for (int i = 1; i < db; ++i) {
      T sacc = (act && i > j) ? S[j * LDS + i] : hs_zero<T>();
      for (int k = 1; k < i; ++k) {
        const T g = S[k * LDS + i];
        if (act && k > j) sacc = hs_fma(sacc, g, X[j * LDS + k]);
      }
      if (act && i > j) X[j * LDS + i] = hs_sub(hs_zero<T>(), sacc);
    }
  } else {
    const int j = tid - 64;  // column j of U⁻¹:  x_ij = −(Σ_{i<k≤j} U_ik·x_kj)/U_ii
    const bool act = j < db;
    if (act) X[j * LDS + j] = hs_recip_pivot(S[j * LDS + j]);
    for (int i = db - 2; i >= 0; --i) {
      T sacc = hs_zero<T>();
      for (int k = i + 1; k < db; ++k) {
        const T g = S[k * LDS + i];
        if (act && k <= j && i < j) sacc = hs_fma(sacc, g, X[j * LDS + k]);
      }
      const T dinv = hs_recip_pivot(S[i * LDS + i]);
      if (act && i < j) X[j * LDS + i] = hs_sub(hs_zero<T>(), hs_mul(sacc, dinv));
    }
  }
  __syncthreads();
  for (int e = tid; e < db * db; e += blockDim.x) {
    const int i = e % db, j = e / db;
    G[(long long)j * ld + i] = X[j * LDS + i];
  }
}

// ------------------------------------------------------------------------------------------------
// building block: for one tile of 32 rows (lane = row), accumulate  Σ_k M[row, k]·v[k]  over k in [k0, k1),
// the columns split round-robin over the NW warps; `tri` masks the triangular diagonal blocks:
//   tri = 0 full,  1 strictly lower (k < row_in_block),  2 upper incl. diagonal (k ≥ row_in_block)
// Returns the CTA-wide sum for this lane's row in warp 0 (other warps return garbage); uses red[NW][32].
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T tile_dot(const T* __restrict__ M, long long ld, int row, bool row_ok, int k0, int k1,
                                      const T* __restrict__ v, int tri, int rdiag, T (*red)[32]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T acc = hs_zero<T>();
  if (row_ok) {
    const T* p = M + row;
    int k = k0 + warp;
#pragma unroll 4
    for (; k < k1; k += NW) {
      bool use = true;
      if (tri == 1) use = (k - k0) < rdiag;
      else if (tri == 2) use = (k - k0) >= rdiag;
      if (use) acc = hs_fma(acc, p[(long long)k * ld], v[k - k0]);
    }
  }
  red[warp][lane] = acc;
  __syncthreads();
  T s = hs_zero<T>();
  if (warp == 0) {
#pragma unroll
    for (int w = 0; w < NW; ++w) s = hs_add(s, red[w][lane]);
  }
  __syncthreads();
  return s;
}

// ------------------------------------------------------------------------------------------------
// small fronts (ni ≤ DB): one CTA per (front, rhs) does the whole front
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NTH) k_sv_small_fwd(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                       const int* __restrict__ gidx, const int* __restrict__ rperm,
                                                       T* __restrict__ x, long long ldx, int f0) {
  constexpr int DB = SolveCfg<T>::DB;
  const Front fr = fronts[f0 + blockIdx.x];
  const int n = fr.n, ni = fr.ni;
  if (ni == 0) return;
  const T* F = pool + fr.off;
  const long long ld = fr.ld;
  T* xr = x + (long long)blockIdx.y * ldx;
  const int* gi = gidx + fr.ioff;
  const int* rp = rperm + fr.ioff;
  __shared__ T tin[DB], t[DB];
  __shared__ T red[NW][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < ni; k += NTH) tin[k] = xr[gi[rp[k]]];
  __syncthreads();
  // t = L11⁻¹·(P x_int): unit lower
  for (int r0 = 0; r0 < ni; r0 += 32) {
    const int r = r0 + lane;
    const T s = tile_dot<T>(F, ld, r, r < ni, 0, min(ni, r0 + 32), tin, 1, r, red);
    if (warp == 0 && r < ni) {
      const T v = hs_add(tin[r], s);
      t[r] = v;
      xr[gi[r]] = v;
    }
  }
  __syncthreads();
  // x_bnd -= L21·t
  for (int r0 = ni; r0 < n; r0 += 32) {
    const int r = r0 + lane;
    const T s = tile_dot<T>(F, ld, r, r < n, 0, ni, t, 0, 0, red);
    if (warp == 0 && r < n) { const int g = gi[r]; xr[g] = hs_sub(xr[g], s); }
  }
}

template <typename T>
__global__ void __launch_bounds__(NTH) k_sv_small_bwd(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                       const int* __restrict__ gidx, T* __restrict__ x, long long ldx,
                                                       int f0) {
  constexpr int DB = SolveCfg<T>::DB;
  const Front fr = fronts[f0 + blockIdx.x];
  const int n = fr.n, ni = fr.ni, nb = n - ni;
  if (ni == 0) return;
  const T* F = pool + fr.off;
  const long long ld = fr.ld;
  T* xr = x + (long long)blockIdx.y * ldx;
  const int* gi = gidx + fr.ioff;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xb = reinterpret_cast<T*>(smem_raw);  // nb entries
  __shared__ T y[DB];
  __shared__ T red[NW][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < nb; k += NTH) xb[k] = xr[gi[ni + k]];
  __syncthreads();
  // y = t − U12·x_bnd
  for (int r0 = 0; r0 < ni; r0 += 32) {
    const int r = r0 + lane;
    const T s = tile_dot<T>(F + ni * ld, ld, r, r < ni, 0, nb, xb, 0, 0, red);
    if (warp == 0 && r < ni) y[r] = hs_sub(xr[gi[r]], s);
  }
  __syncthreads();
  // x_int = U11⁻¹·y: upper incl. diagonal
  for (int r0 = 0; r0 < ni; r0 += 32) {
    const int r = r0 + lane;
    const T s = tile_dot<T>(F + (long long)r0 * ld, ld, r, r < ni, 0, ni - r0, y + r0, 2, lane, red);
    if (warp == 0 && r < ni) xr[gi[r]] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// large fronts (ni > DB): one thread-block CLUSTER per (front, rhs) runs the whole sweep; block steps are separated
// by cluster barriers instead of kernel launches.  A warp owns 32-row tiles (lane = row) and streams its rows of the
// current block column; global warp 0 additionally owns the next diagonal block and applies its inverse, so that the
// next step starts from a final v_b.  The working vector lives in `work` (global, read/written with .cg accesses).
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T ldcg(const T* p);
template <> __device__ __forceinline__ double ldcg<double>(const double* p) { return __ldcg(p); }
template <> __device__ __forceinline__ cplx ldcg<cplx>(const cplx* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return cplx{v.x, v.y};
}
template <typename T> __device__ __forceinline__ void stcg(T* p, T v);
template <> __device__ __forceinline__ void stcg<double>(double* p, double v) { __stcg(p, v); }
template <> __device__ __forceinline__ void stcg<cplx>(cplx* p, cplx v) { __stcg(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

// Σ_k M[r, k]·v[k] for one row per lane, v (≤ 64 values) in shared memory
template <typename T>
__device__ __forceinline__ T row_dot(const T* __restrict__ Mrow, long long ld, int kcnt, const T* __restrict__ v, bool ok) {
  T acc = hs_zero<T>();
  if (ok) {
#pragma unroll 16
    for (int k = 0; k < kcnt; ++k) acc = hs_fma(acc, Mrow[(long long)k * ld], v[k]);
  }
  return acc;
}

// two consecutive rows per lane through one 16-byte load (f64 only; Mrow must be 16-byte aligned): keeps
// 16 × 16 B per thread in flight, which is what lets a handful of SMs pull a useful share of HBM bandwidth
__device__ __forceinline__ void row_dot2(const double* __restrict__ Mrow, long long ld, int kcnt,
                                         const double* __restrict__ v, double& a0, double& a1) {
  double s0 = 0.0, s1 = 0.0;
#pragma unroll 16
  for (int k = 0; k < kcnt; ++k) {
    const double2 m = *reinterpret_cast<const double2*>(Mrow + (long long)k * ld);
    const double vk = v[k];
    s0 = fma(m.x, vk, s0);
    s1 = fma(m.y, vk, s1);
  }
  a0 = s0; a1 = s1;
}

// one tile of the streamed update  w[r] −= Σ_k M[r,k]·v[k]  (or the final scatter into x when `fin`);
// a tile is 64 rows for aligned f64 fronts (2 rows per lane) and 32 rows otherwise
template <typename T> struct TileRows { static constexpr int N = 32; };
template <> struct TileRows<double> { static constexpr int N = 64; };

template <typename T>
__device__ __forceinline__ void tile_update(const T* __restrict__ Mb, long long ld, int kcnt, const T* __restrict__ v,
                                            int rbase, int rend, bool al, T* __restrict__ w, T* __restrict__ xr,
                                            const int* __restrict__ gi, bool fin, const T* __restrict__ src_x) {
  const int lane = threadIdx.x & 31;
  auto put = [&](int r, T acc) {
    const T cur = src_x ? src_x[gi[r]] : ldcg(&w[r]);
    const T val = hs_sub(cur, acc);
    if (fin) xr[gi[r]] = val; else stcg(&w[r], val);
  };
  if constexpr (sizeof(T) == 8) {
    const int r = rbase + 2 * lane;
    if (al && (rbase & 1) == 0 && r + 1 < rend) {
      double a0, a1;
      row_dot2(reinterpret_cast<const double*>(Mb) + r, ld, kcnt, reinterpret_cast<const double*>(v), a0, a1);
      put(r, a0); put(r + 1, a1);
    } else {
      // unaligned front or ragged tail: scalar fallback over the two rows this lane owns
      for (int q = 0; q < 2; ++q) {
        const int rr = r + q;
        if (rr < rend) put(rr, row_dot<T>(Mb + rr, ld, kcnt, v, true));
      }
    }
  } else {
    const int r = rbase + lane;
    if (r < rend) put(r, row_dot<T>(Mb + r, ld, kcnt, v, true));
  }
}

// Work of the "diagonal CTA" in one step, executed by all NTH threads of that CTA:
//   u[i]  = win[i] − Σ_{k<db} M[i, k]·vb[k]                 for the cnt ≤ 64 rows starting at Mrows / win
//   y[i]  = (tri == 1 ? u[i] : 0) + Σ_k Tinv[i, k]·u[k]      for i < dn   (k < i lower-unit, k ≥ i upper)
//   y[i]  = u[i]                                             for dn ≤ i < cnt (rows below a partial last block)
// Returns y[tid] to threads tid < cnt (others get zero).
template <typename T>
__device__ __forceinline__ T diag_cta(const T* __restrict__ Mrows, long long ld, int db, const T* __restrict__ vb,
                                      T win, int cnt, const T* __restrict__ Tinv, int dn, int tri, T* sT, T (*spart)[64],
                                      T* su) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int LDT = 65;
  // all global loads of this step are issued up front with static addressing (16 + 16 per thread in flight)
  {
    const int i = tid & 63, kq = tid >> 6;  // 4 columns per pass
    T pre[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int k = kq + 4 * q;
      pre[q] = (i < dn && k < dn) ? Tinv[(long long)k * ld + i] : hs_zero<T>();
    }
    T a0 = hs_zero<T>(), a1 = hs_zero<T>();
    const bool ok0 = lane < cnt, ok1 = lane + 32 < cnt;
    T m0[8], m1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = warp + NW * q;
      m0[q] = (ok0 && k < db) ? Mrows[(long long)k * ld + lane] : hs_zero<T>();
      m1[q] = (ok1 && k < db) ? Mrows[(long long)k * ld + lane + 32] : hs_zero<T>();
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = warp + NW * q;
      const T vk = k < db ? vb[k] : hs_zero<T>();
      a0 = hs_fma(a0, m0[q], vk);
      a1 = hs_fma(a1, m1[q], vk);
    }
    spart[warp][lane] = a0;
    spart[warp][lane + 32] = a1;
#pragma unroll
    for (int q = 0; q < 16; ++q) sT[(kq + 4 * q) * LDT + i] = pre[q];
  }
  __syncthreads();
  if (tid < 64) {
    T u = hs_zero<T>();
    if (tid < cnt) {
      T sum = hs_zero<T>();
#pragma unroll
      for (int w = 0; w < NW; ++w) sum = hs_add(sum, spart[w][tid]);
      u = hs_sub(win, sum);
    }
    su[tid] = u;
  }
  __syncthreads();
  {
    T a0 = hs_zero<T>(), a1 = hs_zero<T>();
    const int r0 = lane, r1 = lane + 32;
    for (int k = warp; k < dn; k += NW) {
      const T uk = su[k];
      const bool use0 = r0 < dn && (tri == 1 ? k < r0 : k >= r0);
      const bool use1 = r1 < dn && (tri == 1 ? k < r1 : k >= r1);
      if (use0) a0 = hs_fma(a0, sT[k * LDT + r0], uk);
      if (use1) a1 = hs_fma(a1, sT[k * LDT + r1], uk);
    }
    spart[warp][lane] = a0;
    spart[warp][lane + 32] = a1;
  }
  __syncthreads();
  T y = hs_zero<T>();
  if (tid < cnt) {
    if (tid < dn) {
      T sum = tri == 1 ? su[tid] : hs_zero<T>();
#pragma unroll
      for (int w = 0; w < NW; ++w) sum = hs_add(sum, spart[w][tid]);
      y = sum;
    } else {
      y = su[tid];
    }
  }
  __syncthreads();
  return y;
}

// Rectangular part of a large front as ONE streamed mat-vec over the whole GPU (it holds 2/3 of the front's bytes and
// has no dependency chain, so it does not belong inside the block-step loop of the cluster kernels):
//   FWD:  x[bnd] −= L21·t,            t = x[int] after the triangular solve         (rows ni..n, columns 0..ni)
//   BWD:  work[0:ni] = x[int] − U12·x[bnd]                                            (rows 0..ni, columns ni..n)
// One CTA per (front, 32-row tile, rhs): lane = row, the 8 warps split the columns, partial sums meet in shared memory.
constexpr int VB = 2048;  // vector entries staged in shared memory per pass
template <typename T, bool FWD>
__global__ void __launch_bounds__(NTH) k_gemv_rect(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                    const int* __restrict__ gidx, T* __restrict__ x, long long ldx,
                                                    T* __restrict__ work, long long wstride, long long ioff0, int f0) {
  const Front fr = fronts[f0 + blockIdx.x];
  const int n = fr.n, ni = fr.ni, nb = n - ni;
  const int nrows = FWD ? nb : ni, ncols = FWD ? ni : nb;
  const int r0 = blockIdx.y * 32;
  if (r0 >= nrows) return;
  const T* M = pool + fr.off + (FWD ? (long long)ni : (long long)ni * fr.ld);  // first row / first column of the block
  const long long ld = fr.ld;
  T* xr = x + (long long)blockIdx.z * ldx;
  T* w = work + (long long)blockIdx.z * wstride + (fr.ioff - ioff0);
  const int* gi = gidx + fr.ioff;
  const int* gv = gi + (FWD ? 0 : ni);   // indices of the vector entries
  const int* gr = gi + (FWD ? ni : 0);   // indices of the rows
  __shared__ T sv[VB];
  __shared__ T red[NW][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = r0 + lane;
  const bool ok = r < nrows;
  T acc = hs_zero<T>();
  for (int c0 = 0; c0 < ncols; c0 += VB) {
    const int cw = min(VB, ncols - c0);
    __syncthreads();
    for (int k = tid; k < cw; k += NTH) sv[k] = xr[gv[c0 + k]];
    __syncthreads();
    if (ok) {
      const T* p = M + (long long)c0 * ld + r;
#pragma unroll 16
      for (int k = warp; k < cw; k += NW) acc = hs_fma(acc, p[(long long)k * ld], sv[k]);
    }
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && ok) {
    T sum = hs_zero<T>();
#pragma unroll
    for (int q = 0; q < NW; ++q) sum = hs_add(sum, red[q][lane]);
    const int g = gr[r];
    if (FWD) xr[g] = hs_sub(xr[g], sum); else w[r] = hs_sub(xr[g], sum);
  }
}

template <typename T, bool CL>
__global__ void __launch_bounds__(NTH) k_sv_big_fwd(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                     const int* __restrict__ gidx, const int* __restrict__ rperm,
                                                     T* __restrict__ x, long long ldx, T* __restrict__ work,
                                                     long long wstride, long long ioff0, int f0) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = CL ? (int)cluster.num_blocks() : 1;
  const int crank = CL ? (int)cluster.block_rank() : 0;
  const Front fr = fronts[f0 + (CL ? blockIdx.x / C : blockIdx.x)];
  const int n = fr.n, ni = fr.ni;
  const T* F = pool + fr.off;
  const long long ld = fr.ld;
  const bool al = (fr.off & 1) == 0;  // 16-byte aligned columns (false only for the root-boundary pseudo front)
  T* xr = x + (long long)blockIdx.y * ldx;
  T* w = work + (long long)blockIdx.y * wstride + (fr.ioff - ioff0);
  const int* gi = gidx + fr.ioff;
  const int* rp = rperm + fr.ioff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sT = reinterpret_cast<T*>(smem_raw);        // 64·65
  T(*spart)[64] = reinterpret_cast<T(*)[64]>(sT + 64 * 65);  // NW·64
  T* su = sT + 64 * 65 + NW * 64;                // 64
  T* svb = su + 64;                              // 64
  auto sync = [&]() { if (CL) cluster.sync(); else __syncthreads(); };
  // tile workers: the warps of CTAs 1..C-1 (all warps of the only CTA when C == 1)
  const int nworkers = C > 1 ? NW * (C - 1) : NW;
  const int me = C > 1 ? (crank - 1) * NW + warp : warp;
  const bool worker = C == 1 || crank > 0;
  // gather v = P·x_int (the boundary rows are updated afterwards by k_gemv_rect)
  for (int r = crank * NTH + tid; r < ni; r += NTH * C) stcg(&w[r], xr[gi[rp[r]]]);
  sync();
  const int B = (ni + DB - 1) / DB;
  if (crank == 0) {  // v_0 = L00⁻¹·v_0
    const int db0 = min(DB, ni);
    const T win = tid < db0 ? ldcg(&w[tid]) : hs_zero<T>();
    const T y = diag_cta<T>(F, ld, 0, svb, win, db0, F, db0, 1, sT, spart, su);
    if (tid < db0) { stcg(&w[tid], y); xr[gi[tid]] = y; }
  }
  sync();
  for (int b = 0; b < B; ++b) {
    const int b0 = b * DB, b1 = min(b0 + DB, ni), db = b1 - b0;
    const bool last = b1 >= ni;
    if (tid < 64) svb[tid] = tid < db ? ldcg(&w[b0 + tid]) : hs_zero<T>();
    __syncthreads();
    const T* Fb = F + (long long)b0 * ld;
    // rows after b1 in tiles of 32; the first 64 rows (the next diagonal block) belong to CTA 0 unless this is the
    // last block column
    if (!last) {
      if (crank == 0) {
        const int nb1 = min(b1 + DB, ni), dn = nb1 - b1;
        const int cnt = dn;
        const T win = tid < cnt ? ldcg(&w[b1 + tid]) : hs_zero<T>();
        const T y = diag_cta<T>(Fb + b1, ld, db, svb, win, cnt, F + (long long)b1 * ld + b1, dn, 1, sT, spart, su);
        if (tid < cnt) { stcg(&w[b1 + tid], y); if (tid < dn) xr[gi[b1 + tid]] = y; }
      }
    }
    if (worker) {
      constexpr int TR = TileRows<T>::N;
      const int rstart = b1 + 64;  // rows of the pivot block below the next diagonal block
      const int ntiles = last ? 0 : (ni - rstart + TR - 1) / TR;
      for (int t = me; t < ntiles; t += nworkers)
        tile_update<T>(Fb, ld, db, svb, rstart + t * TR, ni, al, w, xr, gi, false, nullptr);
    }
    if (b + 1 < B) sync();
  }
}

template <typename T, bool CL>
__global__ void __launch_bounds__(NTH) k_sv_big_bwd(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                     const int* __restrict__ gidx, T* __restrict__ x, long long ldx,
                                                     T* __restrict__ work, long long wstride, long long ioff0, int f0) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = CL ? (int)cluster.num_blocks() : 1;
  const int crank = CL ? (int)cluster.block_rank() : 0;
  const Front fr = fronts[f0 + (CL ? blockIdx.x / C : blockIdx.x)];
  const int n = fr.n, ni = fr.ni, nb = n - ni;
  const T* F = pool + fr.off;
  const long long ld = fr.ld;
  const bool al = (fr.off & 1) == 0;
  T* xr = x + (long long)blockIdx.y * ldx;
  T* w = work + (long long)blockIdx.y * wstride + (fr.ioff - ioff0);
  const int* gi = gidx + fr.ioff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sT = reinterpret_cast<T*>(smem_raw);
  T(*spart)[64] = reinterpret_cast<T(*)[64]>(sT + 64 * 65);
  T* su = sT + 64 * 65 + NW * 64;
  T* svb = su + 64;
  auto sync = [&]() { if (CL) cluster.sync(); else __syncthreads(); };
  const int nworkers = C > 1 ? NW * (C - 1) : NW;
  const int me = C > 1 ? (crank - 1) * NW + warp : warp;
  const bool worker = C == 1 || crank > 0;
  const int B = (ni + DB - 1) / DB;
  const int l0 = (B - 1) * DB, dl = ni - l0;
  // phase 0: work[0:ni] already holds t = x_int − U12·x_bnd (k_gemv_rect); CTA 0 applies the inverse of the last
  // diagonal block
  if (crank == 0) {
    const T win = tid < dl ? ldcg(&w[l0 + tid]) : hs_zero<T>();
    const T y = diag_cta<T>(F, ld, 0, svb, win, dl, F + (long long)l0 * ld + l0, dl, 2, sT, spart, su);
    if (tid < dl) { stcg(&w[l0 + tid], y); xr[gi[l0 + tid]] = y; }
  }
  for (int b = B - 1; b >= 1; --b) {
    sync();
    const int b0 = b * DB, b1 = min(b0 + DB, ni), db = b1 - b0;
    const int p0 = b0 - DB;  // previous diagonal block [p0, b0), always full
    if (tid < 64) svb[tid] = tid < db ? ldcg(&w[b0 + tid]) : hs_zero<T>();
    __syncthreads();
    const T* Fb = F + (long long)b0 * ld;
    if (crank == 0) {
      const T win = tid < 64 ? ldcg(&w[p0 + tid]) : hs_zero<T>();
      const T y = diag_cta<T>(Fb + p0, ld, db, svb, win, 64, F + (long long)p0 * ld + p0, 64, 2, sT, spart, su);
      if (tid < 64) { stcg(&w[p0 + tid], y); xr[gi[p0 + tid]] = y; }
    }
    if (worker) {
      constexpr int TR = TileRows<T>::N;
      const int ntop = (p0 + TR - 1) / TR;
      for (int t = me; t < ntop; t += nworkers)
        tile_update<T>(Fb, ld, db, svb, t * TR, p0, al, w, xr, gi, false, nullptr);
    }
  }
}

template <typename T> void prep_impl(hs_fac* f, const Level& L, cudaStream_t st) {
  if (L.max_ni == 0) return;
  dim3 grid(L.f1 - L.f0, (L.max_ni + DB - 1) / DB);
  const int dbm = std::min(DB, L.max_ni), lds = dbm | 1;
  k_trtri_diag<T><<<grid, 128, (size_t)2 * lds * lds * sizeof(T), st>>>(f->d_fronts, (T*)f->pool, L.f0, lds);
  CUDA_OK(cudaGetLastError());
  f->stats.launches_factor += 1;
}

template <typename T> constexpr size_t big_smem() { return (64 * 65 + NW * 64 + 128) * sizeof(T); }

static int pow2_ceil_i(int v) { int p = 1; while (p < v) p <<= 1; return p; }

template <typename T, bool FWD> void launch_big(hs_fac* f, const Level& L, int nbig, int64_t nrhs, T* x) {
  cudaStream_t st = f->ctx->stream;
  const int C = std::min(f->ctx->max_cluster, std::max(1, pow2_ceil_i((L.max_n + 255) / 256)));
  const Front* fr = f->d_fronts;
  const T* pool = (const T*)f->pool;
  const int* gidx = f->d_gidx;
  const int* rperm = f->d_rperm;
  T* work = (T*)f->d_work;
  long long ldx = f->xld, ws = f->max_level_idx, ioff0 = L.ioff0;
  int f0 = L.f0;
  if (C == 1) {
    dim3 g(nbig, (unsigned)nrhs);
    if (FWD) k_sv_big_fwd<T, false><<<g, NTH, big_smem<T>(), st>>>(fr, pool, gidx, rperm, x, ldx, work, ws, ioff0, f0);
    else k_sv_big_bwd<T, false><<<g, NTH, big_smem<T>(), st>>>(fr, pool, gidx, x, ldx, work, ws, ioff0, f0);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nbig * C), (unsigned)nrhs);
    cfg.blockDim = dim3(NTH);
    cfg.dynamicSmemBytes = big_smem<T>();
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (FWD) CUDA_OK(cudaLaunchKernelEx(&cfg, k_sv_big_fwd<T, true>, fr, pool, gidx, rperm, x, ldx, work, ws, ioff0, f0));
    else CUDA_OK(cudaLaunchKernelEx(&cfg, k_sv_big_bwd<T, true>, fr, pool, gidx, x, ldx, work, ws, ioff0, f0));
  }
  CUDA_OK(cudaGetLastError());
  ++f->stats.launches_solve;
}

template <typename T> void run_impl(hs_fac* f, int64_t nrhs, void* xv, int which) {
  cudaStream_t st = f->ctx->stream;
  T* x = (T*)xv;
  const T* pool = (const T*)f->pool;
  hs_stats_t& s = f->stats;
  auto nbig_of = [&](const Level& L) {
    return (int)(std::partition_point(L.ni_sorted.begin(), L.ni_sorted.end(), [&](int v) { return v > DB; }) - L.ni_sorted.begin());
  };
  for (size_t li = 0; (which & 1) && li < f->flevels.size(); ++li) {  // post-order
    const Level& L = f->flevels[li];
    const int nf = L.f1 - L.f0, nbig = nbig_of(L);
    if (nbig > 0) {
      launch_big<T, true>(f, L, nbig, nrhs, x);
      if (L.max_nb > 0) {
        dim3 g(nbig, (L.max_nb + 31) / 32, (unsigned)nrhs);
        k_gemv_rect<T, true><<<g, NTH, 0, st>>>(f->d_fronts, pool, f->d_gidx, x, f->xld, (T*)f->d_work, f->max_level_idx, L.ioff0, L.f0);
        ++s.launches_solve;
      }
    }
    if (nf - nbig > 0) {
      dim3 g(nf - nbig, (unsigned)nrhs);
      k_sv_small_fwd<T><<<g, NTH, 0, st>>>(f->d_fronts, pool, f->d_gidx, f->d_rperm, x, f->xld, L.f0 + nbig);
      ++s.launches_solve;
    }
    if (L.thin >= 0) hs_comp_solve(f, f->clevels[L.thin], nrhs, x, true);   // x[bnd] += Qb·(border rows)
  }
  for (size_t li = f->flevels.size(); (which & 2) && li-- > 0;) {  // pre-order
    const Level& L = f->flevels[li];
    const int nf = L.f1 - L.f0, nbig = nbig_of(L);
    if (L.thin >= 0) hs_comp_solve(f, f->clevels[L.thin], nrhs, x, false);  // border rows ← Ri·x[bnd]
    if (nbig > 0) {
      dim3 g(nbig, (L.max_ni + 31) / 32, (unsigned)nrhs);
      k_gemv_rect<T, false><<<g, NTH, 0, st>>>(f->d_fronts, pool, f->d_gidx, x, f->xld, (T*)f->d_work, f->max_level_idx, L.ioff0, L.f0);
      ++s.launches_solve;
      launch_big<T, false>(f, L, nbig, nrhs, x);
    }
    if (nf - nbig > 0) {
      dim3 g(nf - nbig, (unsigned)nrhs);
      const size_t sm = (size_t)std::max(L.max_nb, 1) * sizeof(T);
      k_sv_small_bwd<T><<<g, NTH, sm, st>>>(f->d_fronts, pool, f->d_gidx, x, f->xld, L.f0 + nbig);
      ++s.launches_solve;
    }
  }
  CUDA_OK(cudaGetLastError());
}

}  // namespace

void hs_solve_setup() {
  CUDA_OK(cudaFuncSetAttribute(k_trtri_diag<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (DB + 1) * (DB + 1) * (int)sizeof(double)));
  CUDA_OK(cudaFuncSetAttribute(k_trtri_diag<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (DB + 1) * (DB + 1) * (int)sizeof(cplx)));
  CUDA_OK(cudaFuncSetAttribute(k_sv_big_fwd<cplx, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem<cplx>()));
  CUDA_OK(cudaFuncSetAttribute(k_sv_big_bwd<cplx, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem<cplx>()));
  CUDA_OK(cudaFuncSetAttribute(k_sv_big_fwd<cplx, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem<cplx>()));
  CUDA_OK(cudaFuncSetAttribute(k_sv_big_bwd<cplx, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem<cplx>()));
  CUDA_OK(cudaFuncSetAttribute(k_sv_big_fwd<double, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CUDA_OK(cudaFuncSetAttribute(k_sv_big_bwd<double, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CUDA_OK(cudaFuncSetAttribute(k_sv_big_fwd<cplx, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CUDA_OK(cudaFuncSetAttribute(k_sv_big_bwd<cplx, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CUDA_OK(cudaFuncSetAttribute(k_sv_small_bwd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_sv_small_bwd<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

int hs_solve_block(hs_dtype) { return DB; }

void hs_solve_prep(hs_fac* f, const Level& L, cudaStream_t st) {
  if (f->dtype == HS_F64) prep_impl<double>(f, L, st); else prep_impl<cplx>(f, L, st);
}

void hs_solve_run(hs_fac* f, int64_t nrhs, void* x, int which) {
  if (f->dtype == HS_F64) run_impl<double>(f, nrhs, x, which); else run_impl<cplx>(f, nrhs, x, which);
}
