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
// (esz·Σ(ni² + 2·ni·nb) bytes per RHS).  Fronts with ni ≤ DB run in one fused CTA per front; larger fronts run their
// triangular part in super-block steps (k_sv_tri_*) and their rectangular part as one streamed mat-vec (k_gemv_rect).
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
// large fronts (ni > DB): triangular part of the sweeps in SUPER-BLOCKS of SB = 256 pivot rows, one launch per
// super-block step over all large fronts of the level.  In a step
//   * CTA 0 of a front (the "diagonal CTA", 1024 threads) finishes the rows of the current super-block: it subtracts the
//     contribution of the previous super-block's 256 solution entries (a 256×256 mat-vec split over 4 thread groups) and
//     solves the 256×256 triangular system with the inverted 64×64 diagonal blocks (4 dependent sub-steps);
//   * the other CTAs stream the same 256 columns over the remaining rows of the pivot block, 256 rows each.
// The chain of dependent steps per front is ni/256 kernel boundaries instead of ni/64 cluster barriers with a global
// exchange (round 1: 188 barriers of ~6 µs per sweep at the 2048² workload), and every panel is streamed by as many CTAs
// as it has 256-row tiles.  The rectangular parts (2/3 of the bytes) stay with k_gemv_rect.
// ------------------------------------------------------------------------------------------------
constexpr int SB = 256;        // super-block
constexpr int TRI_T = 1024;    // threads per CTA of the triangular step kernels

template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int o);
template <> __device__ __forceinline__ double shfl_xor_t<double>(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ cplx shfl_xor_t<cplx>(cplx v, int o) {
  return cplx{__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o)};
}

// acc[t] = Σ_{c<ncol} M[t, c]·v[c] for the 256 rows t of a tile (row t valid when t < nrow): thread (g, t) takes the columns
// c ≡ its quarter, the four partial sums meet in `red`.  Returns the sum to threads with g == 0.
template <typename T>
__device__ __forceinline__ T tile_matvec(const T* __restrict__ M, long long ld, int nrow, int ncol, const T* __restrict__ v, T (*red)[SB]) {
  const int g = threadIdx.x >> 8, t = threadIdx.x & 255;
  T acc = hs_zero<T>();
  if (t < nrow) {
    const int c0 = g * (SB / 4), c1 = min(ncol, c0 + SB / 4);
    const T* p = M + t;
#pragma unroll 16
    for (int c = c0; c < c1; ++c) acc = hs_fma(acc, p[(long long)c * ld], v[c]);
  }
  red[g][t] = acc;
  __syncthreads();
  T s = hs_zero<T>();
  if (g == 0) s = hs_add(hs_add(red[0][t], red[1][t]), hs_add(red[2][t], red[3][t]));
  __syncthreads();
  return s;
}

// forward: step s finishes super-block s of every large front (rows [s·SB, …) of v = L11⁻¹·P·x_int)
template <typename T>
__global__ void __launch_bounds__(TRI_T) k_sv_tri_fwd(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                       const int* __restrict__ gidx, const int* __restrict__ rperm,
                                                       T* __restrict__ x, long long ldx, T* __restrict__ work,
                                                       long long wstride, long long ioff0, int f0, int s) {
  const Front fr = fronts[f0 + blockIdx.x];
  const int ni = fr.ni;
  const int R0 = s * SB;
  if (R0 >= ni) return;
  const int R1 = min(R0 + SB, ni), cnt = R1 - R0;
  const T* F = pool + fr.off;
  const long long ld = fr.ld;
  T* xr = x + (long long)blockIdx.z * ldx;
  T* w = work + (long long)blockIdx.z * wstride + (fr.ioff - ioff0);
  const int* gi = gidx + fr.ioff;
  const int* rp = rperm + fr.ioff;
  const int tid = threadIdx.x, g = tid >> 8, t = tid & 255;
  __shared__ T sv[SB], su[SB];
  __shared__ T red[4][SB];
  const int P0 = R0 - SB;   // previous super-block (always full)
  if (s > 0 && tid < SB) sv[tid] = w[P0 + tid];
  __syncthreads();
  if (blockIdx.y > 0) {
    // worker tile: rows behind the current super-block
    const int r0 = R1 + (blockIdx.y - 1) * SB;
    if (r0 >= ni) return;
    const int nrow = min(SB, ni - r0);
    if (s == 0) {   // gather v = P·x_int
      if (g == 0 && t < nrow) w[r0 + t] = xr[gi[rp[r0 + t]]];
      return;
    }
    const T sum = tile_matvec<T>(F + (long long)P0 * ld + r0, ld, nrow, SB, sv, red);
    if (g == 0 && t < nrow) w[r0 + t] = hs_sub(w[r0 + t], sum);
    return;
  }
  // diagonal CTA
  {
    T sum = hs_zero<T>();
    if (s > 0) sum = tile_matvec<T>(F + (long long)P0 * ld + R0, ld, cnt, SB, sv, red);
    if (g == 0 && t < cnt) su[t] = hs_sub(s == 0 ? xr[gi[rp[R0 + t]]] : w[R0 + t], sum);
  }
  __syncthreads();
  const T* D = F + (long long)R0 * ld + R0;   // the super-block's diagonal 256×256 block
  for (int B0 = 0; B0 < cnt; B0 += DB) {
    const int bw = min(DB, cnt - B0);
    {  // (i) y_B = u_B + strict_lower(L_BB⁻¹)·u_B : row i = tid & 63, 16 column groups of 4
      const int i = tid & 63, pp = tid >> 6;
      T acc = hs_zero<T>();
      if (i < bw) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = pp * 4 + kk;
          if (k < i) acc = hs_fma(acc, D[(long long)(B0 + k) * ld + (B0 + i)], su[B0 + k]);
        }
      }
      red[pp >> 2][(pp & 3) * 64 + i] = acc;
      __syncthreads();
      if (tid < bw) {
        T y = su[B0 + tid];
#pragma unroll
        for (int q = 0; q < 16; ++q) y = hs_add(y, red[q >> 2][(q & 3) * 64 + tid]);
        su[B0 + tid] = y;
      }
      __syncthreads();
    }
    if (B0 + bw < cnt) {  // (ii) rows behind block B inside the super-block: u −= L[:, B]·y_B, 4 column groups of 16
      const int rr = B0 + bw + t;
      T acc = hs_zero<T>();
      if (rr < cnt) {
        const int c0 = g * 16, c1 = min(bw, c0 + 16);
        const T* p = D + (long long)B0 * ld + rr;
#pragma unroll 16
        for (int c = c0; c < c1; ++c) acc = hs_fma(acc, p[(long long)c * ld], su[B0 + c]);
      }
      red[g][t] = acc;
      __syncthreads();
      if (g == 0 && rr < cnt) su[rr] = hs_sub(su[rr], hs_add(hs_add(red[0][t], red[1][t]), hs_add(red[2][t], red[3][t])));
      __syncthreads();
    }
  }
  if (tid < cnt) { const T v = su[tid]; w[R0 + tid] = v; xr[gi[R0 + tid]] = v; }
}

// backward: step s finishes super-block b = nsb − 1 − s of every large front (x_int = U11⁻¹·t, t = work[0:ni] from k_gemv_rect)
template <typename T>
__global__ void __launch_bounds__(TRI_T) k_sv_tri_bwd(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                       const int* __restrict__ gidx, T* __restrict__ x, long long ldx,
                                                       T* __restrict__ work, long long wstride, long long ioff0, int f0, int s) {
  const Front fr = fronts[f0 + blockIdx.x];
  const int ni = fr.ni;
  const int nsb = (ni + SB - 1) / SB;
  const int b = nsb - 1 - s;
  if (b < 0) return;
  const int R0 = b * SB, R1 = min(R0 + SB, ni), cnt = R1 - R0;
  const T* F = pool + fr.off;
  const long long ld = fr.ld;
  T* xr = x + (long long)blockIdx.z * ldx;
  T* w = work + (long long)blockIdx.z * wstride + (fr.ioff - ioff0);
  const int* gi = gidx + fr.ioff;
  const int tid = threadIdx.x, g = tid >> 8, t = tid & 255;
  __shared__ T sv[SB], su[SB];
  __shared__ T red[4][SB];
  const int pc = s > 0 ? min(SB, ni - R1) : 0;   // columns of the super-block solved in the previous step: [R1, R1 + pc)
  if (tid < SB) sv[tid] = tid < pc ? w[R1 + tid] : hs_zero<T>();
  __syncthreads();
  if (blockIdx.y > 0) {
    // worker tile: rows above the current super-block
    if (s == 0) return;
    const int r0 = (blockIdx.y - 1) * SB;
    if (r0 >= R0) return;
    const int nrow = min(SB, R0 - r0);
    const T sum = tile_matvec<T>(F + (long long)R1 * ld + r0, ld, nrow, pc, sv, red);
    if (g == 0 && t < nrow) w[r0 + t] = hs_sub(w[r0 + t], sum);
    return;
  }
  {
    T sum = hs_zero<T>();
    if (s > 0) sum = tile_matvec<T>(F + (long long)R1 * ld + R0, ld, cnt, pc, sv, red);
    if (g == 0 && t < cnt) su[t] = hs_sub(w[R0 + t], sum);
  }
  __syncthreads();
  const T* D = F + (long long)R0 * ld + R0;
  const int nblk = (cnt + DB - 1) / DB;
  for (int q = nblk - 1; q >= 0; --q) {
    const int B0 = q * DB, bw = min(DB, cnt - B0);
    {  // (i) y_B = upper(U_BB⁻¹)·u_B (diagonal included)
      const int i = tid & 63, pp = tid >> 6;
      T acc = hs_zero<T>();
      if (i < bw) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = pp * 4 + kk;
          if (k >= i && k < bw) acc = hs_fma(acc, D[(long long)(B0 + k) * ld + (B0 + i)], su[B0 + k]);
        }
      }
      red[pp >> 2][(pp & 3) * 64 + i] = acc;
      __syncthreads();
      if (tid < bw) {
        T y = hs_zero<T>();
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) y = hs_add(y, red[qq >> 2][(qq & 3) * 64 + tid]);
        su[B0 + tid] = y;
      }
      __syncthreads();
    }
    if (B0 > 0) {  // (ii) rows above block B inside the super-block: u −= U[:, B]·y_B
      T acc = hs_zero<T>();
      if (t < B0) {
        const int c0 = g * 16, c1 = min(bw, c0 + 16);
        const T* p = D + (long long)B0 * ld + t;
#pragma unroll 16
        for (int c = c0; c < c1; ++c) acc = hs_fma(acc, p[(long long)c * ld], su[B0 + c]);
      }
      red[g][t] = acc;
      __syncthreads();
      if (g == 0 && t < B0) su[t] = hs_sub(su[t], hs_add(hs_add(red[0][t], red[1][t]), hs_add(red[2][t], red[3][t])));
      __syncthreads();
    }
  }
  if (tid < cnt) { const T v = su[tid]; w[R0 + tid] = v; xr[gi[R0 + tid]] = v; }
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

template <typename T> void prep_impl(hs_fac* f, const Level& L, cudaStream_t st) {
  if (L.max_ni == 0) return;
  dim3 grid(L.f1 - L.f0, (L.max_ni + DB - 1) / DB);
  const int dbm = std::min(DB, L.max_ni), lds = dbm | 1;
  k_trtri_diag<T><<<grid, 128, (size_t)2 * lds * lds * sizeof(T), st>>>(f->d_fronts, (T*)f->pool, L.f0, lds);
  CUDA_OK(cudaGetLastError());
  f->stats.launches_factor += 1;
}

// triangular part of the large fronts of a level: one launch per super-block step
template <typename T, bool FWD> void launch_big(hs_fac* f, const Level& L, int nbig, int64_t nrhs, T* x) {
  cudaStream_t st = f->ctx->stream;
  const int nsb = (L.max_ni + SB - 1) / SB;
  for (int s = 0; s < nsb; ++s) {
    // rows still to be updated behind (FWD) / above (BWD) the super-block of this step, for the largest front
    const int rest = std::max(0, L.max_ni - (s + 1) * SB);
    dim3 g(nbig, 1 + (rest + SB - 1) / SB, (unsigned)nrhs);
    if (FWD) k_sv_tri_fwd<T><<<g, TRI_T, 0, st>>>(f->d_fronts, (const T*)f->pool, f->d_gidx, f->d_rperm, x, f->xld, (T*)f->d_work, f->max_level_idx, L.ioff0, L.f0, s);
    else k_sv_tri_bwd<T><<<g, TRI_T, 0, st>>>(f->d_fronts, (const T*)f->pool, f->d_gidx, x, f->xld, (T*)f->d_work, f->max_level_idx, L.ioff0, L.f0, s);
    ++f->stats.launches_solve;
  }
  CUDA_OK(cudaGetLastError());
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
