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
// in-place inversion of the DB×DB diagonal blocks of L11 (unit lower) and U11 (upper); one CTA of 4 warps per
// (front, block): warps 0,1 invert the lower triangle, warps 2,3 the upper one, everything inside ONE shared copy of the
// block (33 KB in f64: six CTAs per SM).  Each 64×64 triangle is inverted as a 2×2 block matrix of 32×32 blocks,
//   inv [A 0; B C] = [A⁻¹ 0; −C⁻¹·B·A⁻¹  C⁻¹],        inv [A B; 0 C] = [A⁻¹  −A⁻¹·B·C⁻¹; 0 C⁻¹]:
//   phase 1   warp h: the 32×32 diagonal triangle h, lane j = column j of its inverse by RIGHT-looking substitution with
//             the column in registers — 31 − k independent FMAs per step instead of one dependent chain per entry
//   phase 2a  T = B·A⁻¹ (lower) / B·C⁻¹ (upper), 2b  off-diagonal block = −C⁻¹·T / −A⁻¹·T; lane j = column j, warp h =
//             rows ≡ h (mod 2), so the statically known zeros of the broadcast factor are skipped evenly in both warps
// Results replace their operands in shared memory (registers → barrier → store), the block goes back in one pass.
// Round 2 before this: one thread per column with a left-looking dependent FMA chain, 6.0 ms at the 2048² workload.
// ------------------------------------------------------------------------------------------------
// HB = half block: a CTA inverts one (2·HB)×(2·HB) block.  HB = 32 is the DB-blocked layout of the large fronts; levels
// whose pivot blocks all fit one smaller block (the bottom of the tree: ni ≤ 14, 30, 49 at the 2048² workload) use
// HB = 8, 16, 25 — the work goes with HB³.
template <typename T, int HB>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? 4 : 3) k_trtri_diag(const Front* __restrict__ fronts, T* __restrict__ pool, int f0) {
  constexpr int DBT = 2 * HB, LDS = DBT + 1, HH = (HB + 1) / 2;
  const Front fr = fronts[f0 + blockIdx.x];
  const int b0 = blockIdx.y * DBT;
  if (b0 >= fr.ni) return;
  const int db = min(DBT, fr.ni - b0);
  extern __shared__ __align__(16) unsigned char smem_x[];
  T* S = reinterpret_cast<T*>(smem_x);  // the block, column-major with pitch LDS; padded with the identity beyond db
  T* G = pool + fr.off + (long long)b0 * fr.ld + b0;
  const long long ld = fr.ld;
  const int tid = threadIdx.x, j = tid & 31, h = (tid >> 5) & 1;
  const bool lower = tid < 64, act = j < HB;
  {
    // the whole block is requested before the first shared-memory store (a rolled loop pays one latency per element)
    constexpr int NE = (DBT * DBT + 127) / 128, NBATCH = (NE * (int)sizeof(T) > 256) ? 2 : 1, NQ = (NE + NBATCH - 1) / NBATCH;
#pragma unroll
    for (int bt = 0; bt < NBATCH; ++bt) {
      T tmp[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int e = tid + (bt * NQ + q) * 128, r = e % DBT, c = e / DBT;
        tmp[q] = (r < db && c < db) ? G[(long long)c * ld + r] : (r == c ? hs_one<T>() : hs_zero<T>());
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int e = tid + (bt * NQ + q) * 128;
        if (e < DBT * DBT) S[(e / DBT) * LDS + (e % DBT)] = tmp[q];
      }
    }
  }
  __syncthreads();
  const int o = HB * h;
  if (act) {  // phase 1
    T x[HB];
#pragma unroll
    for (int i = 0; i < HB; ++i) x[i] = i == j ? hs_one<T>() : hs_zero<T>();
    if (lower) {
#pragma unroll
      for (int k = 0; k < HB - 1; ++k) {
        const T* col = S + (o + k) * LDS + o;
#pragma unroll
        for (int i = k + 1; i < HB; ++i) x[i] = hs_fnma(x[i], col[i], x[k]);
      }
      __syncwarp(__activemask());
#pragma unroll
      for (int i = 0; i < HB; ++i)
        if (i > j) S[(o + j) * LDS + o + i] = x[i];
    } else {
#pragma unroll
      for (int k = HB - 1; k >= 0; --k) {
        const T* col = S + (o + k) * LDS + o;
        x[k] = hs_mul(x[k], hs_recip_pivot(col[k]));
#pragma unroll
        for (int i = 0; i < k; ++i) x[i] = hs_fnma(x[i], col[i], x[k]);
      }
      __syncwarp(__activemask());
#pragma unroll
      for (int i = 0; i < HB; ++i)
        if (i <= j) S[(o + j) * LDS + o + i] = x[i];
    }
  }
  __syncthreads();
  if (db > HB) {   // uniform: a block that fits the first half has no off-diagonal part
    // phase 2: this thread's rows i = 2·ii + h of column j of the off-diagonal block
    T acc[HH];
#pragma unroll
    for (int ii = 0; ii < HH; ++ii) acc[ii] = hs_zero<T>();
    // B = the off-diagonal block of the factor: rows HB.. × columns 0.. (lower), rows 0.. × columns HB.. (upper)
    T* B = lower ? S + HB : S + HB * LDS;
    if (act) {
      if (lower) {  // T[i, j] = Σ_k B[i, k]·A⁻¹[k, j],  A⁻¹ unit lower: entries k > j stored, 1 at k = j
        const T* own = S + j * LDS;
#pragma unroll
        for (int k = 0; k < HB; ++k) {
          const T a = k > j ? own[k] : (k == j ? hs_one<T>() : hs_zero<T>());
#pragma unroll
          for (int ii = 0; ii < HH; ++ii)
            if (2 * ii + 1 < HB || h == 0) acc[ii] = hs_fma(acc[ii], B[k * LDS + min(2 * ii + h, HB - 1)], a);
        }
      } else {      // T[i, j] = Σ_k B[i, k]·C⁻¹[k, j],  C⁻¹ upper: entries k ≤ j stored
        const T* own = S + (HB + j) * LDS + HB;
#pragma unroll
        for (int k = 0; k < HB; ++k) {
          const T a = k <= j ? own[k] : hs_zero<T>();
#pragma unroll
          for (int ii = 0; ii < HH; ++ii)
            if (2 * ii + 1 < HB || h == 0) acc[ii] = hs_fma(acc[ii], B[k * LDS + min(2 * ii + h, HB - 1)], a);
        }
      }
    }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int ii = 0; ii < HH; ++ii) {
        if (2 * ii + h < HB) B[j * LDS + 2 * ii + h] = acc[ii];
        acc[ii] = hs_zero<T>();
      }
    }
    __syncthreads();
    if (act) {
      const T* tcol = B + j * LDS;   // column j of T
      if (lower) {  // −C⁻¹·T,  C⁻¹[i, k] unit lower at rows/columns HB..: stored for k < i
        const T* Ci = S + HB * LDS + HB;
#pragma unroll
        for (int k = 0; k < HB; ++k) {
          const T t = tcol[k];
#pragma unroll
          for (int ii = 0; ii < HH; ++ii) {
            if (2 * ii + 1 < k) continue;   // rows 2·ii + h ≤ 2·ii + 1 < k: structurally zero
            const int i = min(2 * ii + h, HB - 1);
            const T c = k < i ? Ci[k * LDS + i] : (k == i ? hs_one<T>() : hs_zero<T>());
            acc[ii] = hs_fma(acc[ii], c, t);
          }
        }
      } else {      // −A⁻¹·T,  A⁻¹[i, k] upper at rows/columns 0..: stored for k ≥ i
#pragma unroll
        for (int k = 0; k < HB; ++k) {
          const T t = tcol[k];
#pragma unroll
          for (int ii = 0; ii < HH; ++ii) {
            if (2 * ii > k) continue;       // rows 2·ii + h ≥ 2·ii > k: structurally zero
            const int i = min(2 * ii + h, HB - 1);
            const T c = k >= i ? S[k * LDS + i] : hs_zero<T>();
            acc[ii] = hs_fma(acc[ii], c, t);
          }
        }
      }
    }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int ii = 0; ii < HH; ++ii)
        if (2 * ii + h < HB) B[j * LDS + 2 * ii + h] = hs_sub(hs_zero<T>(), acc[ii]);
    }
    __syncthreads();
  }
  for (int e = tid; e < db * db; e += 128) {
    const int r = e % db, c = e / db;
    G[(long long)c * ld + r] = S[c * LDS + r];
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
// large fronts (ni > DB): triangular part of the sweeps in SUPER-BLOCKS of SB = 128 pivot rows, one launch per
// super-block step over all large fronts of the level.  In a step
//   * CTA 0 of a front (the "diagonal CTA") finishes the rows of the current super-block: it subtracts the contribution
//     of the previous super-block's 128 solution entries (a 128×128 mat-vec, 16 entries per thread held in registers)
//     and solves the 128×128 triangular system with the two inverted 64×64 diagonal blocks;
//   * the other CTAs stream the same 128 columns over the remaining rows of the pivot block, 128 rows each.
// Everything the diagonal CTA needs is requested at kernel entry — the panel into registers, the three 64×64 tiles of the
// triangle into shared memory with cp.async — so a step costs ONE global-memory latency plus a handful of shared-memory
// phases, instead of one latency per dependent phase.  The chain of dependent steps per front is ni/128 kernel
// boundaries; round 1 ran ni/64 cluster barriers with a global-memory exchange each (~6 µs per barrier, 188 per sweep at
// the 2048² workload).  The rectangular parts (2/3 of the bytes) stay with k_gemv_rect.
// ------------------------------------------------------------------------------------------------
constexpr int SB = 128;        // super-block
constexpr int TRI_T = 1024;    // threads per CTA of the triangular step kernels
constexpr int TG = TRI_T / SB; // column groups of the panel mat-vec (8), 16 columns each

template <typename T> __device__ __forceinline__ void cp_async_elem(T* smem, const T* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  if constexpr (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_fence_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory"); }

// 64×64 tile G[r0 + i, c0 + k] (i < nr, k < nc, zero elsewhere) → shared memory, column-major with pitch 64
template <typename T>
__device__ __forceinline__ void stage_tile(T* dst, const T* __restrict__ G, long long ld, int nr, int nc) {
  for (int e = threadIdx.x; e < 64 * 64; e += TRI_T) {
    const int i = e & 63, k = e >> 6;
    if (i < nr && k < nc) cp_async_elem(dst + e, G + (long long)k * ld + i);
    else dst[e] = hs_zero<T>();
  }
}

// Σ_c M[t, c]·v[c] over a 128×128 panel tile for row t = tid & 127 (valid when t < nrow, columns c < ncol): the loads of
// this thread's 16 columns are issued by `panel_load`, the products are summed over the 8 column groups through `red`.
template <typename T>
__device__ __forceinline__ void panel_load(T (&m)[SB / TG], const T* __restrict__ M, long long ld, int nrow, int ncol) {
  const int g = threadIdx.x >> 7, t = threadIdx.x & 127;
#pragma unroll
  for (int q = 0; q < SB / TG; ++q) {
    const int c = g * (SB / TG) + q;
    m[q] = (t < nrow && c < ncol) ? M[(long long)c * ld + t] : hs_zero<T>();
  }
}
template <typename T>
__device__ __forceinline__ T panel_reduce(const T (&m)[SB / TG], const T* __restrict__ v, T (*red)[SB]) {
  const int g = threadIdx.x >> 7, t = threadIdx.x & 127;
  T acc = hs_zero<T>();
#pragma unroll
  for (int q = 0; q < SB / TG; ++q) acc = hs_fma(acc, m[q], v[g * (SB / TG) + q]);
  red[g][t] = acc;
  __syncthreads();
  T sum = hs_zero<T>();
  if (g == 0) {
#pragma unroll
    for (int q = 0; q < TG; ++q) sum = hs_add(sum, red[q][t]);
  }
  __syncthreads();
  return sum;
}

// y[i] = Σ_k Tile[i, k]·u[k] over a staged 64×64 tile restricted to k < i (lower == true) or k ≥ i (upper), i < nr, k < nc.
// Row i = tid & 63, 16 column groups of 4; the result is returned to threads tid < 64.
template <typename T>
__device__ __forceinline__ T tile_tri_matvec(const T* __restrict__ tile, const T* __restrict__ u, int nr, int nc, int mode, T (*red)[SB]) {
  const int i = threadIdx.x & 63, pp = threadIdx.x >> 6;
  T acc = hs_zero<T>();
  if (i < nr) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = pp * 4 + kk;
      const bool use = k < nc && (mode == 0 ? true : (mode == 1 ? k < i : k >= i));
      if (use) acc = hs_fma(acc, tile[k * 64 + i], u[k]);
    }
  }
  red[pp >> 1][(pp & 1) * 64 + i] = acc;
  __syncthreads();
  T y = hs_zero<T>();
  if (threadIdx.x < 64) {
#pragma unroll
    for (int q = 0; q < 16; ++q) y = hs_add(y, red[q >> 1][(q & 1) * 64 + threadIdx.x]);
  }
  __syncthreads();
  return y;
}

template <typename T> constexpr size_t tri_smem() { return (3 * 64 * 64 + TG * SB + 2 * SB) * sizeof(T); }

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
  const int tid = threadIdx.x, g = tid >> 7, t = tid & 127;
  extern __shared__ __align__(16) unsigned char smem_tri[];
  T* tiles = reinterpret_cast<T*>(smem_tri);                     // T00, T11, T10: 3 × 64×64
  T(*red)[SB] = reinterpret_cast<T(*)[SB]>(tiles + 3 * 64 * 64);  // [TG][SB]
  T* sv = tiles + 3 * 64 * 64 + TG * SB;                         // [SB] solution entries of the previous super-block
  T* su = sv + SB;                                               // [SB] this super-block
  const int P0 = R0 - SB;   // previous super-block (always full)
  T m[SB / TG];
  if (blockIdx.y > 0) {
    // worker tile: rows behind the current super-block
    const int r0 = R1 + (blockIdx.y - 1) * SB;
    if (r0 >= ni) return;
    const int nrow = min(SB, ni - r0);
    if (s == 0) {   // gather v = P·x_int
      if (g == 0 && t < nrow) w[r0 + t] = xr[gi[rp[r0 + t]]];
      return;
    }
    panel_load<T>(m, F + (long long)P0 * ld + r0, ld, nrow, SB);
    if (tid < SB) sv[tid] = w[P0 + tid];
    __syncthreads();
    const T sum = panel_reduce<T>(m, sv, red);
    if (g == 0 && t < nrow) w[r0 + t] = hs_sub(w[r0 + t], sum);
    return;
  }
  // diagonal CTA: request everything at once
  const int bw0 = min(DB, cnt), bw1 = cnt - bw0;
  const T* D = F + (long long)R0 * ld + R0;
  stage_tile<T>(tiles, D, ld, bw0, bw0);                                              // L00⁻¹ (strictly lower part used)
  stage_tile<T>(tiles + 4096, D + (long long)DB * ld + DB, ld, bw1, bw1);             // L11⁻¹
  stage_tile<T>(tiles + 8192, D + DB, ld, bw1, bw0);                                  // L10
  if (s > 0) panel_load<T>(m, F + (long long)P0 * ld + R0, ld, cnt, SB);
  T base = hs_zero<T>();
  if (tid < cnt) base = s == 0 ? xr[gi[rp[R0 + tid]]] : w[R0 + tid];
  if (s > 0 && tid < SB) sv[tid] = w[P0 + tid];
  cp_async_fence_all();
  __syncthreads();
  T sum = hs_zero<T>();
  if (s > 0) sum = panel_reduce<T>(m, sv, red);
  if (tid < SB) su[tid] = tid < cnt ? hs_sub(base, sum) : hs_zero<T>();   // g == 0 ⇔ tid < 128
  __syncthreads();
  {  // block 0: y0 = u0 + strict_lower(L00⁻¹)·u0
    const T y = tile_tri_matvec<T>(tiles, su, bw0, bw0, 1, red);
    if (tid < bw0) su[tid] = hs_add(su[tid], y);
    __syncthreads();
  }
  if (bw1 > 0) {
    {  // u1 −= L10·y0
      const T y = tile_tri_matvec<T>(tiles + 8192, su, bw1, bw0, 0, red);
      if (tid < bw1) su[DB + tid] = hs_sub(su[DB + tid], y);
      __syncthreads();
    }
    {  // block 1
      const T y = tile_tri_matvec<T>(tiles + 4096, su + DB, bw1, bw1, 1, red);
      if (tid < bw1) su[DB + tid] = hs_add(su[DB + tid], y);
      __syncthreads();
    }
  }
  // Step 0 must not touch x while the worker tiles of the same launch are still gathering P·x_int from it (a row
  // interchanged out of the first super-block is read from x[gi[0 .. SB)) by a worker): its entries are written by the
  // diagonal CTA of step 1, from the copy in `work`.
  if (tid < cnt) {
    const T v = su[tid];
    w[R0 + tid] = v;
    if (s > 0 || R1 >= ni) xr[gi[R0 + tid]] = v;
  }
  if (s == 1 && tid < SB) xr[gi[tid]] = sv[tid];
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
  const int tid = threadIdx.x, g = tid >> 7, t = tid & 127;
  extern __shared__ __align__(16) unsigned char smem_tri[];
  T* tiles = reinterpret_cast<T*>(smem_tri);
  T(*red)[SB] = reinterpret_cast<T(*)[SB]>(tiles + 3 * 64 * 64);
  T* sv = tiles + 3 * 64 * 64 + TG * SB;
  T* su = sv + SB;
  const int pc = s > 0 ? min(SB, ni - R1) : 0;   // columns of the super-block solved in the previous step: [R1, R1 + pc)
  T m[SB / TG];
  if (blockIdx.y > 0) {
    // worker tile: rows above the current super-block
    if (s == 0) return;
    const int r0 = (blockIdx.y - 1) * SB;
    if (r0 >= R0) return;
    const int nrow = min(SB, R0 - r0);
    panel_load<T>(m, F + (long long)R1 * ld + r0, ld, nrow, pc);
    if (tid < SB) sv[tid] = tid < pc ? w[R1 + tid] : hs_zero<T>();
    __syncthreads();
    const T sum = panel_reduce<T>(m, sv, red);
    if (g == 0 && t < nrow) w[r0 + t] = hs_sub(w[r0 + t], sum);
    return;
  }
  const int bw0 = min(DB, cnt), bw1 = cnt - bw0;
  const T* D = F + (long long)R0 * ld + R0;
  stage_tile<T>(tiles, D, ld, bw0, bw0);                                              // U00⁻¹ (upper part incl. diagonal used)
  stage_tile<T>(tiles + 4096, D + (long long)DB * ld + DB, ld, bw1, bw1);             // U11⁻¹
  stage_tile<T>(tiles + 8192, D + (long long)DB * ld, ld, bw0, bw1);                  // U01
  if (s > 0) panel_load<T>(m, F + (long long)R1 * ld + R0, ld, cnt, pc);
  T base = hs_zero<T>();
  if (tid < cnt) base = w[R0 + tid];
  if (tid < SB) sv[tid] = tid < pc ? w[R1 + tid] : hs_zero<T>();
  cp_async_fence_all();
  __syncthreads();
  T sum = hs_zero<T>();
  if (s > 0) sum = panel_reduce<T>(m, sv, red);
  if (tid < SB) su[tid] = tid < cnt ? hs_sub(base, sum) : hs_zero<T>();
  __syncthreads();
  if (bw1 > 0) {
    {  // block 1 first: y1 = upper(U11⁻¹)·u1
      const T y = tile_tri_matvec<T>(tiles + 4096, su + DB, bw1, bw1, 2, red);
      if (tid < bw1) su[DB + tid] = y;
      __syncthreads();
    }
    {  // u0 −= U01·y1
      const T y = tile_tri_matvec<T>(tiles + 8192, su + DB, bw0, bw1, 0, red);
      if (tid < bw0) su[tid] = hs_sub(su[tid], y);
      __syncthreads();
    }
  }
  {  // block 0
    const T y = tile_tri_matvec<T>(tiles, su, bw0, bw0, 2, red);
    if (tid < bw0) su[tid] = y;
    __syncthreads();
  }
  if (tid < cnt) { const T v = su[tid]; w[R0 + tid] = v; xr[gi[R0 + tid]] = v; }
}

// Rectangular part of a large front as ONE streamed mat-vec over the whole GPU (it holds 2/3 of the front's bytes and
// has no dependency chain, so it does not belong inside the super-block step kernels):
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

template <typename T, int HB> static void launch_trtri(hs_fac* f, const Level& L, cudaStream_t st) {
  constexpr int DBT = 2 * HB;
  static bool attr = false;
  if (!attr) {
    CUDA_OK(cudaFuncSetAttribute(k_trtri_diag<T, HB>, cudaFuncAttributeMaxDynamicSharedMemorySize, DBT * (DBT + 1) * (int)sizeof(T)));
    attr = true;
  }
  dim3 grid(L.f1 - L.f0, (L.max_ni + DBT - 1) / DBT);
  k_trtri_diag<T, HB><<<grid, 128, (size_t)DBT * (DBT + 1) * sizeof(T), st>>>(f->d_fronts, (T*)f->pool, L.f0);
}

template <typename T> void prep_impl(hs_fac* f, const Level& L, cudaStream_t st) {
  if (L.max_ni == 0) return;
  // one smaller block per front where every pivot block of the level fits it; DB-blocked otherwise
  if (L.max_ni <= 16) launch_trtri<T, 8>(f, L, st);
  else if (L.max_ni <= 32) launch_trtri<T, 16>(f, L, st);
  else if (L.max_ni <= 50) launch_trtri<T, 25>(f, L, st);
  else launch_trtri<T, DB / 2>(f, L, st);
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
    if (FWD) k_sv_tri_fwd<T><<<g, TRI_T, tri_smem<T>(), st>>>(f->d_fronts, (const T*)f->pool, f->d_gidx, f->d_rperm, x, f->xld, (T*)f->d_work, f->max_level_idx, L.ioff0, L.f0, s);
    else k_sv_tri_bwd<T><<<g, TRI_T, tri_smem<T>(), st>>>(f->d_fronts, (const T*)f->pool, f->d_gidx, x, f->xld, (T*)f->d_work, f->max_level_idx, L.ioff0, L.f0, s);
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
      // shared-memory staging of x[bnd]: sized by the small fronts that do eliminate something (external leaves of the
      // subtree-per-GPU top trees have ni = 0 and boundaries of 10^4 rows: they return at once and must not size it)
      size_t sm = (size_t)std::max(L.max_nb, 1) * sizeof(T);
      if (sm > 48 * 1024) {
        int mnb = 1;
        for (int i = L.f0 + nbig; i < L.f1; ++i)
          if (f->fronts[i].ni > 0) mnb = std::max(mnb, f->fronts[i].n - f->fronts[i].ni);
        sm = (size_t)mnb * sizeof(T);
        if (sm > 200 * 1024) throw hs_error(HS_ESIZE, "tree solve: a front with at most 64 pivot rows has a boundary too long for the small-front kernel");
      }
      k_sv_small_bwd<T><<<g, NTH, sm, st>>>(f->d_fronts, pool, f->d_gidx, x, f->xld, L.f0 + nbig);
      ++s.launches_solve;
    }
  }
  CUDA_OK(cudaGetLastError());
}

}  // namespace

void hs_solve_setup() {
  CUDA_OK(cudaFuncSetAttribute(k_sv_tri_fwd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_smem<double>()));
  CUDA_OK(cudaFuncSetAttribute(k_sv_tri_bwd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_smem<double>()));
  CUDA_OK(cudaFuncSetAttribute(k_sv_tri_fwd<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_smem<cplx>()));
  CUDA_OK(cudaFuncSetAttribute(k_sv_tri_bwd<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_smem<cplx>()));
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
