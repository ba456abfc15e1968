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

#include <algorithm>

#include "hs_fac.cuh"

namespace {

template <typename T> struct SolveCfg;
template <> struct SolveCfg<double> { static constexpr int DB = 128; };
template <> struct SolveCfg<cplx> { static constexpr int DB = 96; };

constexpr int NW = 8;          // warps per CTA
constexpr int NTH = NW * 32;

// ------------------------------------------------------------------------------------------------
// in-place inversion of the diagonal blocks of L11 (unit lower) and U11 (upper); one CTA per (front, block)
// LAPACK trti2 order: L from the last column to the first, U from the first to the last; both in one sweep.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_trtri_diag(const Front* __restrict__ fronts, T* __restrict__ pool, int f0) {
  constexpr int DB = SolveCfg<T>::DB;
  const Front fr = fronts[f0 + blockIdx.x];
  const int b0 = blockIdx.y * DB;
  if (b0 >= fr.ni) return;
  const int db = min(DB, fr.ni - b0);
  const int lds = db | 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* S = reinterpret_cast<T*>(smem_raw);  // db × lds, column-major
  T* vl = S + (size_t)db * lds;           // column of L being replaced
  T* vu = vl + db;                        // column of U being replaced
  T* G = pool + fr.off + (long long)b0 * fr.ld + b0;
  const int tid = threadIdx.x;
  for (int e = tid; e < db * db; e += blockDim.x) {
    const int i = e % db, j = e / db;
    S[j * lds + i] = G[(long long)j * fr.ld + i];
  }
  __syncthreads();
  for (int s = 0; s < db; ++s) {
    const int jl = db - 1 - s;  // L column handled in this step (rows jl+1..db-1 use the inverted trailing block)
    const int ju = s;           // U column handled in this step (rows 0..ju-1 use the inverted leading block)
    for (int i = tid; i < db; i += blockDim.x) {
      if (i > jl) vl[i] = S[jl * lds + i];
      if (i < ju) vu[i] = S[ju * lds + i];
    }
    __syncthreads();
    for (int i = tid; i < db; i += blockDim.x) {
      if (i > jl) {  // y_i = v_i + Σ_{k=jl+1}^{i-1} X_ik v_k ;  X_i,jl = -y_i
        T y = vl[i];
        for (int k = jl + 1; k < i; ++k) y = hs_fma(y, S[k * lds + i], vl[k]);
        S[jl * lds + i] = hs_sub(hs_zero<T>(), y);
      }
    }
    // U: needs 1/U_jj of this column: every thread recomputes it (cheap, avoids a broadcast)
    const T xjj = hs_recip(S[ju * lds + ju]);
    __syncthreads();  // S[ju*lds+ju] read by all before thread ju/… overwrites it
    for (int i = tid; i < db; i += blockDim.x) {
      if (i < ju) {  // y_i = Σ_{k=i}^{ju-1} X_ik v_k ;  X_i,ju = -x_jj·y_i
        T y = hs_zero<T>();
        for (int k = i; k < ju; ++k) y = hs_fma(y, S[k * lds + i], vu[k]);
        S[ju * lds + i] = hs_sub(hs_zero<T>(), hs_mul(y, xjj));
      } else if (i == ju) {
        S[ju * lds + ju] = xjj;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < db * db; e += blockDim.x) {
    const int i = e % db, j = e / db;
    G[(long long)j * fr.ld + i] = S[j * lds + i];
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
// large fronts (ni > DB): one launch per block step, CTA per 32-row tile; the CTA that owns the next diagonal block
// (blockIdx.y == 0) also applies its inverse so that the next step can start from a final v_b.
// The working vector of every front lives in `work` (one slice per front, offset ioff − ioff0).
// ------------------------------------------------------------------------------------------------
// step = -1: gather v = [P x_int; x_bnd] into work and finalize block 0.  step = b ≥ 0: apply block column b.
template <typename T>
__global__ void __launch_bounds__(NTH) k_sv_fwd_step(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                      const int* __restrict__ gidx, const int* __restrict__ rperm,
                                                      T* __restrict__ x, long long ldx, T* __restrict__ work,
                                                      long long wstride, long long ioff0, int f0, int step) {
  constexpr int DB = SolveCfg<T>::DB;
  const Front fr = fronts[f0 + blockIdx.x];
  const int n = fr.n, ni = fr.ni;
  const T* F = pool + fr.off;
  const long long ld = fr.ld;
  T* xr = x + (long long)blockIdx.z * ldx;
  T* w = work + (long long)blockIdx.z * wstride + (fr.ioff - ioff0);
  const int* gi = gidx + fr.ioff;
  const int* rp = rperm + fr.ioff;
  __shared__ T vb[DB], u[DB];
  __shared__ T red[NW][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (step < 0) {
    // gather; tile 0 additionally finalizes block 0
    const int rows_per_cta = 32 * NW * 4;
    const int r_begin = blockIdx.y * rows_per_cta;
    if (r_begin >= n) return;
    for (int r = r_begin + tid; r < min(n, r_begin + rows_per_cta); r += NTH)
      if (!(blockIdx.y == 0 && r < DB)) w[r] = xr[gi[r < ni ? rp[r] : r]];
    if (blockIdx.y == 0) {
      const int db = min(DB, ni);
      for (int k = tid; k < db; k += NTH) u[k] = xr[gi[rp[k]]];
      __syncthreads();
      for (int r0 = 0; r0 < db; r0 += 32) {
        const int r = r0 + lane;
        const T s = tile_dot<T>(F, ld, r, r < db, 0, min(db, r0 + 32), u, 1, r, red);
        if (warp == 0 && r < db) w[r] = hs_add(u[r], s);  // x itself is written in step 0 (other CTAs still gather from it)
      }
    }
    return;
  }
  const int b0 = step * DB;
  if (b0 >= ni) return;
  const int b1 = min(b0 + DB, ni), db = b1 - b0;
  const bool last = b1 >= ni;
  // rows handled: blockIdx.y == 0 → the next diagonal block [b1, b1+DB) ∩ [b1, ni) (if any); others → 32-row tiles after it
  const int nb1 = last ? b1 : min(b1 + DB, ni);  // end of the next diagonal block
  for (int k = tid; k < db; k += NTH) vb[k] = w[b0 + k];
  __syncthreads();
  if (blockIdx.y == 0) {
    for (int k = tid; k < db; k += NTH) xr[gi[b0 + k]] = vb[k];  // block b is final: t = L11⁻¹·P·x[int]
    if (last) return;
    const int dn = nb1 - b1;
    for (int r0 = 0; r0 < dn; r0 += 32) {
      const int r = b1 + r0 + lane;
      const T s = tile_dot<T>(F + (long long)b0 * ld, ld, r, r < nb1, 0, db, vb, 0, 0, red);
      if (warp == 0 && r < nb1) u[r0 + lane] = hs_sub(w[r], s);
    }
    __syncthreads();
    for (int r0 = 0; r0 < dn; r0 += 32) {
      const int r = r0 + lane;
      const T s = tile_dot<T>(F + (long long)b1 * ld + b1, ld, r, r < dn, 0, min(dn, r0 + 32), u, 1, r, red);
      if (warp == 0 && r < dn) w[b1 + r] = hs_add(u[r], s);
    }
    return;
  }
  const int r0 = nb1 + (blockIdx.y - 1) * 32;
  if (r0 >= n) return;
  const int r = r0 + lane;
  const T s = tile_dot<T>(F + (long long)b0 * ld, ld, r, r < n, 0, db, vb, 0, 0, red);
  if (warp == 0 && r < n) {
    const T v = hs_sub(w[r], s);
    if (last) xr[gi[r]] = v; else w[r] = v;   // after the last block column the boundary rows are final
  }
}

// step = -1: t = x_int − U12·x_bnd into work (all rows), the CTA owning the last diagonal block applies its inverse.
// step = b ≥ 1 (descending): t[0:b0] -= F[0:b0, b]·t_b, the CTA owning block b-1 applies its inverse.
template <typename T>
__global__ void __launch_bounds__(NTH) k_sv_bwd_step(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                      const int* __restrict__ gidx, T* __restrict__ x, long long ldx,
                                                      T* __restrict__ work, long long wstride, long long ioff0, int f0,
                                                      int step) {
  constexpr int DB = SolveCfg<T>::DB;
  const Front fr = fronts[f0 + blockIdx.x];
  const int n = fr.n, ni = fr.ni, nb = n - ni;
  const T* F = pool + fr.off;
  const long long ld = fr.ld;
  T* xr = x + (long long)blockIdx.z * ldx;
  T* w = work + (long long)blockIdx.z * wstride + (fr.ioff - ioff0);
  const int* gi = gidx + fr.ioff;
  __shared__ T vb[DB], u[DB];
  __shared__ T red[NW][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = (ni + DB - 1) / DB;
  if (step < 0) {
    // rows of the last diagonal block go to blockIdx.y == 0, the rest in 32-row tiles
    const int l0 = (B - 1) * DB, dl = ni - l0;
    const bool diag = blockIdx.y == 0;
    const int rbase = diag ? l0 : (blockIdx.y - 1) * 32;
    const int rend = diag ? ni : min(l0, rbase + 32);
    if (rbase >= rend) return;
    // U12·x_bnd in chunks of DB boundary entries
    T acc[(DB + 31) / 32];
#pragma unroll
    for (int q = 0; q < (DB + 31) / 32; ++q) acc[q] = hs_zero<T>();
    for (int c0 = 0; c0 < nb; c0 += DB) {
      const int cw = min(DB, nb - c0);
      __syncthreads();
      for (int k = tid; k < cw; k += NTH) vb[k] = xr[gi[ni + c0 + k]];
      __syncthreads();
      int q = 0;
      for (int r0 = rbase; r0 < rend; r0 += 32, ++q) {
        const int r = r0 + lane;
        const T s = tile_dot<T>(F + (long long)(ni + c0) * ld, ld, r, r < rend, 0, cw, vb, 0, 0, red);
        if (warp == 0) acc[q] = hs_add(acc[q], s);
      }
    }
    int q = 0;
    for (int r0 = rbase; r0 < rend; r0 += 32, ++q) {
      const int r = r0 + lane;
      if (warp == 0 && r < rend) {
        const T v = hs_sub(xr[gi[r]], acc[q]);
        if (diag) u[r - l0] = v; else w[r] = v;
      }
    }
    if (!diag) return;
    __syncthreads();
    for (int r0 = 0; r0 < dl; r0 += 32) {
      const int r = r0 + lane;
      const T s = tile_dot<T>(F + (long long)(l0 + r0) * ld + l0, ld, r, r < dl, 0, dl - r0, u + r0, 2, lane, red);
      if (warp == 0 && r < dl) { w[l0 + r] = s; xr[gi[l0 + r]] = s; }
    }
    return;
  }
  // apply block column `step` (≥ 1) to the rows above it
  const int b0 = step * DB;
  if (b0 >= ni || step < 1) return;
  const int b1 = min(b0 + DB, ni), db = b1 - b0;
  const int p0 = b0 - DB;  // previous diagonal block [p0, b0)
  for (int k = tid; k < db; k += NTH) vb[k] = w[b0 + k];
  __syncthreads();
  if (blockIdx.y == 0) {
    for (int r0 = 0; r0 < DB; r0 += 32) {
      const int r = p0 + r0 + lane;
      const T s = tile_dot<T>(F + (long long)b0 * ld, ld, r, true, 0, db, vb, 0, 0, red);
      if (warp == 0) u[r0 + lane] = hs_sub(w[r], s);
    }
    __syncthreads();
    for (int r0 = 0; r0 < DB; r0 += 32) {
      const int r = r0 + lane;
      const T s = tile_dot<T>(F + (long long)(p0 + r0) * ld + p0, ld, r, true, 0, DB - r0, u + r0, 2, lane, red);
      if (warp == 0) { w[p0 + r] = s; xr[gi[p0 + r]] = s; }
    }
    return;
  }
  const int r0 = (blockIdx.y - 1) * 32;
  if (r0 >= p0) return;
  const int r = r0 + lane;
  const T s = tile_dot<T>(F + (long long)b0 * ld, ld, r, r < p0, 0, db, vb, 0, 0, red);
  if (warp == 0 && r < p0) w[r] = hs_sub(w[r], s);
}

template <typename T> void prep_impl(hs_fac* f, const Level& L) {
  constexpr int DB = SolveCfg<T>::DB;
  if (L.max_ni == 0) return;
  const int db = std::min(DB, L.max_ni);
  const size_t sm = ((size_t)db * (db | 1) + 2 * db) * sizeof(T);
  dim3 grid(L.f1 - L.f0, (L.max_ni + DB - 1) / DB);
  k_trtri_diag<T><<<grid, 128, sm, f->ctx->stream>>>(f->d_fronts, (T*)f->pool, L.f0);
  CUDA_OK(cudaGetLastError());
  f->stats.launches_factor += 1;
}

template <typename T> void run_impl(hs_fac* f, int64_t nrhs, void* xv) {
  constexpr int DB = SolveCfg<T>::DB;
  cudaStream_t st = f->ctx->stream;
  T* x = (T*)xv;
  const T* pool = (const T*)f->pool;
  T* work = (T*)f->d_work;
  const long long ws = f->max_level_idx;
  hs_stats_t& s = f->stats;
  auto nbig_of = [&](const Level& L) {
    return (int)(std::partition_point(L.ni_sorted.begin(), L.ni_sorted.end(), [&](int v) { return v > DB; }) - L.ni_sorted.begin());
  };
  for (size_t li = 0; li < f->levels.size(); ++li) {  // post-order
    const Level& L = f->levels[li];
    const int nf = L.f1 - L.f0, nbig = nbig_of(L);
    if (nbig > 0) {
      const int B = (L.max_ni + DB - 1) / DB;
      dim3 g0(nbig, (L.max_n + 32 * NW * 4 - 1) / (32 * NW * 4), (unsigned)nrhs);
      k_sv_fwd_step<T><<<g0, NTH, 0, st>>>(f->d_fronts, pool, f->d_gidx, f->d_rperm, x, f->n, work, ws, L.ioff0, L.f0, -1);
      ++s.launches_solve;
      for (int b = 0; b < B; ++b) {
        const int nact = (int)(std::partition_point(L.ni_sorted.begin(), L.ni_sorted.begin() + nbig, [&](int v) { return v > b * DB; }) -
                               L.ni_sorted.begin());
        if (nact == 0) break;
        const int rows_after = std::max(0, L.max_n - b * DB);
        dim3 g(nact, 1 + (rows_after + 31) / 32, (unsigned)nrhs);
        k_sv_fwd_step<T><<<g, NTH, 0, st>>>(f->d_fronts, pool, f->d_gidx, f->d_rperm, x, f->n, work, ws, L.ioff0, L.f0, b);
        ++s.launches_solve;
      }
    }
    if (nf - nbig > 0) {
      dim3 g(nf - nbig, (unsigned)nrhs);
      k_sv_small_fwd<T><<<g, NTH, 0, st>>>(f->d_fronts, pool, f->d_gidx, f->d_rperm, x, f->n, L.f0 + nbig);
      ++s.launches_solve;
    }
  }
  for (size_t li = f->levels.size(); li-- > 0;) {  // pre-order
    const Level& L = f->levels[li];
    const int nf = L.f1 - L.f0, nbig = nbig_of(L);
    if (nbig > 0) {
      const int B = (L.max_ni + DB - 1) / DB;
      dim3 g0(nbig, 1 + (L.max_ni + 31) / 32, (unsigned)nrhs);
      k_sv_bwd_step<T><<<g0, NTH, 0, st>>>(f->d_fronts, pool, f->d_gidx, x, f->n, work, ws, L.ioff0, L.f0, -1);
      ++s.launches_solve;
      for (int b = B - 1; b >= 1; --b) {
        const int nact = (int)(std::partition_point(L.ni_sorted.begin(), L.ni_sorted.begin() + nbig, [&](int v) { return v > b * DB; }) -
                               L.ni_sorted.begin());
        if (nact == 0) continue;
        dim3 g(nact, 1 + ((b - 1) * DB + 31) / 32, (unsigned)nrhs);
        k_sv_bwd_step<T><<<g, NTH, 0, st>>>(f->d_fronts, pool, f->d_gidx, x, f->n, work, ws, L.ioff0, L.f0, b);
        ++s.launches_solve;
      }
    }
    if (nf - nbig > 0) {
      dim3 g(nf - nbig, (unsigned)nrhs);
      const size_t sm = (size_t)std::max(L.max_nb, 1) * sizeof(T);
      k_sv_small_bwd<T><<<g, NTH, sm, st>>>(f->d_fronts, pool, f->d_gidx, x, f->n, L.f0 + nbig);
      ++s.launches_solve;
    }
  }
  CUDA_OK(cudaGetLastError());
}

}  // namespace

void hs_solve_setup() {
  CUDA_OK(cudaFuncSetAttribute(k_trtri_diag<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_trtri_diag<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_sv_small_bwd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_sv_small_bwd<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

int hs_solve_block(hs_dtype dt) { return dt == HS_F64 ? SolveCfg<double>::DB : SolveCfg<cplx>::DB; }

void hs_solve_prep(hs_fac* f, const Level& L) {
  if (f->dtype == HS_F64) prep_impl<double>(f, L); else prep_impl<cplx>(f, L);
}

void hs_solve_run(hs_fac* f, int64_t nrhs, void* x) {
  if (f->dtype == HS_F64) run_impl<double>(f, nrhs, x); else run_impl<cplx>(f, nrhs, x);
}
