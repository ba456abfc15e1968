// Compressed fronts: low-rank Gauss transforms (reference src/factorization.jl:78-112,171-182,228-249).
//
// A front flagged for compression (factorization.jl:15) is assembled densely like every other front.  Then
//   1. Abi ≈ Qb·Rb and Aib ≈ Qi·Ri by truncated column-pivoted QR (`pqrfact(...; sketch=:none)`, :172,:178), computed
//      here as a pivoted Cholesky factorization of the Gram matrix MᴴM evaluated lazily one pivot row per step —
//      the same pivots and the same R as Householder QR with column pivoting in exact arithmetic, but made of
//      full-width streaming passes over M (HBM-bound, every SM busy) instead of a serial chain of reflectors;
//      the stopping rule is pqrfact's: first k with |R[k,k]| ≤ max(atol, rtol·|R[1,1]|).  Q = M[:,piv]·R11⁻¹.
//   2. the thin bordered front [Aii Qi; Rb 0] is built and handed to the ordinary level-batched LU (hs_api.cu);
//   3. the Schur complement operator of :228-249, S = Abb − (Abi·(Aii⁻¹Qi))·Ri, is evaluated with the tensor-core
//      GEMM (descriptor mode of k_gemm) and left in the front's dense slot, where the parent's assembly finds it.
// The accuracy of the lazily evaluated Gram matrix limits useful tolerances to about 1e-7 relative.
#include <algorithm>
#include <cstring>
#include <vector>

#include "hs_fac.cuh"
#include "hs_kernels.cuh"

namespace {

__device__ __forceinline__ double abs2(double a) { return a * a; }
__device__ __forceinline__ double abs2(cplx a) { return a.x * a.x + a.y * a.y; }
// acc + conj(a)·b
__device__ __forceinline__ double cjfma(double acc, double a, double b) { return fma(a, b, acc); }
__device__ __forceinline__ cplx cjfma(cplx acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
  return acc;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ cplx wsum(cplx v) { return cplx{wsum(v.x), wsum(v.y)}; }
__device__ __forceinline__ double scal(double a, double s) { return a * s; }
__device__ __forceinline__ cplx scal(cplx a, double s) { return cplx{a.x * s, a.y * s}; }
__device__ __forceinline__ double from_real(double r, double*) { return r; }
__device__ __forceinline__ cplx from_real(double r, cplx*) { return cplx{r, 0.0}; }
__device__ __forceinline__ double real_of(double a) { return a; }
__device__ __forceinline__ double real_of(cplx a) { return a.x; }

// ---- 1. truncated column-pivoted QR through the Gram matrix ----------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_id_init(const IdRun* __restrict__ runs, const T* __restrict__ pool,
                                                  double* __restrict__ state, int* __restrict__ ints) {
  const IdRun R = runs[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.y == 0 && threadIdx.x == 0) ints[1 + R.slot] = -1;
  for (int j = blockIdx.y * 8 + warp; j < R.ncol; j += gridDim.y * 8) {
    const T* c = pool + R.moff + (long long)j * R.ld;
    double s = 0.0;
    for (int i = lane; i < R.m; i += 32) s += abs2(c[i]);
    s = wsum(s);
    if (lane == 0) state[R.st + j] = s;
  }
}

// one step k of every unfinished run: pick the column with the largest remaining norm, produce row k of R for all
// columns and downdate the norms.  Every CTA of a run repeats the (cheap) pivot search so that no grid-wide barrier
// is needed; the column norms are double-buffered by the parity of k.
// STAGE: the pivot column (and its R entries) is staged in shared memory; for fronts too tall for that it is read
// through L2 instead.
template <typename T, bool STAGE>
__global__ void __launch_bounds__(256) k_id_step(const IdRun* __restrict__ runs, const T* __restrict__ pool, T* __restrict__ ws,
                                                  double* __restrict__ state, int* __restrict__ ints, int k, double atol,
                                                  double rtol) {
  const IdRun R = runs[blockIdx.x];
  int* rank = ints + 1 + R.slot;
  if (*rank >= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool lead = blockIdx.y == 0 && tid == 0;
  if (k >= R.rcap) {
    if (lead) { *rank = R.rcap; atomicAdd(ints, 1); }
    return;
  }
  extern __shared__ __align__(16) unsigned char smem_id[];
  T* scol = reinterpret_cast<T*>(smem_id);
  T* srp = scol + (STAGE ? R.m : 0);
  __shared__ double s_v[8];
  __shared__ int s_i[8];
  __shared__ double s_red[8];
  const double* din = state + R.st + (long long)(k & 1) * R.ncol;
  double* dout = state + R.st + (long long)((k + 1) & 1) * R.ncol;
  double best = -1.0;
  int bj = 0x7fffffff;
  for (int j = tid; j < R.ncol; j += 256) {
    const double v = din[j];
    if (v > best) { best = v; bj = j; }
  }
  warp_argmax(best, bj);
  if (lane == 0) { s_v[warp] = best; s_i[warp] = bj; }
  __syncthreads();
  best = s_v[0]; bj = s_i[0];
#pragma unroll
  for (int w = 1; w < 8; ++w)
    if (s_v[w] > best || (s_v[w] == best && s_i[w] < bj)) { best = s_v[w]; bj = s_i[w]; }
  if (!(best > 0.0)) {
    if (lead) { *rank = k; atomicAdd(ints, 1); }
    return;
  }
  const int p = bj;
  const T* Mp = pool + R.moff + (long long)p * R.ld;
  const T* Rp = ws + R.rws + (long long)p * R.ldr;
  double part = 0.0;
  for (int i = tid; i < R.m; i += 256) { const T v = Mp[i]; if (STAGE) scol[i] = v; part += abs2(v); }
  for (int l = tid; l < k; l += 256) { const T v = Rp[l]; if (STAGE) srp[l] = v; part -= abs2(v); }
  const T* colp = STAGE ? scol : Mp;
  const T* rp = STAGE ? srp : Rp;
  part = wsum(part);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  double y = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) y += s_red[w];
  const double rkk = sqrt(fmax(y, 0.0));                       // |R[k,k]|, recomputed rather than downdated
  const double r11 = k == 0 ? rkk : state[R.st + 2ll * R.ncol];
  const double tscale = R.pad ? 0.5 : 1.0;   // the right transform's sketch of the sparse couplings halves once more (:202)
  const double ptol = fmax(tscale * atol, tscale * rtol * r11);
  if (!(rkk > ptol)) {
    if (lead) { *rank = k; atomicAdd(ints, 1); }
    return;
  }
  if (lead) {
    if (k == 0) state[R.st + 2ll * R.ncol] = rkk;
    ints[R.ip + k] = p;
  }
  const double inv = 1.0 / rkk;
  // two columns per warp and trip: twice the loads in flight
  const int stride = gridDim.y * 8;
  for (int ja = blockIdx.y * 8 + warp; ja < R.ncol; ja += 2 * stride) {
    const int jb = ja + stride;
    const bool hb = jb < R.ncol;
    const double da = din[ja], db = hb ? din[jb] : -1.0;
    T* Ra = ws + R.rws + (long long)ja * R.ldr;
    T* Rb = ws + R.rws + (long long)(hb ? jb : ja) * R.ldr;
    const bool sa = da < 0.0 || ja == p, sb = !hb || db < 0.0 || jb == p;  // earlier pivots / this step's pivot
    T acca = hs_zero<T>(), accb = hs_zero<T>();
    if (!sa && !sb) {
      const T* Ma = pool + R.moff + (long long)ja * R.ld;
      const T* Mb = pool + R.moff + (long long)jb * R.ld;
      T a2 = hs_zero<T>(), b2 = hs_zero<T>();
#pragma unroll 4
      for (int i = lane; i < R.m; i += 32) { const T c = colp[i]; acca = cjfma(acca, c, Ma[i]); accb = cjfma(accb, c, Mb[i]); }
      for (int l = lane; l < k; l += 32) { const T c = rp[l]; a2 = cjfma(a2, c, Ra[l]); b2 = cjfma(b2, c, Rb[l]); }
      acca = hs_sub(acca, a2); accb = hs_sub(accb, b2);
    } else {
      if (!sa) {
        const T* Ma = pool + R.moff + (long long)ja * R.ld;
        T a2 = hs_zero<T>();
#pragma unroll 4
        for (int i = lane; i < R.m; i += 32) acca = cjfma(acca, colp[i], Ma[i]);
        for (int l = lane; l < k; l += 32) a2 = cjfma(a2, rp[l], Ra[l]);
        acca = hs_sub(acca, a2);
      }
      if (!sb) {
        const T* Mb = pool + R.moff + (long long)jb * R.ld;
        T b2 = hs_zero<T>();
#pragma unroll 4
        for (int i = lane; i < R.m; i += 32) accb = cjfma(accb, colp[i], Mb[i]);
        for (int l = lane; l < k; l += 32) b2 = cjfma(b2, rp[l], Rb[l]);
        accb = hs_sub(accb, b2);
      }
    }
    acca = wsum(acca); accb = wsum(accb);
    if (lane == 0) {
      if (sa) { Ra[k] = ja == p ? from_real(rkk, (T*)nullptr) : hs_zero<T>(); dout[ja] = -1.0; }
      else { const T v = scal(acca, inv); Ra[k] = v; dout[ja] = fmax(da - abs2(v), 0.0); }
      if (hb) {
        if (sb) { Rb[k] = jb == p ? from_real(rkk, (T*)nullptr) : hs_zero<T>(); dout[jb] = -1.0; }
        else { const T v = scal(accb, inv); Rb[k] = v; dout[jb] = fmax(db - abs2(v), 0.0); }
      }
    }
  }
}

// Q = M[:, piv]·R11⁻¹, one thread per row (coalesced over rows); Qb goes to the side buffer, Qi straight into the
// thin front's border columns.
template <typename T>
__global__ void __launch_bounds__(128) k_form_q(const IdRun* __restrict__ runs, const LrDesc* __restrict__ lr, T* pool,
                                                 const T* __restrict__ ws, const int* __restrict__ ints,
                                                 const Front* __restrict__ fronts) {
  const IdRun R = runs[blockIdx.x];
  const LrDesc D = lr[blockIdx.x >> 1];
  const int which = blockIdx.x & 1;
  const int r = which ? D.r2x : D.r1x;
  const int i = blockIdx.y * 128 + threadIdx.x;
  if (i >= R.m) return;
  T* Q;
  long long ldq;
  if (!which) { ldq = D.qb_ld; Q = pool + D.qb + (long long)D.off1 * ldq; }
  else { const Front th = fronts[D.thin]; ldq = th.ld; Q = pool + th.off + (long long)(D.ni + D.off2) * ldq; }
  const int* piv = ints + R.ip;
  const T* M = pool + R.moff;
  for (int q = 0; q < r; ++q) {
    const int p = piv[q];
    const T* Rp = ws + R.rws + (long long)p * R.ldr;
    T v = M[(long long)p * R.ld + i];
    for (int l = 0; l < q; ++l) v = hs_fnma(v, Q[(long long)l * ldq + i], Rp[l]);
    Q[(long long)q * ldq + i] = scal(v, 1.0 / real_of(Rp[q]));
  }
}

// thin front = [Aii Qi; Rb 0] (Qi comes from k_form_q), and the persistent copies of Ri and Riᵀ
template <typename T>
__global__ void __launch_bounds__(256) k_build_thin(const LrDesc* __restrict__ lr, const IdRun* __restrict__ runs,
                                                     const Front* __restrict__ fronts, T* pool, const T* __restrict__ ws) {
  const int c = blockIdx.x;
  const LrDesc D = lr[c];
  const Front fd = fronts[D.fi], th = fronts[D.thin];
  const IdRun Rb = runs[2 * c], Ri = runs[2 * c + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nmax = max(D.ni, D.nb);
  for (int j = blockIdx.y * 8 + warp; j < nmax; j += gridDim.y * 8) {
    if (j < D.ni) {
      const T* src = pool + fd.off + (long long)j * fd.ld;
      T* dst = pool + th.off + (long long)j * th.ld;
      for (int i = lane; i < D.ni; i += 32) dst[i] = src[i];
      const T* rb = ws + Rb.rws + (long long)j * Rb.ldr;
      for (int l = lane; l < D.r1x; l += 32) dst[D.ni + D.off1 + l] = rb[l];
    }
    if (j < D.nb) {
      const T* rr = ws + Ri.rws + (long long)j * Ri.ldr;
      for (int l = lane; l < D.r2x; l += 32) {
        const T v = rr[l];
        pool[D.ri + (long long)j * D.ri_ld + D.off2 + l] = v;
        pool[D.vit + (long long)(D.off2 + l) * D.qb_ld + j] = v;
      }
    }
  }
}

// HSS-children fronts (factorization.jl:184-209): the pivoted QR runs on the sparse couplings alone — the anti-diagonal
// blocks A[bnd1,int2], A[bnd2,int1] of Abi and A[int1,bnd2], A[int2,bnd1] of Aib — copied out of the assembled front
template <typename T>
__global__ void __launch_bounds__(256) k_copy_anti(const LrDesc* __restrict__ lr, const Front* __restrict__ fronts, T* pool) {
  const LrDesc D = lr[blockIdx.x];
  if (!D.hchild) return;
  const Front fd = fronts[D.fi];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* F = pool + fd.off;
  for (int j = blockIdx.y * 8 + warp; j < D.ni + D.nb; j += gridDim.y * 8) {
    if (j < D.ni) {   // column j of X1 (nb × ni): rows of the other child only
      T* dst = pool + D.x1 + (long long)j * D.ldx1;
      const T* src = F + (long long)j * fd.ld + D.ni;
      const bool cl = j < D.ni_l;
      for (int r = lane; r < D.nb; r += 32) dst[r] = ((r < D.nb_l) != cl) ? src[r] : hs_zero<T>();
    } else {          // column j − ni of X2 (ni × nb)
      const int c = j - D.ni;
      T* dst = pool + D.x2 + (long long)c * D.ldx2;
      const T* src = F + (long long)(D.ni + c) * fd.ld;
      const bool cl = c < D.nb_l;
      for (int r = lane; r < D.ni; r += 32) dst[r] = ((r < D.ni_l) != cl) ? src[r] : hs_zero<T>();
    }
  }
}

// ---- 3. Schur complement ----------------------------------------------------------------------------------------
// Y = L11⁻¹·P·Qi (border columns of the factored thin front) → workspace, ld = ldy
template <typename T>
__global__ void __launch_bounds__(256) k_copy_y(const LrDesc* __restrict__ lr, const IdRun* __restrict__ runs,
                                                 const Front* __restrict__ fronts, T* pool, T* __restrict__ ws) {
  const int c = blockIdx.x;
  const LrDesc D = lr[c];
  const Front th = fronts[D.thin];
  const int ldy = max(2, (D.ni + 1) & ~1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = blockIdx.y * 8 + warp; q < D.r2; q += gridDim.y * 8) {
    const T* src = pool + th.off + (long long)(D.ni + q) * th.ld;
    T* dst = pool + D.y + (long long)q * ldy;
    for (int i = lane; i < D.ni; i += 32) dst[i] = src[i];
  }
}

// rows [b0, b0+db) of Y ← U_bb⁻¹ · (those rows); U_bb⁻¹ is the upper triangle the solve preparation left in the
// diagonal block.  One warp per column of Y.
template <typename T>
__global__ void __launch_bounds__(128) k_diag_apply(const LrDesc* __restrict__ lr, const IdRun* __restrict__ runs,
                                                     const Front* __restrict__ fronts, T* pool, T* __restrict__ ws,
                                                     int b) {
  const int c = blockIdx.x;
  const LrDesc D = lr[c];
  const int b0 = b * 64;
  if (b0 >= D.ni) return;
  const int db = min(64, D.ni - b0);
  const Front th = fronts[D.thin];
  const T* U = pool + th.off + (long long)b0 * th.ld + b0;
  const int ldy = max(2, (D.ni + 1) & ~1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ T ys[4][64];
  for (int q = blockIdx.y * 4 + warp; q < D.r2; q += gridDim.y * 4) {
    T* y = pool + D.y + (long long)q * ldy + b0;
    ys[warp][lane] = lane < db ? y[lane] : hs_zero<T>();
    ys[warp][lane + 32] = lane + 32 < db ? y[lane + 32] : hs_zero<T>();
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = lane + 32 * h;
      if (i < db) {
        T acc = hs_zero<T>();
        for (int kk = i; kk < db; ++kk) acc = hs_fma(acc, U[(long long)kk * th.ld + i], ys[warp][kk]);
        y[i] = acc;
      }
    }
    __syncwarp();
  }
}

// ---- solve: the two thin products around the ordinary sweeps over the thin front -----------------------------------
// forward: x[bnd] += Qb · v, v = the border slots after the thin front's forward step (= −Rb·Aii⁻¹·x_int)
template <typename T>
__global__ void __launch_bounds__(256) k_lr_fwd(const LrDesc* __restrict__ lr, const Front* __restrict__ fronts,
                                                 const T* pool, const int* __restrict__ gidx, T* x, long long ldx,
                                                 long long n) {
  const LrDesc D = lr[blockIdx.x];
  if (blockIdx.y * 256 >= D.nb) return;
  const Front fd = fronts[D.fi];
  T* xr = x + (long long)blockIdx.z * ldx;
  extern __shared__ __align__(16) unsigned char smem_lr[];
  T* sv = reinterpret_cast<T*>(smem_lr);
  for (int k = threadIdx.x; k < D.r1; k += 256) sv[k] = xr[n + D.voff + k];
  __syncthreads();
  const int i = blockIdx.y * 256 + threadIdx.x;
  if (i >= D.nb) return;
  const T* Q = pool + D.qb + i;
  T acc = hs_zero<T>();
  for (int k = 0; k < D.r1; ++k) acc = hs_fma(acc, Q[(long long)k * D.qb_ld], sv[k]);
  const int g = gidx[fd.ioff + D.ni + i];
  xr[g] = hs_add(xr[g], acc);
}

// backward: border slots ← Ri · x[bnd]; the thin front's backward step then subtracts (L11⁻¹PQi)·that
template <typename T>
__global__ void __launch_bounds__(256) k_lr_bwd(const LrDesc* __restrict__ lr, const Front* __restrict__ fronts,
                                                 const T* pool, const int* __restrict__ gidx, T* x, long long ldx,
                                                 long long n) {
  const LrDesc D = lr[blockIdx.x];
  const Front fd = fronts[D.fi];
  T* xr = x + (long long)blockIdx.z * ldx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int* gb = gidx + fd.ioff + D.ni;
  for (int k = blockIdx.y * 8 + warp; k < D.r; k += gridDim.y * 8) {
    T acc = hs_zero<T>();
    if (k < D.r2) {
      const T* V = pool + D.vit + (long long)k * D.qb_ld;
      for (int j = lane; j < D.nb; j += 32) acc = hs_fma(acc, V[j], xr[gb[j]]);
      acc = wsum(acc);
    }
    if (lane == 0) xr[n + D.voff + k] = acc;
  }
}

template <typename T> void gen_gemm(hs_fac* f, const GemmDesc* d_items, int nitems, int maxM, int maxN) {
  if (nitems <= 0 || maxM <= 0 || maxN <= 0) return;
  using Cfg = GemmCfg<T>;
  dim3 grid(nitems, (maxM + Cfg::TM - 1) / Cfg::TM, (maxN + Cfg::TN - 1) / Cfg::TN);
  k_gemm<T, true><<<grid, gemm_threads<T>(), gemm_smem_bytes<T>(), f->ctx->stream>>>(nullptr, (T*)f->pool, 0, 0, 0, 0, 0, 0, d_items);
  CUDA_OK(cudaGetLastError());
  ++f->stats.launches_factor;
  ++f->stats.gemm_launches;
}

inline long long up32(long long v) { return (v + 31) / 32 * 32; }
inline int even_up(int v) { return std::max(2, (v + 1) & ~1); }

template <typename T> void prepare_impl(hs_fac* f, CompLevel& C) {
  cudaStream_t st = f->ctx->stream;
  const int nc = C.c1 - C.c0, nruns = 2 * nc;
  if (nc == 0) return;
  const IdRun* d_runs = f->d_runs + 2 * C.c0;
  const T* pool = (const T*)f->pool;
  T* ws = (T*)f->d_cws;
  int max_ncol = 0, max_m = 0, max_rcap = 0;
  bool any_hchild = false;
  for (int c = C.c0; c < C.c1; ++c) {
    const CompFront& cf = f->comp[c];
    max_ncol = std::max(max_ncol, std::max(cf.ni, cf.nb));
    max_m = std::max(max_m, std::max(cf.ni, cf.nb));
    max_rcap = std::max(max_rcap, cf.rcap);
    any_hchild = any_hchild || cf.hchild;
  }
  std::vector<LrDesc> lr(nc);
  for (int c = C.c0; c < C.c1; ++c) {   // static part; ranks and buffer offsets follow once the pivoted QRs are done
    const CompFront& cf = f->comp[c];
    LrDesc& d = lr[c - C.c0];
    d = LrDesc{};
    d.fi = cf.fi; d.thin = f->nfr + cf.fi; d.ni = cf.ni; d.nb = cf.nb; d.voff = cf.voff;
    d.hchild = cf.hchild; d.ni_l = cf.ni_l; d.nb_l = cf.nb_l; d.x1 = cf.x1; d.x2 = cf.x2; d.ldx1 = cf.ldx1; d.ldx2 = cf.ldx2;
  }
  if (any_hchild) {
    // children in HSS form: their low-rank blocks are taken over as they are, only the sparse couplings are compressed
    CUDA_OK(cudaMemcpyAsync(f->d_lr + C.c0, lr.data(), (size_t)nc * sizeof(LrDesc), cudaMemcpyHostToDevice, st));
    dim3 g(nc, std::min((max_m * 2 + 7) / 8, 128));
    k_copy_anti<T><<<g, 256, 0, st>>>(f->d_lr + C.c0, f->d_fronts, (T*)f->pool);
    CUDA_OK(cudaGetLastError());
    ++f->stats.launches_factor;
  }
  size_t smem = (size_t)(max_m + max_rcap) * sizeof(T);
  static const bool no_stage = getenv("HS_ID_NOSTAGE") != nullptr;
  const bool stage = smem <= 200 * 1024 && !no_stage;  // taller fronts read the pivot column through L2
  if (!stage) smem = 0;
  CUDA_OK(cudaMemsetAsync(f->d_cint, 0, sizeof(int), st));
  {
    dim3 g(nruns, std::min((max_ncol + 7) / 8, 64));
    k_id_init<T><<<g, 256, 0, st>>>(d_runs, pool, f->d_cstate, f->d_cint);
    CUDA_OK(cudaGetLastError());
    ++f->stats.launches_factor;
  }
  const int ny = std::max(1, std::min((max_ncol + 15) / 16, (6 * 148 + nruns - 1) / nruns));
  const double atol = 0.5 * f->opts.atol, rtol = 0.5 * f->opts.rtol;  // factorization.jl:99-100
  int k = 0, done = 0;
  while (done < nruns && k <= max_rcap) {
    const int ke = std::min(k + 32, max_rcap + 1);
    for (; k < ke; ++k) {
      if (stage) k_id_step<T, true><<<dim3(nruns, ny), 256, smem, st>>>(d_runs, pool, ws, f->d_cstate, f->d_cint, k, atol, rtol);
      else k_id_step<T, false><<<dim3(nruns, ny), 256, 0, st>>>(d_runs, pool, ws, f->d_cstate, f->d_cint, k, atol, rtol);
      ++f->stats.launches_factor;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(&done, f->d_cint, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
  std::vector<int> ranks(nruns);
  CUDA_OK(cudaMemcpyAsync(ranks.data(), f->d_cint + 1, (size_t)nruns * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  // side buffer: thin fronts, Qb, Riᵀ, Ri, Y, Z
  Level& Lt = f->flevels[C.flevel];
  long long off = 0;
  Lt.max_n = 0; Lt.max_nb = 0;
  std::vector<long long> thin_off(nc);
  for (int c = C.c0; c < C.c1; ++c) {
    CompFront& cf = f->comp[c];
    LrDesc& d = lr[c - C.c0];
    cf.rx1 = ranks[2 * (c - C.c0)];
    cf.rx2 = ranks[2 * (c - C.c0) + 1];
    if (cf.rx1 < 0 || cf.rx2 < 0) throw hs_error(HS_ECUDA, "pivoted QR did not terminate");
    d.off1 = d.off2 = 0;
    if (cf.hchild) {
      // L = [blkdiag(Ub1, Ub2)  Q]·[blkdiag(Vi1, Vi2)  Rᴴ]ᴴ (factorization.jl:185-191), R likewise (:198-204): the ranks add up
      const HssFront &Hl = f->hss[f->comp[cf.cl].hss], &Hr = f->hss[f->comp[cf.cr].hss];
      d.off1 = Hl.ra1 + Hr.ra1;
      d.off2 = Hl.rb1 + Hr.rb1;
    }
    d.r1x = cf.rx1; d.r2x = cf.rx2;
    cf.r1 = d.off1 + cf.rx1;
    cf.r2 = d.off2 + cf.rx2;
    cf.r = std::max(cf.r1, cf.r2);
    if (cf.r > cf.vcap) throw hs_error(HS_ESIZE, "low-rank Gauss transform of a front exceeds its reserved border rows");
    f->stats.maxrank = std::max<int64_t>(f->stats.maxrank, cf.r);
    const int nt = cf.ni + cf.r, ldt = even_up(nt);
    thin_off[c - C.c0] = off; off += up32((long long)ldt * nt);
    cf.qb_ld = even_up(cf.nb);
    cf.ri_ld = even_up(cf.r2);
    cf.qb = off; off += up32((long long)cf.qb_ld * std::max(cf.r1, 1));
    cf.vit = off; off += up32((long long)cf.qb_ld * std::max(cf.r2, 1));
    cf.ri = off; off += up32((long long)cf.ri_ld * cf.nb);
    cf.yoff = off; off += up32((long long)even_up(cf.ni) * std::max(cf.r2, 1));
    cf.zoff = off; off += up32((long long)even_up(cf.nb) * std::max(cf.r2, 1));
    Lt.max_n = std::max(Lt.max_n, nt);
    Lt.max_nb = std::max(Lt.max_nb, cf.r);
  }
  const size_t need = (size_t)std::max<long long>(off, 32) * sizeof(T);
  if (need > C.side_bytes) {
    cudaFree(C.side);
    C.side = nullptr; C.side_bytes = 0;
    CUDA_OK(cudaMalloc(&C.side, need));
    C.side_bytes = need;
  }
  CUDA_OK(cudaMemsetAsync(C.side, 0, need, st));
  const long long base = (long long)(((char*)C.side - (char*)f->pool) / (long long)sizeof(T));
  for (int c = C.c0; c < C.c1; ++c) {
    CompFront& cf = f->comp[c];
    cf.qb += base; cf.vit += base; cf.ri += base; cf.yoff += base; cf.zoff += base;
    Front& th = f->fronts[f->nfr + cf.fi];
    th.n = cf.ni + cf.r;
    th.ld = even_up(th.n);
    th.off = base + thin_off[c - C.c0];
    LrDesc& d = lr[c - C.c0];
    d.r1 = cf.r1; d.r2 = cf.r2; d.r = cf.r;
    d.qb = cf.qb; d.vit = cf.vit; d.ri = cf.ri; d.qb_ld = cf.qb_ld; d.ri_ld = cf.ri_ld;
    d.y = cf.yoff; d.z = cf.zoff;
  }
  // thin descriptors of a level are contiguous (the level's compressed fronts are)
  const int t0 = f->nfr + f->comp[C.c0].fi;
  CUDA_OK(cudaMemcpyAsync(f->d_fronts + t0, f->fronts.data() + t0, (size_t)nc * sizeof(Front), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(f->d_lr + C.c0, lr.data(), (size_t)nc * sizeof(LrDesc), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));  // `lr` goes out of scope
  const LrDesc* d_lr = f->d_lr + C.c0;
  {
    dim3 g(nruns, (max_m + 127) / 128);
    k_form_q<T><<<g, 128, 0, st>>>(d_runs, d_lr, (T*)f->pool, ws, f->d_cint, f->d_fronts);
    dim3 g2(nc, std::min((max_ncol + 7) / 8, 128));
    k_build_thin<T><<<g2, 256, 0, st>>>(d_lr, d_runs, f->d_fronts, (T*)f->pool, ws);
    CUDA_OK(cudaGetLastError());
    f->stats.launches_factor += 2;
  }
  if (any_hchild) {
    // the children's generator blocks (:129-137, kept in the children's HSS stores by hs_hss.cu) in front of the pivoted-QR
    // factors:  Qb = [blkdiag(Û22·B21)ₗ,ᵣ  Qx],  Rb = [blkdiag(V̂11ᴴ)ₗ,ᵣ ; Rx],  Qi = [blkdiag(Û11·B12)ₗ,ᵣ  Qx'],  Ri = [blkdiag(V̂22ᴴ)ₗ,ᵣ ; Rx']
    std::vector<CopyDesc> cp;
    auto add = [&](long long src, int lds, long long dst, int ldd, int rows, int cols, int mode) {
      if (rows <= 0 || cols <= 0) return;
      CopyDesc d{};
      d.src = src; d.lds = lds; d.dst = dst; d.ldd = ldd; d.rows = rows; d.cols = cols; d.gap_at = rows; d.gap_skip = 0; d.mode = mode;
      cp.push_back(d);
    };
    for (int c = C.c0; c < C.c1; ++c) {
      const CompFront& cf = f->comp[c];
      if (!cf.hchild) continue;
      const Front& th = f->fronts[f->nfr + cf.fi];
      const HssFront* Hc[2] = {&f->hss[f->comp[cf.cl].hss], &f->hss[f->comp[cf.cr].hss]};
      int io = 0, bo = 0, ao = 0, co = 0;   // offsets: int rows, bnd rows, columns of the L part, columns of the R part
      for (int s2 = 0; s2 < 2; ++s2) {
        const HssFront& H = *Hc[s2];
        const int n1 = H.n1, nbc = H.m - H.n1;
        add(H.sbase + H.tb, H.ld_tb, cf.qb + bo + (long long)ao * cf.qb_ld, cf.qb_ld, nbc, H.ra1, 0);                                   // Qb
        add(H.sbase + H.vha, H.ld_vha, th.off + (cf.ni + ao) + (long long)io * th.ld, th.ld, H.ra1, n1, 0);                              // Rb
        add(H.sbase + H.ta, H.ld_ta, th.off + io + (long long)(cf.ni + co) * th.ld, th.ld, n1, H.rb1, 0);                                // Qi
        add(H.sbase + H.vhb, H.ld_vhb, cf.ri + co + (long long)bo * cf.ri_ld, cf.ri_ld, H.rb1, nbc, 0);                                  // Ri
        add(H.sbase + H.vhb, H.ld_vhb, cf.vit + bo + (long long)co * cf.qb_ld, cf.qb_ld, H.rb1, nbc, 2);                                 // Riᵀ
        io += n1; bo += nbc; ao += H.ra1; co += H.rb1;
      }
    }
    hs_hss_copy(f, cp);
  }
}

template <typename T> void schur_impl(hs_fac* f, CompLevel& C) {
  cudaStream_t st = f->ctx->stream;
  const int nc = C.c1 - C.c0;
  if (nc == 0) return;
  const IdRun* d_runs = f->d_runs + 2 * C.c0;
  const LrDesc* d_lr = f->d_lr + C.c0;
  T* ws = (T*)f->d_cws;
  int max_r2 = 0, max_ni = 0, max_nb = 0;
  for (int c = C.c0; c < C.c1; ++c) {
    const CompFront& cf = f->comp[c];
    max_r2 = std::max(max_r2, cf.r2); max_ni = std::max(max_ni, cf.ni); max_nb = std::max(max_nb, cf.nb);
  }
  if (max_r2 == 0) return;  // R = 0: S = Abb
  // descriptors: back substitution updates per 64-block (descending), then Z = Abi·Y, then S −= Z·Ri
  const int nblk = (max_ni + 63) / 64;
  std::vector<GemmDesc> gd;
  std::vector<int> g0(nblk + 1, 0), gmaxM(nblk, 0);
  for (int b = nblk - 1; b >= 1; --b) {
    g0[b] = (int)gd.size();
    for (int c = C.c0; c < C.c1; ++c) {
      const CompFront& cf = f->comp[c];
      const int b0 = b * 64;
      if (b0 >= cf.ni || cf.r2 == 0) continue;
      const Front& th = f->fronts[f->nfr + cf.fi];
      const int ldy = even_up(cf.ni);
      const long long y = cf.yoff;
      GemmDesc d{};
      d.a = th.off + (long long)b0 * th.ld; d.lda = th.ld;      // U[0:b0, b0:b0+db]
      d.b = y + b0; d.ldb = ldy;                                // Y[b0:b0+db, :]
      d.c = y; d.ldc = ldy;                                     // Y[0:b0, :]
      d.M = b0; d.N = cf.r2; d.K = std::min(64, cf.ni - b0); d.sign = -1;
      gd.push_back(d);
      gmaxM[b] = std::max(gmaxM[b], b0);
    }
  }
  std::vector<int> gcount(nblk, 0);
  for (int b = nblk - 1; b >= 1; --b) gcount[b] = (b == 1 ? (int)gd.size() : g0[b - 1]) - g0[b];
  // (g0 was filled in descending b, so the items of block b end where those of block b−1 begin)
  const int gz = (int)gd.size();
  for (int c = C.c0; c < C.c1; ++c) {
    const CompFront& cf = f->comp[c];
    const Front& fd = f->fronts[cf.fi];
    GemmDesc d{};
    d.a = fd.off + cf.ni; d.lda = fd.ld;                                    // Abi
    d.b = cf.yoff; d.ldb = even_up(cf.ni);                                  // Y = Aii⁻¹·Qi
    d.c = cf.zoff; d.ldc = even_up(cf.nb);                                  // Z (zeroed with the side buffer)
    d.M = cf.nb; d.N = cf.r2; d.K = cf.ni; d.sign = +1;
    gd.push_back(d);
  }
  const int gs = (int)gd.size();
  for (int c = C.c0; c < C.c1; ++c) {
    const CompFront& cf = f->comp[c];
    if (cf.hss >= 0) continue;   // HSS Schur complement: the operator stays matrix-free (Abb, Z, Ri), see hs_hss.cu
    const Front& fd = f->fronts[cf.fi];
    GemmDesc d{};
    d.a = cf.zoff; d.lda = even_up(cf.nb);                                  // Z
    d.b = cf.ri; d.ldb = cf.ri_ld;                                          // Ri
    d.c = fd.off + (long long)cf.ni * fd.ld + cf.ni; d.ldc = fd.ld;         // S (dense slot)
    d.M = cf.nb; d.N = cf.nb; d.K = cf.r2; d.sign = -1;
    gd.push_back(d);
  }
  if (gd.size() > f->gd_cap) {
    cudaFree(f->d_gd);
    f->d_gd = nullptr; f->gd_cap = 0;
    CUDA_OK(cudaMalloc(&f->d_gd, gd.size() * sizeof(GemmDesc)));
    f->gd_cap = gd.size();
  }
  CUDA_OK(cudaMemcpyAsync(f->d_gd, gd.data(), gd.size() * sizeof(GemmDesc), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));
  const GemmDesc* dg = (const GemmDesc*)f->d_gd;
  {
    dim3 g(nc, std::min((max_r2 + 7) / 8, 64));
    k_copy_y<T><<<g, 256, 0, st>>>(d_lr, d_runs, f->d_fronts, (T*)f->pool, ws);
    ++f->stats.launches_factor;
  }
  for (int b = nblk - 1; b >= 0; --b) {
    dim3 g(nc, std::min((max_r2 + 3) / 4, 64));
    k_diag_apply<T><<<g, 128, 0, st>>>(d_lr, d_runs, f->d_fronts, (T*)f->pool, ws, b);
    ++f->stats.launches_factor;
    if (b >= 1) gen_gemm<T>(f, dg + g0[b], gcount[b], gmaxM[b], max_r2);
  }
  CUDA_OK(cudaGetLastError());
  gen_gemm<T>(f, dg + gz, nc, max_nb, max_r2);
  gen_gemm<T>(f, dg + gs, (int)gd.size() - gs, max_nb, max_nb);
}

template <typename T> void solve_impl_c(hs_fac* f, const CompLevel& C, int64_t nrhs, T* x, bool fwd) {
  cudaStream_t st = f->ctx->stream;
  const int nc = C.c1 - C.c0;
  if (nc == 0) return;
  int max_nb = 0, max_r = 0, max_r1 = 0;
  for (int c = C.c0; c < C.c1; ++c) {
    max_nb = std::max(max_nb, f->comp[c].nb); max_r = std::max(max_r, f->comp[c].r); max_r1 = std::max(max_r1, f->comp[c].r1);
  }
  const LrDesc* d_lr = f->d_lr + C.c0;
  if (fwd) {
    if (max_r1 == 0) return;
    dim3 g(nc, (max_nb + 255) / 256, (unsigned)nrhs);
    k_lr_fwd<T><<<g, 256, (size_t)max_r1 * sizeof(T), st>>>(d_lr, f->d_fronts, (const T*)f->pool, f->d_gidx, x, f->xld, f->n);
  } else {
    if (max_r == 0) return;
    dim3 g(nc, std::min((max_r + 7) / 8, 64), (unsigned)nrhs);
    k_lr_bwd<T><<<g, 256, 0, st>>>(d_lr, f->d_fronts, (const T*)f->pool, f->d_gidx, x, f->xld, f->n);
  }
  CUDA_OK(cudaGetLastError());
  ++f->stats.launches_solve;
}

}  // namespace

void hs_comp_setup() {
  CUDA_OK(cudaFuncSetAttribute(k_gemm<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<double>()));
  CUDA_OK(cudaFuncSetAttribute(k_gemm<cplx, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<cplx>()));
  CUDA_OK(cudaFuncSetAttribute(k_id_step<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_id_step<cplx, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_lr_fwd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_lr_fwd<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
}

// workspaces and the static run descriptors; `comp` / `clevels` were filled by build_plan
void hs_comp_plan(hs_fac* f) {
  if (f->comp.empty()) return;
  cudaStream_t st = f->ctx->stream;
  f->runs.assign(2 * f->comp.size(), IdRun{});
  long long max_ws = 0, max_st = 0;
  int max_int = 0;
  for (CompLevel& C : f->clevels) {
    long long ws = 0, stt = 0;
    int ip = 1 + 2 * (C.c1 - C.c0);
    for (int c = C.c0; c < C.c1; ++c) {
      const CompFront& cf = f->comp[c];
      const Front& fd = f->fronts[cf.fi];
      const int ldr = even_up(cf.rcap);
      for (int which = 0; which < 2; ++which) {
        IdRun& R = f->runs[2 * c + which];
        if (!which) { R.moff = fd.off + cf.ni; R.m = cf.nb; R.ncol = cf.ni; }                      // Abi
        else { R.moff = fd.off + (long long)cf.ni * fd.ld; R.m = cf.ni; R.ncol = cf.nb; }          // Aib
        R.ld = fd.ld; R.rcap = cf.rcap; R.ldr = ldr;
        R.rws = ws; ws += up32((long long)(ldr + 2) * (R.ncol + 2));
        R.st = stt; stt += 2ll * R.ncol + 2;
        R.ip = ip; ip += cf.rcap;
        R.slot = 2 * (c - C.c0) + which;
        R.pad = 0;
      }
    }
    max_ws = std::max(max_ws, ws); max_st = std::max(max_st, stt); max_int = std::max(max_int, ip);
    // HSS-children fronts: the pivoted QRs run on copies of the sparse coupling blocks alone (factorization.jl:186-190,199-203)
    long long xo = 0;
    for (int c = C.c0; c < C.c1; ++c) {
      CompFront& cf = f->comp[c];
      if (!cf.hchild) continue;
      cf.ldx1 = even_up(cf.nb); cf.x1 = xo; xo += up32((long long)cf.ldx1 * cf.ni);
      cf.ldx2 = even_up(cf.ni); cf.x2 = xo; xo += up32((long long)cf.ldx2 * cf.nb);
    }
    if (xo > 0) {
      C.xws_bytes = (size_t)xo * f->esz;
      CUDA_OK(cudaMalloc(&C.xws, C.xws_bytes));
      const long long xb = (long long)(((char*)C.xws - (char*)f->pool) / (long long)f->esz);
      for (int c = C.c0; c < C.c1; ++c) {
        CompFront& cf = f->comp[c];
        if (!cf.hchild) continue;
        cf.x1 += xb; cf.x2 += xb;
        IdRun& Ra = f->runs[2 * c];
        IdRun& Rb = f->runs[2 * c + 1];
        Ra.moff = cf.x1; Ra.ld = cf.ldx1;
        Rb.moff = cf.x2; Rb.ld = cf.ldx2;
        Rb.pad = 1;   // pqrfact(…; atol = 0.5·atol, rtol = 0.5·rtol) on top of the halved tolerances (:202)
      }
    }
  }
  CUDA_OK(cudaMalloc(&f->d_cws, (size_t)std::max<long long>(max_ws, 32) * f->esz));
  CUDA_OK(cudaMalloc((void**)&f->d_cstate, (size_t)std::max<long long>(max_st, 1) * sizeof(double)));
  CUDA_OK(cudaMalloc((void**)&f->d_cint, (size_t)std::max(max_int, 1) * sizeof(int)));
  CUDA_OK(cudaMalloc((void**)&f->d_runs, f->runs.size() * sizeof(IdRun)));
  CUDA_OK(cudaMalloc((void**)&f->d_lr, f->comp.size() * sizeof(LrDesc)));
  CUDA_OK(cudaMemcpyAsync(f->d_runs, f->runs.data(), f->runs.size() * sizeof(IdRun), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));
}

void hs_gen_gemm(hs_fac* f, const GemmDesc* d_items, int nitems, int maxM, int maxN) {
  if (f->dtype == HS_F64) gen_gemm<double>(f, d_items, nitems, maxM, maxN); else gen_gemm<cplx>(f, d_items, nitems, maxM, maxN);
}

void hs_comp_prepare(hs_fac* f, CompLevel& C) {
  if (f->dtype == HS_F64) prepare_impl<double>(f, C); else prepare_impl<cplx>(f, C);
}
void hs_comp_schur(hs_fac* f, CompLevel& C) {
  if (f->dtype == HS_F64) schur_impl<double>(f, C); else schur_impl<cplx>(f, C);
}
void hs_comp_solve(hs_fac* f, const CompLevel& C, int64_t nrhs, void* x, bool fwd) {
  if (f->dtype == HS_F64) solve_impl_c<double>(f, C, nrhs, (double*)x, fwd); else solve_impl_c<cplx>(f, C, nrhs, (cplx*)x, fwd);
}
