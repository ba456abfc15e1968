// Fused factorization of small fronts: one CTA per front, the whole front in REGISTERS.
//
// Covers what the reference does per leaf / small branch in _factor_leaf and _factor_branch
// (factorization.jl:30-42, 62-75): LU of the pivot block with partial pivoting (`D \ …`), L = A_bi·D⁻¹, R = D⁻¹·A_ib
// and the Schur complement S = A_bb − A_bi·R — here as one right-looking partial LU of the assembled front, read
// once from HBM and written once.  Tens of thousands of such fronts sit at the bottom of the elimination tree, so
// the kernel is organised for many resident CTAs per SM rather than for a single fast front:
//
//   thread (tr, tc) of a TR×TC grid owns rows tr + TR·i, columns tc + TC·k of the front (RPT×CPT registers);
//   per pivot column: the column's owners publish it to shared memory → every warp finds the pivot redundantly
//   (no broadcast) → the pivot row's owners publish it → rank-1 update from registers.  Two barriers per column.
//   Pivoting is implicit (rows are never moved until the write-back), which yields the same L, U and row order as
//   LAPACK's explicit interchanges.
#include <cuda_runtime.h>

#include <cstdlib>

#include "hs_fac.cuh"

namespace {

template <typename T, int TR, int TC, int RPT, int CPT, int MINB>
__global__ void __launch_bounds__(TR* TC, MINB) k_front_small(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                         int* __restrict__ ipiv, int* __restrict__ rperm, int f0,
                                                         int* __restrict__ info) {
  constexpr int NT = TR * TC;
  constexpr int NMAX = TR * RPT;  // rows covered
  constexpr int CMAX = TC * CPT;  // columns covered
  const int fi = f0 + blockIdx.x;
  const Front fr = fronts[fi];
  const int n = fr.n, ni = fr.ni;
  if (n == 0) return;
  T* F = pool + fr.off;
  const long long ld = fr.ld;
  const int tid = threadIdx.x, lane = tid & 31;
  const int tr = tid % TR, tc = tid / TR;

  __shared__ T s_col[2][NMAX];
  __shared__ double s_abs[2][NMAX];
  __shared__ T s_row[2][CMAX];
  __shared__ int s_piv[NMAX];   // s_piv[k] = physical row chosen as k-th pivot
  __shared__ int s_pos[NMAX];   // final position of every physical row
  __shared__ int s_what[NMAX], s_where[NMAX];

  T a[RPT][CPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = tr + TR * i;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tc + TC * k;
      a[i][k] = (r < n && c < n) ? F[(long long)c * ld + r] : hs_zero<T>();
    }
  }
  unsigned mydone = 0;  // bit i: my row i has already been a pivot
  bool singular = false;

#pragma unroll
  for (int kj = 0; kj < CPT; ++kj) {
    for (int jt = 0; jt < TC; ++jt) {
      const int j = kj * TC + jt;
      if (j >= ni) break;
      const int par = j & 1;
      // (a) owners of column j publish it (|.| = -1 for rows that are not pivot candidates)
      if (tc == jt) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = tr + TR * i;
          s_col[par][r] = a[i][kj];
          s_abs[par][r] = (r < ni && !((mydone >> i) & 1u)) ? hs_abs1(a[i][kj]) : -1.0;
        }
      }
      __syncthreads();
      // (b) every warp searches the pivot (largest |.|, smallest row on ties)
      double best = -1.0;
      int p = 0x7fffffff;
      for (int r = lane; r < ni; r += 32) {
        const double v = s_abs[par][r];
        if (v > best) { best = v; p = r; }
      }
      warp_argmax(best, p);
      const bool ok = best > 0.0;
      if (!ok) {  // exactly singular column: take the first remaining candidate as a formal pivot, no elimination
        singular = true;
        p = 0x7fffffff;
        for (int r = lane; r < ni; r += 32)
          if (s_abs[par][r] >= 0.0) { p = r; break; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p = min(p, __shfl_xor_sync(0xffffffffu, p, o));
        if (tid == 0 && atomicCAS(&info[0], 0, 1) == 0) { info[1] = fi; info[2] = j; }
      }
      // (c) owners of the pivot row publish it
      const int ip = p / TR;
      if (tr == p % TR) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          if (i == ip) {
#pragma unroll
            for (int k = kj; k < CPT; ++k) s_row[par][tc + TC * k] = a[i][k];
            mydone |= 1u << i;
          }
        }
      }
      if (tid == 0) s_piv[j] = p;
      __syncthreads();
      // (d) elimination from registers
      if (ok) {
        const T inv = hs_recip_pivot(s_col[par][p]);
        T l[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = tr + TR * i;
          const bool act = r < n && !((mydone >> i) & 1u);
          l[i] = act ? hs_mul(s_col[par][r], inv) : hs_zero<T>();
          if (act && tc == jt) a[i][kj] = l[i];
        }
#pragma unroll
        for (int k = kj; k < CPT; ++k) {
          if (k == kj && tc <= jt) continue;
          const T u = s_row[par][tc + TC * k];
#pragma unroll
          for (int i = 0; i < RPT; ++i) a[i][k] = hs_fnma(a[i][k], l[i], u);
        }
      }
    }
  }
  (void)singular;
  __syncthreads();
  // final positions: k-th pivot → row k; boundary rows stay.  LAPACK-style ipiv from the pivot order.
  for (int r = tid; r < n; r += NT) { s_pos[r] = r; s_what[r] = r; s_where[r] = r; }
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < ni; ++k) {
      const int pr = s_piv[k];
      s_pos[pr] = k;
      const int q = s_where[pr];          // current position of the row that becomes pivot k
      ipiv[fr.ioff + k] = q;
      const int other = s_what[k];        // row currently sitting at position k
      s_what[k] = pr; s_what[q] = other;
      s_where[pr] = k; s_where[other] = q;
      rperm[fr.ioff + k] = pr;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int r = tr + TR * i;
    if (r >= n) continue;
    const int pr = s_pos[r];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tc + TC * k;
      if (c < n) F[(long long)c * ld + pr] = a[i][k];
    }
  }
}

template <typename T, int TR, int TC, int RPT, int CPT, int MINB = 1> void launch(hs_fac* f, int f0, int nf) {
  k_front_small<T, TR, TC, RPT, CPT, MINB><<<nf, TR * TC, 0, f->ctx->stream>>>(f->d_fronts, (T*)f->pool, f->d_ipiv, f->d_rperm, f0, f->d_info);
  CUDA_OK(cudaGetLastError());
}

}  // namespace

// largest front (rows) the fused register kernel takes for this scalar type; 0 disables it
int hs_small_max_n(hs_dtype dt) {
  if (getenv("HS_NO_SMALL")) return 0;
  return dt == HS_F64 ? 128 : 96;
}

// whole level in one launch; the caller guarantees max_n ≤ hs_small_max_n()
void hs_small_factor(hs_fac* f, const Level& L) {
  const int nf = L.f1 - L.f0, n = L.max_n;
  if (f->dtype == HS_F64) {
    static const int mb = getenv("HS_SMALL_MINB") ? atoi(getenv("HS_SMALL_MINB")) : 3;
    if (n <= 80) { if (mb >= 4) launch<double, 16, 8, 5, 10, 4>(f, L.f0, nf); else if (mb == 3) launch<double, 16, 8, 5, 10, 3>(f, L.f0, nf); else launch<double, 16, 8, 5, 10, 1>(f, L.f0, nf); }
    else if (n <= 96) launch<double, 16, 16, 6, 6>(f, L.f0, nf);
    else launch<double, 16, 16, 8, 8>(f, L.f0, nf);
  } else {
    if (n <= 80) launch<cplx, 16, 16, 5, 5>(f, L.f0, nf);
    else launch<cplx, 16, 16, 6, 6>(f, L.f0, nf);
  }
  f->stats.launches_factor += 1;
}
