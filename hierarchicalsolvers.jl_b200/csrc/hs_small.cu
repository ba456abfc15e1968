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

// ------------------------------------------------------------------------------------------------
// real fronts with n ≤ 64 (the two bottom levels of a 2D tree: 3/4 of all fronts): ONE THREAD PER FRONT ROW, the row in
// 64 registers, two warps per front.  Per pivot column: each warp finds its candidate with three REDUX operations, the
// candidate's owner publishes the rest of its row, one barrier, both warps pick the winner from shared memory and
// eliminate from registers — the pivot row is a broadcast LDS.128 stream, and there are no per-element predicates, no
// select chains over register tiles and no second barrier (the TR×TC tile kernel above spends 85% of its issue slots on
// those: 43 000 warp instructions per leaf front, this one ≈ 14 000).  Register indices must be static, so the column
// loop is unrolled over blocks of 8 columns only: after a block the row is ROTATED by 8 registers (finished columns go
// to a shared-memory stash until the row's final position is known), which keeps the loop body at ~40 KB of code.
// Measured variants: 6 CTAs per SM (168 registers, spills) 5.3 ms for the two bottom levels of the 2048² workload, 4 CTAs
// without spills 4.5 ms, blocks of 4 columns 6.4 ms, pivot search of column j+1 issued inside step j 5.9 ms; the TR×TC
// tile kernel 6.8 ms.
// Rows are not moved until the write-back; rperm records the pivot order (the LAPACK interchange sequence is derived
// from it on the host when a caller asks for it).
// ------------------------------------------------------------------------------------------------
constexpr int RW_NC = 64, RW_JB = 8;
__global__ void __launch_bounds__(RW_NC, 4) k_front_rows(const Front* __restrict__ fronts, double* __restrict__ pool,
                                                         int* __restrict__ rperm, int f0, int* __restrict__ info) {
  constexpr int NC = RW_NC, JB = RW_JB;
  const int fi = f0 + blockIdx.x;
  const Front fr = fronts[fi];
  const int n = fr.n, ni = fr.ni;
  if (n == 0) return;
  double* F = pool + fr.off;
  const long long ld = fr.ld;
  const int r = threadIdx.x, lane = r & 31, warp = r >> 5;
  __shared__ __align__(16) double s_u[2][2][NC];     // [column parity][warp] rest of the warp's candidate row
  __shared__ unsigned long long s_key[2][2];         // candidate |.| as an ordered key (0: the warp has none)
  __shared__ int s_row[2][2];
  __shared__ double s_stash[(NC - JB) * NC];         // finished columns of every row, column-major

  double a[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) a[c] = (r < n && c < n) ? F[(long long)c * ld + r] : 0.0;
  bool cand = r < ni;   // may still be chosen as a pivot
  bool live = true;     // has not been a pivot yet: takes part in the elimination
  int mypos = r;        // final row position
  int base = 0;         // a[k] holds column base + k
  for (;;) {
#pragma unroll
    for (int jt = 0; jt < JB; ++jt) {
      const int j = base + jt;
      if (j >= ni) break;
      const int par = jt & 1;
      double v = cand ? fabs(a[jt]) : -1.0;
      int idx = r;
      warp_argmax(v, idx);
      if (r == idx && v >= 0.0) {
        // columns jt.. of the candidate row, 16 bytes at a time from the even column at or below jt
#pragma unroll
        for (int k = jt & ~1; k < NC; k += 2) *reinterpret_cast<double2*>(&s_u[par][warp][k]) = make_double2(a[k], a[k + 1]);
      }
      if (lane == 0) {
        s_key[par][warp] = v < 0.0 ? 0ull : (unsigned long long)__double_as_longlong(v) + 1ull;
        s_row[par][warp] = idx;
      }
      __syncthreads();
      const unsigned long long k0 = s_key[par][0], k1 = s_key[par][1];
      const int w = k1 > k0 ? 1 : 0;                 // ties go to the smaller row, which lives in warp 0
      const int p = s_row[par][w];
      const bool ok = (w ? k1 : k0) >= 2ull;         // key 1 is |.| = 0: exactly singular column, formal pivot only
      if (r == p) { cand = false; live = false; mypos = j; }
      if (!ok) {
        if (r == 0 && atomicCAS(&info[0], 0, 1) == 0) { info[1] = fi; info[2] = j; }
      } else if (live) {
        const double* u = s_u[par][w];
        const double l = a[jt] * hs_recip_pivot(u[jt]);
        a[jt] = l;
        if ((jt & 1) == 0) a[jt + 1] = fma(-l, u[jt + 1], a[jt + 1]);
#pragma unroll
        for (int k = (jt + 2) & ~1; k < NC; k += 2) {
          const double2 uu = *reinterpret_cast<const double2*>(&u[k]);
          a[k] = fma(-l, uu.x, a[k]);
          a[k + 1] = fma(-l, uu.y, a[k + 1]);
        }
      }
    }
    if (base + JB >= ni) break;
    // rotate: the block's columns are final for this row
#pragma unroll
    for (int c = 0; c < JB; ++c) s_stash[(base + c) * NC + r] = a[c];
#pragma unroll
    for (int k = 0; k < NC - JB; ++k) a[k] = a[k + JB];
#pragma unroll
    for (int k = NC - JB; k < NC; ++k) a[k] = 0.0;
    base += JB;
  }
  if (r >= n) return;
  if (r < ni) rperm[fr.ioff + mypos] = r;
  double* dst = F + mypos;
  for (int c = 0; c < base; ++c) dst[(long long)c * ld] = s_stash[c * NC + r];
#pragma unroll
  for (int k = 0; k < NC; ++k)
    if (base + k < n) dst[(long long)(base + k) * ld] = a[k];
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
    static const bool rows = !(getenv("HS_SMALL_ROWS") && atoi(getenv("HS_SMALL_ROWS")) == 0);
    if (n <= RW_NC && L.max_ni <= RW_NC && rows) {
      k_front_rows<<<nf, RW_NC, 0, f->ctx->stream>>>(f->d_fronts, (double*)f->pool, f->d_rperm, L.f0, f->d_info);
      CUDA_OK(cudaGetLastError());
    } else if (n <= 80) { if (mb >= 4) launch<double, 16, 8, 5, 10, 4>(f, L.f0, nf); else if (mb == 3) launch<double, 16, 8, 5, 10, 3>(f, L.f0, nf); else launch<double, 16, 8, 5, 10, 1>(f, L.f0, nf); }
    else if (n <= 96) launch<double, 16, 16, 6, 6>(f, L.f0, nf);
    else launch<double, 16, 16, 8, 8>(f, L.f0, nf);
  } else {
    if (n <= 80) launch<cplx, 16, 16, 5, 5>(f, L.f0, nf);
    else launch<cplx, 16, 16, 6, 6>(f, L.f0, nf);
  }
  f->stats.launches_factor += 1;
}
