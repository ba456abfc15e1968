// Hand-written sm_100a kernels of the multifrontal factor + tree solve.
//
//   assembly   k_fill_owner / k_scatter_A / k_extend_add      (factorization.jl:33-40,115-123)  HBM-bound
//   LU panel   k_panel   — register-resident panel, one thread-block cluster per front,
//                          pivot search over distributed shared memory                (blockmatrix.jl:118, `\`)
//   row ops    k_swap_trsm / k_swap_trsm_warp / k_laswp — row interchanges + unit-lower triangular solve of the U row
//                          panel (thread per column for the wide launches, warp per column on the panel chain)
//   update     k_gemm    — FP64 / complex-FP64 tensor-core (DMMA m8n8k4) Schur update C -= A·B
//                                                                                  (factorization.jl:40,72)
//   elsewhere  hs_panel.cuh: k_panel, k_trsm_rows; hs_small.cu: k_front_small, k_front_rows (whole small fronts);
//              hs_solve.cu: k_trtri_diag, k_sv_small_*, k_sv_tri_*, k_gemv_rect  (factornode.jl:77-99)  HBM-bound
#pragma once

#include <cooperative_groups.h>

#include "hs_types.cuh"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------
// assembly
// ------------------------------------------------------------------------------------------------

// own[g] = front that holds DOF g on this level, pos[g] = its row inside that front
static __global__ void k_fill_owner(const Front* __restrict__ fronts, const int* __restrict__ gidx, int* __restrict__ own,
                             int* __restrict__ pos, int f0) {
  const int fi = f0 + blockIdx.x;
  const Front fr = fronts[fi];
  const int k = blockIdx.y * blockDim.x + threadIdx.x;
  if (k >= fr.n) return;
  const int g = gidx[fr.ioff + k];
  own[g] = fi;
  pos[g] = k;
}

// Gather of the original sparse A into the fronts of one level:
//   leaf   F = A[[int;bnd],[int;bnd]]                                   (factorization.jl:33-40)
//   branch only the couplings between the two children, A[l.bnd, r.bnd] and A[r.bnd, l.bnd]
//          (the off-diagonal blocks of _assemble_blocks, factorization.jl:118-121)
// one thread per front column; A is CSC so a column of A is contiguous.
template <typename T>
__global__ void k_scatter_A(const Front* __restrict__ fronts, T* __restrict__ pool, const int* __restrict__ gidx,
                            const int* __restrict__ own, const int* __restrict__ pos,
                            const long long* __restrict__ colptr, const long long* __restrict__ rowval,
                            const T* __restrict__ nzval, int f0) {
  const int fi = f0 + blockIdx.x;
  const Front fr = fronts[fi];
  const int k = blockIdx.y * blockDim.x + threadIdx.x;
  if (k >= fr.n || (fr.flags & 4)) return;  // bit2: external leaf, its front is an imported Schur block
  const int g = gidx[fr.ioff + k];
  T* col = pool + fr.off + (long long)k * fr.ld;
  const bool leaf = fr.ni_l < 0;
  // side of a front row: 0 = came from the left child, 1 = from the right child
  auto side = [&](int r) { return r < fr.ni ? (r >= fr.ni_l) : (r - fr.ni >= fr.nb_l); };
  const int sk = leaf ? 0 : side(k);
  for (long long p = colptr[g]; p < colptr[g + 1]; ++p) {
    const int i = (int)rowval[p];
    if (own[i] != fi) continue;
    const int r = pos[i];
    if (leaf || side(r) != sk) col[r] = nzval[p];
  }
}

// "Extend-add": copy a child's Schur complement into its parent's front.  In this formulation child
// boundaries are disjoint, so nothing is accumulated — it is a collision-free permuted block copy
// (factorization.jl:118-121 diagonal blocks; the S[perm,perm] of :41,:74 is folded into the map).
// cmap[a] = row of the parent front that receives the child's boundary row a (-1: dropped).
template <typename T, int CB, int LW>
__global__ void __launch_bounds__(256) k_extend_add(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                     const int* __restrict__ cmap, int c0, int cols_per_cta) {
  const int ci = c0 + blockIdx.x;
  const Front ch = fronts[ci];
  if (ch.parent < 0) return;
  const int nb = ch.n - ch.ni;
  const int b0 = blockIdx.y * cols_per_cta;
  if (b0 >= nb) return;
  const int b1 = min(nb, b0 + cols_per_cta);
  const Front pa = fronts[ch.parent];
  const int* map = cmap + ch.ioff + ch.ni;
  const T* S = pool + ch.off + (long long)ch.ni * ch.ld + ch.ni;
  T* P = pool + pa.off;
  // LW lanes walk the rows of one column of the child's Schur block (no per-element div/mod), so a warp covers 32/LW
  // columns at a time (LW = 8 / 16 for the small fronts of the deep levels, 32 above).  CB row chunks are loaded before
  // the first store (source and destination live in the same pool).  f64: two rows per lane through one 16-byte load
  // when the column start is 16-byte aligned (ni even; the leading dimension always is even).
  constexpr int CPW = 32 / LW;                  // columns per warp and pass
  constexpr int RPL = sizeof(T) == 8 ? 2 : 1;   // rows per lane and chunk
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lr = lane % LW, lc = lane / LW;
  const bool vec = RPL == 2 && ((ch.off + (long long)ch.ni * ch.ld + ch.ni) & 1) == 0;
  for (int b = b0 + warp * CPW + lc; b < b1; b += 8 * CPW) {
    const int pc = map[b];
    if (pc < 0) continue;
    const T* src = S + (long long)b * ch.ld;
    T* dst = P + (long long)pc * pa.ld;
    for (int a0 = 0; a0 < nb; a0 += LW * RPL * CB) {
      T v[CB][RPL];
      int d[CB][RPL];
#pragma unroll
      for (int q = 0; q < CB; ++q) {
        const int a = a0 + (q * LW + lr) * RPL;
        if constexpr (RPL == 2) {
          if (vec && a + 1 < nb) {
            const double2 t = *reinterpret_cast<const double2*>(src + a);
            v[q][0] = t.x; v[q][1] = t.y;
            d[q][0] = map[a]; d[q][1] = map[a + 1];
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const bool ok = a + h < nb;
              v[q][h] = ok ? src[a + h] : hs_zero<T>();
              d[q][h] = ok ? map[a + h] : -1;
            }
          }
        } else {
          const bool ok = a < nb;
          v[q][0] = ok ? src[a] : hs_zero<T>();
          d[q][0] = ok ? map[a] : -1;
        }
      }
#pragma unroll
      for (int q = 0; q < CB; ++q)
#pragma unroll
        for (int h = 0; h < RPL; ++h)
          if (d[q][h] >= 0) dst[d[q][h]] = v[q][h];
    }
  }
}

// rperm[k] = local row of the ORIGINAL pivot block that ends at row k after all interchanges (so P·x is a gather).
// The identity before the first panel; every k_panel applies its row moves to it.
static __global__ void k_rperm_init(const Front* __restrict__ fronts, int* __restrict__ rperm, int f0) {
  const Front fr = fronts[f0 + blockIdx.x];
  const int k = blockIdx.y * blockDim.x + threadIdx.x;
  if (k < fr.ni) rperm[fr.ioff + k] = k;
}

// ------------------------------------------------------------------------------------------------
// row interchanges + U row panel:  U(j0:j0+wc, c) = L11⁻¹·(P·A)(j0:j0+wc, c);  one thread per front column.
// Two-level blocked LU (outer block [J0, BE), BE = min(J0+NB, ni)):
//   cmode 0 (while the block is being factored): the interchanges of panel j0 go to the block's own columns
//            [J0, j0) ∪ [j0+wc, BE); the columns right of the panel inside the block also get the triangular solve.
//   cmode 1 (after the block): columns [BE, n) get the triangular solve of sub-block j0 — their interchanges were
//            applied for the whole block at once by k_laswp (they must all precede the first solve).
// ------------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(128) k_swap_trsm(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                    const int* __restrict__ ipiv, int f0, int J0, int j0, int NB,
                                                    int cmode) {
  const int fi = f0 + blockIdx.x;
  const Front fr = fronts[fi];
  if (fr.ni <= j0) return;
  const int wc = min(W, fr.ni - j0);
  const int BE = min(J0 + NB, fr.ni);
  const int nleft = cmode == 0 ? j0 - J0 : 0;
  const int ncols = cmode == 0 ? nleft + (BE - j0 - wc) : fr.n - BE;
  if ((int)(blockIdx.y * blockDim.x) >= ncols) return;
  T* F = pool + fr.off;
  __shared__ T sL[W * W];
  __shared__ int spiv[W];
  {
    // all loads of the W×W block are issued before the first store (a rolled loop pays one L2 latency per element:
    // ~20 µs per launch, most of what these launches cost in round 1)
    constexpr int NT = 128, NE = (W * W + NT - 1) / NT;
    T tmp[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = threadIdx.x + q * NT;
      const int i = e % W, k = e / W;
      tmp[q] = (e < W * W && i < wc && k < wc && i > k) ? F[(long long)(j0 + k) * fr.ld + (j0 + i)] : hs_zero<T>();
    }
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = threadIdx.x + q * NT;
      if (e < W * W) sL[e] = tmp[q];
    }
  }
  if (threadIdx.x < W) spiv[threadIdx.x] = threadIdx.x < wc ? ipiv[fr.ioff + j0 + threadIdx.x] : 0;
  __syncthreads();
  const int cc = blockIdx.y * blockDim.x + threadIdx.x;
  if (cc >= ncols) return;
  int c;
  bool solve = true;
  if (cmode == 0) {
    solve = cc >= nleft;
    c = solve ? j0 + wc + (cc - nleft) : J0 + cc;
  } else {
    c = BE + cc;
  }
  T* col = F + (long long)c * fr.ld;
  if (cmode == 0) {
    for (int j = 0; j < wc; ++j) {
      const int p = spiv[j];
      if (p != j0 + j) { const T t = col[j0 + j]; col[j0 + j] = col[p]; col[p] = t; }
    }
  }
  if (!solve) return;
  T x[W];
#pragma unroll
  for (int i = 0; i < W; ++i) x[i] = i < wc ? col[j0 + i] : hs_zero<T>();
#pragma unroll
  for (int k = 0; k < W - 1; ++k) {
    if (k < wc - 1) {
      const T xk = x[k];
#pragma unroll
      for (int i = k + 1; i < W; ++i) x[i] = hs_fnma(x[i], sL[k * W + i], xk);
    }
  }
#pragma unroll
  for (int i = 0; i < W; ++i)
    if (i < wc) col[j0 + i] = x[i];
}

// The same operation with ONE WARP PER COLUMN, for the launches on the panel chain of the upper levels (a few hundred
// columns in all): lane i keeps x[i] (and x[32+i]), step k broadcasts x[k] with a shuffle.  400 warp instructions per column
// in a rolled loop instead of 4 000 straight-line ones per 32 columns — three times the work of the thread-per-column
// kernel, but a small launch finishes in ~1/2 of the time because it is not bound by fetching 64 KB of cold code.
template <typename T, int W>
__global__ void __launch_bounds__(256) k_swap_trsm_warp(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                         const int* __restrict__ ipiv, int f0, int J0, int j0, int NB,
                                                         int cmode) {
  static_assert(W == 32 || W == 64, "one or two rows per lane");
  constexpr int NT = 256, WPC = NT / 32;
  const int fi = f0 + blockIdx.x;
  const Front fr = fronts[fi];
  if (fr.ni <= j0) return;
  const int wc = min(W, fr.ni - j0);
  const int BE = min(J0 + NB, fr.ni);
  const int nleft = cmode == 0 ? j0 - J0 : 0;
  const int ncols = cmode == 0 ? nleft + (BE - j0 - wc) : fr.n - BE;
  if ((int)(blockIdx.y * WPC) >= ncols) return;
  T* F = pool + fr.off;
  __shared__ T sL[W * W];
  __shared__ int spiv[W];
  {
    constexpr int NE = (W * W + NT - 1) / NT;
    T tmp[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = threadIdx.x + q * NT;
      const int i = e % W, k = e / W;
      tmp[q] = (e < W * W && i < wc && k < wc && i > k) ? F[(long long)(j0 + k) * fr.ld + (j0 + i)] : hs_zero<T>();
    }
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      const int e = threadIdx.x + q * NT;
      if (e < W * W) sL[e] = tmp[q];
    }
  }
  if (threadIdx.x < W) spiv[threadIdx.x] = threadIdx.x < wc ? ipiv[fr.ioff + j0 + threadIdx.x] : j0 + threadIdx.x;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cc = blockIdx.y * WPC + warp;
  if (cc >= ncols) return;
  int c;
  bool solve = true;
  if (cmode == 0) {
    solve = cc >= nleft;
    c = solve ? j0 + wc + (cc - nleft) : J0 + cc;
  } else {
    c = BE + cc;
  }
  T* col = F + (long long)c * fr.ld;
  if (cmode == 0) {
    // interchanges are sequential; skipped altogether when the panel moved no row
    bool moved = spiv[lane] != j0 + lane;
    if (W == 64) moved = moved || spiv[32 + lane] != j0 + 32 + lane;
    if (__any_sync(0xffffffffu, moved)) {
      if (lane == 0) {
        for (int j = 0; j < wc; ++j) {
          const int p = spiv[j];
          if (p != j0 + j) { const T t = col[j0 + j]; col[j0 + j] = col[p]; col[p] = t; }
        }
      }
      __syncwarp();
    }
  }
  if (!solve) return;
  T x0 = lane < wc ? col[j0 + lane] : hs_zero<T>();
  T x1 = hs_zero<T>();
  if (W == 64 && 32 + lane < wc) x1 = col[j0 + 32 + lane];
#pragma unroll 4
  for (int k = 0; k < 32; ++k) {
    const T xk = hs_shfl(x0, k);
    if (lane > k) x0 = hs_fnma(x0, sL[k * W + lane], xk);
    if (W == 64) x1 = hs_fnma(x1, sL[k * W + 32 + lane], xk);
  }
  if (W == 64) {
#pragma unroll 4
    for (int k = 32; k < 64; ++k) {
      const T xk = hs_shfl(x1, k - 32);
      if (lane > k - 32) x1 = hs_fnma(x1, sL[k * W + 32 + lane], xk);
    }
  }
  if (lane < wc) col[j0 + lane] = x0;
  if (W == 64 && 32 + lane < wc) col[j0 + 32 + lane] = x1;
}

// interchanges of a whole outer block [J0, BE) applied to the columns outside it, [0, J0) ∪ [BE, n)
template <typename T>
__global__ void __launch_bounds__(128) k_laswp(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                const int* __restrict__ ipiv, int f0, int J0, int NB) {
  const Front fr = fronts[f0 + blockIdx.x];
  if (fr.ni <= J0) return;
  const int BE = min(J0 + NB, fr.ni);
  const int ncols = J0 + fr.n - BE;
  if ((int)(blockIdx.y * blockDim.x) >= ncols) return;
  // the block's interchanges that really move a row, compacted in shared memory with one coalesced read (a thread
  // walking ipiv itself pays one L2 latency per pivot: ~17 µs per launch on the panel chain of the top fronts)
  __shared__ int s_j[128], s_p[128];
  __shared__ int s_n;
  const int* pv = ipiv + fr.ioff;
  const int cc = blockIdx.y * blockDim.x + threadIdx.x;
  const int c = cc < J0 ? cc : BE + (cc - J0);
  T* col = pool + fr.off + (long long)c * fr.ld;
  for (int b = J0; b < BE; b += 128) {   // 128 pivots per pass, kept in order
    const int j = b + threadIdx.x;
    const int p = j < BE ? pv[j] : j;
    const bool mv = p != j;
    const unsigned m = __ballot_sync(0xffffffffu, mv);
    __shared__ int s_wcnt[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_wcnt[w] = __popc(m);
    __syncthreads();
    int base = 0;
    for (int q = 0; q < w; ++q) base += s_wcnt[q];
    if (mv) { const int k = base + __popc(m & ((1u << lane) - 1u)); s_j[k] = j; s_p[k] = p; }
    if (threadIdx.x == 0) s_n = s_wcnt[0] + s_wcnt[1] + s_wcnt[2] + s_wcnt[3];
    __syncthreads();
    const int nm = s_n;
    if (cc < ncols)
      for (int k = 0; k < nm; ++k) { const int jj = s_j[k], pp = s_p[k]; const T t = col[jj]; col[jj] = col[pp]; col[pp] = t; }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Schur / trailing update on the FP64 tensor cores:  C -= A·B  with A = L panel, B = U row panel of the same front.
// DMMA m8n8k4 (the native FP64 MMA shape on sm_100a: `DMMA.8x8x4`); complex = 4 real DMMAs on interleaved operands.
// 128×64 (f64) / 64×64 (c64) CTA tile, 4 warps, two CTAs per SM, K in chunks staged by a 4-deep cp.async pipeline,
// the C tile read-modify-written through shared memory.  GEN = true (GemmDesc) turns the same kernel into a free-standing
// batched product C ∓= A·B for the compressed path (hs_compress.cu).
//
// Two-level blocking: inside an outer block [J0, BE) of pivot columns (BE = min(J0+NB, ni)) the inner panels of
// width W update only the L-shaped region they must (mode 0); the big trailing block [BE,n)² is updated once per
// outer block with K = BE−J0 (mode 1), which is where almost all flops go.
//   mode 0: K = [j0, j0+wc),  C = rows [j0+wc, n)  × cols [j0+wc, BE)   (inside the block, while it is factored)
//   mode 2: K = [j0, j0+wc),  C = rows [j0+wc, BE) × cols [BE, n)      (top strip of the outside columns, after it)
//   mode 1: K = [J0, BE),     C = [BE, n)²                             (the big update, K up to NB)
//   mode 5 / 6: mode 0 split by rows: [j0+wc, plim) — all the next panel needs — and [plim, n), plim = end of the rows
//               pivots are taken from; mode 6 runs on the below-rows stream together with k_trsm_rows
//   mode 3 / 4: the big update split for look-ahead: columns of the NEXT outer block [BE, BE2) / the rest [BE2, n),
//               BE2 = min(BE+NB, ni); mode 4 runs on a second stream while the next block's panels are factored
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T> struct GemmCfg;
template <> struct GemmCfg<double> {
  // 128×64 tile, 4 warps of 64×32, 164 registers × 128 threads.  Two pipeline stages keep the operand ring (54 KB) below
  // the staged C tile (66 KB), so THREE CTAs fit on an SM and cover each other's prologue/epilogue/barrier stalls; with
  // four stages (108 KB, two CTAs) the Schur updates of the 2048² workload took 39.5 instead of 35.4 ms.
  static constexpr int TM = 128, TN = 64, WM = 64, WN = 32, KC = 16, ST = 2, LDA = 132, LDB = 20;
};
template <> struct GemmCfg<cplx> {
  static constexpr int TM = 64, TN = 64, WM = 32, WN = 32, KC = 8, ST = 4, LDA = 66, LDB = 12;
};
template <typename T> constexpr int gemm_threads() { return (GemmCfg<T>::TM / GemmCfg<T>::WM) * (GemmCfg<T>::TN / GemmCfg<T>::WN) * 32; }
template <typename T> constexpr int gemm_smem_bytes() {
  // operand pipeline, or the C tile staged for the epilogue, whichever is larger
  constexpr int pipe = GemmCfg<T>::ST * (GemmCfg<T>::KC * GemmCfg<T>::LDA + GemmCfg<T>::TN * GemmCfg<T>::LDB) * (int)sizeof(T);
  constexpr int ctile = GemmCfg<T>::TN * (GemmCfg<T>::TM + 2) * (int)sizeof(T);
  return pipe > ctile ? pipe : ctile;
}

// GEN = true: free-standing batched product C ∓= A·B described by GemmDesc (compressed path); blockIdx.x is the item.
struct GemmDesc {
  long long a, b, c;  // element offsets from `pool`: A is M×K (lda), B is K×N (ldb), C is M×N (ldc), all column-major
  int lda, ldb, ldc;
  int M, N, K;
  int sign;           // -1: C -= A·B, +1: C += A·B
  int pad;
};

template <typename T, bool GEN = false>
__global__ void __launch_bounds__(GemmCfg<T>::TM / GemmCfg<T>::WM * GemmCfg<T>::TN / GemmCfg<T>::WN * 32) k_gemm(const Front* __restrict__ fronts, T* __restrict__ pool, int f0, int J0,
                                                  int j0, int NB, int W, int mode, const GemmDesc* __restrict__ gdesc) {
  using Cfg = GemmCfg<T>;
  constexpr int TM = Cfg::TM, TN = Cfg::TN, WM = Cfg::WM, WN = Cfg::WN, KC = Cfg::KC, ST = Cfg::ST, LDA = Cfg::LDA,
                LDB = Cfg::LDB;
  constexpr bool CX = hs_traits<T>::is_complex;
  constexpr int EPV = 16 / (int)sizeof(T);  // elements per 16-byte vector: 2 (f64) or 1 (c64)
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int NT = (TM / WM) * (TN / WN) * 32;  // threads per CTA
  int kbase, kcount, lo, rhi, clo, chi;  // C = rows [lo, rhi) × cols [clo, chi)
  const T *FA, *FB;   // A(m, k) = FA[k·lda + m], B(k, c) = FB[c·ldb + k]
  T* F;               // C(m, c) = F[c·ld + m]
  long long lda, ldb, ld;
  bool al, plus = false;
  if constexpr (GEN) {
    const GemmDesc gd = gdesc[blockIdx.x];
    kbase = 0; kcount = gd.K; lo = 0; rhi = gd.M; clo = 0; chi = gd.N;
    FA = pool + gd.a; FB = pool + gd.b; F = pool + gd.c;
    lda = gd.lda; ldb = gd.ldb; ld = gd.ldc;
    al = ((gd.a | gd.b | gd.c | gd.lda | gd.ldb | gd.ldc) & (EPV - 1)) == 0;
    plus = gd.sign > 0;
    if (kcount <= 0) return;
  } else {
    const Front fr = fronts[f0 + blockIdx.x];
    const int n = fr.n, ni = fr.ni;
    const int BE = min(J0 + NB, ni);
    if (mode == 1 || mode == 3 || mode == 4) {
      if (ni <= J0) return;
      const int BE2 = min(BE + NB, ni);
      kbase = J0; kcount = BE - J0; lo = BE; rhi = n;
      clo = mode == 4 ? BE2 : BE;
      chi = mode == 3 ? BE2 : n;
    } else {
      if (ni <= j0) return;
      const int wc = min(W, ni - j0);
      kbase = j0; kcount = wc; lo = j0 + wc;
      if (mode == 2) { rhi = BE; clo = BE; chi = n; }
      else {
        // mode 0 split by rows at the end of the rows pivots are taken from: 5 = the pivot rows (what the next panel
        // waits for), 6 = the rows below (boundary rows; they ride on a second stream, off the panel chain)
        const int plim = hs_plim(fr, j0);
        clo = lo; chi = BE; rhi = n;
        if (mode == 5) rhi = plim;
        else if (mode == 6) lo = max(lo, plim);
      }
    }
    // rows are loaded in aligned 16-byte pairs; a front whose base is not 16-byte aligned (only the root-boundary
    // pseudo front can be) takes the 8-byte copy path
    al = (fr.off & (EPV - 1)) == 0;
    F = pool + fr.off;
    FA = F; FB = F;
    lda = ldb = ld = fr.ld;
  }
  if (lo >= rhi || clo >= chi) return;
  const int rbase = al ? (lo & ~(EPV - 1)) : lo;
  const int m0 = rbase + blockIdx.y * TM, n0 = clo + blockIdx.z * TN;
  if (m0 >= rhi || n0 >= chi) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  constexpr int STAGE = KC * LDA + TN * LDB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp % (TM / WM)) * WM, wn = (warp / (TM / WM)) * WN;
  const int g = lane >> 2, q = lane & 3;

  // warm L2 with the C tile: the epilogue's read-modify-write then does not wait on HBM
  {
    const int lines_per_col = (TM * (int)sizeof(T)) / 128;
    for (int e = tid; e < TN * lines_per_col; e += NT) {
      const int c = n0 + e / lines_per_col, r = m0 + (e % lines_per_col) * (128 / (int)sizeof(T));
      if (c < chi && r < rhi) asm volatile("prefetch.global.L2 [%0];" ::"l"(F + (long long)c * ld + r));
    }
  }

  auto load_stage = [&](int stage, int chunk) {
    T* As = smem + stage * STAGE;
    T* Bs = As + KC * LDA;
    const int krem = kcount - chunk * KC;
    const int kg = kbase + chunk * KC;
    if (EPV > 1 && !al) {
      for (int c = tid; c < KC * TM; c += NT) {
        const int k = c / TM, m = c % TM;
        const bool ok = k < krem && m0 + m < rhi;
        cp_async8(As + k * LDA + m, ok ? FA + (long long)(kg + k) * lda + (m0 + m) : FA, ok ? 8 : 0);
      }
      for (int c = tid; c < TN * KC; c += NT) {
        const int nn = c / KC, k = c % KC;
        const bool ok = n0 + nn < chi && k < krem;
        cp_async8(Bs + nn * LDB + k, ok ? FB + (long long)(n0 + nn) * ldb + (kg + k) : FB, ok ? 8 : 0);
      }
      return;
    }
    constexpr int AV = TM / EPV;  // 16-byte vectors per k-row of the A tile
#pragma unroll
    for (int i = 0; i < (KC * AV) / NT; ++i) {
      const int c = tid + i * NT;
      const int k = c / AV, m = (c % AV) * EPV;
      const int left = rhi - (m0 + m);
      int bytes = 0;
      if (k < krem && left > 0) bytes = left >= EPV ? 16 : (int)sizeof(T);
      const T* src = bytes ? FA + (long long)(kg + k) * lda + (m0 + m) : FA;
      cp_async16(As + k * LDA + m, src, bytes);
    }
    constexpr int BV = KC / EPV;  // 16-byte vectors per column of the B tile
#pragma unroll
    for (int i = 0; i < (TN * BV) / NT; ++i) {
      const int c = tid + i * NT;
      const int nn = c / BV, k = (c % BV) * EPV;
      int bytes = 0;
      if (n0 + nn < chi && k < krem) bytes = (krem - k) >= EPV ? 16 : (int)sizeof(T);
      const T* src = bytes ? FB + (long long)(n0 + nn) * ldb + (kg + k) : FB;
      cp_async16(Bs + nn * LDB + k, src, bytes);
    }
  };

  double acc[CX ? 2 : 1][MI][NI][2];
#pragma unroll
  for (int z = 0; z < (CX ? 2 : 1); ++z)
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) acc[z][i][j][0] = acc[z][i][j][1] = 0.0;

  const int nchunks = (kcount + KC - 1) / KC;
#pragma unroll
  for (int s = 0; s < ST - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<ST - 2>();
    __syncthreads();
    if (c + ST - 1 < nchunks) load_stage((c + ST - 1) % ST, c + ST - 1);
    cp_async_commit();
    const T* As = smem + (c % ST) * STAGE;
    const T* Bs = As + KC * LDA;
#pragma unroll
    for (int ks = 0; ks < KC / 4; ++ks) {
      const int k = ks * 4 + q;
      T af[MI], bf[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) af[i] = As[k * LDA + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < NI; ++j) bf[j] = Bs[(wn + j * 8 + g) * LDB + k];
      if constexpr (!CX) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) dmma884(acc[0][i][j][0], acc[0][i][j][1], af[i], bf[j]);
      } else {
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const double nai = -af[i].y;
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            dmma884(acc[0][i][j][0], acc[0][i][j][1], af[i].x, bf[j].x);
            dmma884(acc[0][i][j][0], acc[0][i][j][1], nai, bf[j].y);
            dmma884(acc[1][i][j][0], acc[1][i][j][1], af[i].x, bf[j].y);
            dmma884(acc[1][i][j][0], acc[1][i][j][1], af[i].y, bf[j].x);
          }
        }
      }
    }
  }
  cp_async_wait<0>();
  // epilogue: C -= acc.  The C tile is pulled into the (now idle) pipeline buffers with one burst of cp.async — one
  // L2/HBM round trip instead of one per row group — then read-modify-written from shared memory.
  __syncthreads();  // every warp is done with the operand stages
  constexpr int LDC = TM + 2;  // ≡ 2 (mod 8) words of 8 B: conflict-free for the accumulator fragment layout
  T* Cs = smem;                // [TN][LDC]
  if (al) {
    constexpr int CV = TM / EPV;
#pragma unroll 4
    for (int c = tid; c < TN * CV; c += NT) {
      const int nn = c / CV, m = (c % CV) * EPV;
      const int left = rhi - (m0 + m);
      int bytes = 0;
      if (n0 + nn < chi && left > 0) bytes = left >= EPV ? 16 : (int)sizeof(T);
      const T* src = bytes ? F + (long long)(n0 + nn) * ld + (m0 + m) : F;
      cp_async16(Cs + nn * LDC + m, src, bytes);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int rl = wm + i * 8 + g, r = m0 + rl;
    const bool rok = r >= lo && r < rhi;
    T cv[NI][2];
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cl = wn + j * 8 + q * 2 + h, c = n0 + cl;
        if (al) cv[j][h] = Cs[cl * LDC + rl];
        else cv[j][h] = (rok && c < chi) ? F[(long long)c * ld + r] : hs_zero<T>();
      }
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = n0 + wn + j * 8 + q * 2 + h;
        if (rok && c < chi) {
          T v = cv[j][h];
          if constexpr (!CX) {
            v = (GEN && plus) ? v + acc[0][i][j][h] : v - acc[0][i][j][h];
          } else {
            if (GEN && plus) { v.x += acc[0][i][j][h]; v.y += acc[1][i][j][h]; }
            else { v.x -= acc[0][i][j][h]; v.y -= acc[1][i][j][h]; }
          }
          F[(long long)c * ld + r] = v;
        }
      }
  }
}

// small helpers
template <typename T>
__global__ void k_copy_strided(const T* __restrict__ src, long long lds, T* __restrict__ dst, long long ldd, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[(long long)blockIdx.y * ldd + i] = src[(long long)blockIdx.y * lds + i];
}
