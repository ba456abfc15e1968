// Hand-written sm_100a kernels of the multifrontal factor + tree solve.
//
//   assembly   k_fill_owner / k_scatter_A / k_extend_add      (factorization.jl:33-40,115-123)  HBM-bound
//   LU panel   k_panel   — register-resident panel, one thread-block cluster per front,
//                          pivot search over distributed shared memory                (blockmatrix.jl:118, `\`)
//   row ops    k_swap_trsm — row interchanges + unit-lower triangular solve of the U row panel
//   update     k_gemm    — FP64 / complex-FP64 tensor-core (DMMA m8n8k4) Schur update C -= A·B
//                                                                                  (factorization.jl:40,72)
//   solve      k_solve_fwd / k_solve_bwd                        (factornode.jl:77-99)            HBM-bound
#pragma once

#include <cooperative_groups.h>

#include "hs_types.cuh"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------
// assembly
// ------------------------------------------------------------------------------------------------

// own[g] = front that holds DOF g on this level, pos[g] = its row inside that front
__global__ void k_fill_owner(const Front* __restrict__ fronts, const int* __restrict__ gidx, int* __restrict__ own,
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
  if (k >= fr.n) return;
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
template <typename T, int CB>
__global__ void __launch_bounds__(256) k_extend_add(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                     const int* __restrict__ cmap, int c0) {
  const int ci = c0 + blockIdx.x;
  const Front ch = fronts[ci];
  if (ch.parent < 0) return;
  const int nb = ch.n - ch.ni;
  const int b0 = blockIdx.y * CB;
  if (b0 >= nb) return;
  const Front pa = fronts[ch.parent];
  const int* map = cmap + ch.ioff + ch.ni;
  const T* S = pool + ch.off + (long long)ch.ni * ch.ld + ch.ni;
  T* P = pool + pa.off;
  __shared__ int s_pb[CB];
  if (threadIdx.x < CB) s_pb[threadIdx.x] = (b0 + threadIdx.x < nb) ? map[b0 + threadIdx.x] : -1;
  __syncthreads();
  for (int a = threadIdx.x; a < nb; a += blockDim.x) {
    const int pr = map[a];
    if (pr < 0) continue;
#pragma unroll
    for (int b = 0; b < CB; ++b) {
      const int pc = s_pb[b];
      if (pc >= 0) P[(long long)pc * pa.ld + pr] = S[(long long)(b0 + b) * ch.ld + a];
    }
  }
}

// rperm[k] = local row of the ORIGINAL pivot block that ends at row k after all interchanges (so P·x is a gather)
__global__ void k_rperm(const Front* __restrict__ fronts, const int* __restrict__ ipiv, int* __restrict__ rperm,
                        int f0) {
  extern __shared__ int s_p[];
  const Front fr = fronts[f0 + blockIdx.x];
  const int ni = fr.ni;
  for (int k = threadIdx.x; k < ni; k += blockDim.x) s_p[k] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int* pv = ipiv + fr.ioff;
    for (int k = 0; k < ni; ++k) {
      const int p = pv[k];
      if (p != k) { const int t = s_p[k]; s_p[k] = s_p[p]; s_p[p] = t; }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ni; k += blockDim.x) rperm[fr.ioff + k] = s_p[k];
}

// ------------------------------------------------------------------------------------------------
// row interchanges outside the panel + U row panel:  U12 = L11⁻¹ · (P·A)(j0:j0+wc, j0+wc:n)
// one thread per front column
// ------------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(128) k_swap_trsm(const Front* __restrict__ fronts, T* __restrict__ pool,
                                                    const int* __restrict__ ipiv, int f0, int j0) {
  const int fi = f0 + blockIdx.x;
  const Front fr = fronts[fi];
  if (fr.ni <= j0) return;
  const int wc = min(W, fr.ni - j0);
  const int nother = fr.n - wc;
  if ((int)(blockIdx.y * blockDim.x) >= nother) return;
  T* F = pool + fr.off;
  __shared__ T sL[W * W];
  __shared__ int spiv[W];
  for (int e = threadIdx.x; e < W * W; e += blockDim.x) {
    const int i = e % W, k = e / W;
    sL[e] = (i < wc && k < wc && i > k) ? F[(long long)(j0 + k) * fr.ld + (j0 + i)] : hs_zero<T>();
  }
  if (threadIdx.x < W) spiv[threadIdx.x] = threadIdx.x < wc ? ipiv[fr.ioff + j0 + threadIdx.x] : 0;
  __syncthreads();
  const int cc = blockIdx.y * blockDim.x + threadIdx.x;
  if (cc >= nother) return;
  const int c = cc < j0 ? cc : cc + wc;
  T* col = F + (long long)c * fr.ld;
  for (int j = 0; j < wc; ++j) {
    const int p = spiv[j];
    if (p != j0 + j) { const T t = col[j0 + j]; col[j0 + j] = col[p]; col[p] = t; }
  }
  if (c < j0) return;
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

// ------------------------------------------------------------------------------------------------
// Schur / trailing update on the FP64 tensor cores:  C(j0+wc:n, j0+wc:n) -= F(j0+wc:n, j0:j0+wc)·F(j0:j0+wc, j0+wc:n)
// DMMA m8n8k4; complex = 4 real DMMAs on interleaved operands.  64×64 tile per CTA, 4 warps of 32×32.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <typename T> struct GemmCfg;
template <> struct GemmCfg<double> { static constexpr int KC = 64, LDA = 68, LDB = 68; };
template <> struct GemmCfg<cplx> { static constexpr int KC = 32, LDA = 66, LDB = 36; };

template <typename T>
__global__ void __launch_bounds__(128) k_gemm(const Front* __restrict__ fronts, T* __restrict__ pool, int f0, int j0,
                                               int W) {
  constexpr int KC = GemmCfg<T>::KC, LDA = GemmCfg<T>::LDA, LDB = GemmCfg<T>::LDB;
  constexpr bool CX = hs_traits<T>::is_complex;
  const int fi = f0 + blockIdx.x;
  const Front fr = fronts[fi];
  if (fr.ni <= j0) return;
  const int wc = min(W, fr.ni - j0);
  const int t0 = j0 + wc;          // first trailing row/col
  const int mt = fr.n - t0;        // trailing size
  const int m0 = blockIdx.y * 64, n0 = blockIdx.z * 64;
  if (m0 >= mt || n0 >= mt) return;
  T* F = pool + fr.off;
  const long long ld = fr.ld;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* As = reinterpret_cast<T*>(smem_raw);  // [KC][LDA]  (k-major, m contiguous)
  T* Bs = As + KC * LDA;                   // [64][LDB]  (n-major, k contiguous)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  const int g = lane >> 2, q = lane & 3;

  double acc[CX ? 2 : 1][4][4][2];
#pragma unroll
  for (int z = 0; z < (CX ? 2 : 1); ++z)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[z][i][j][0] = acc[z][i][j][1] = 0.0;

  for (int kc = 0; kc < wc; kc += KC) {
    const int kn = min(KC, wc - kc);
    // A tile: rows t0+m0.. (64), cols j0+kc.. (KC)
    for (int e = tid; e < 64 * KC; e += 128) {
      const int mm = e & 63, kk = e >> 6;
      T v = hs_zero<T>();
      if (m0 + mm < mt && kk < kn) v = F[(long long)(j0 + kc + kk) * ld + (t0 + m0 + mm)];
      As[kk * LDA + mm] = v;
    }
    // B tile: rows j0+kc.. (KC), cols t0+n0.. (64)
    for (int e = tid; e < 64 * KC; e += 128) {
      const int kk = e % KC, nn = e / KC;
      T v = hs_zero<T>();
      if (n0 + nn < mt && kk < kn) v = F[(long long)(t0 + n0 + nn) * ld + (j0 + kc + kk)];
      Bs[nn * LDB + kk] = v;
    }
    __syncthreads();
    const int ksteps = (kn + 3) >> 2;
    for (int ks = 0; ks < ksteps; ++ks) {
      const int k = ks * 4 + q;
      T af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[k * LDA + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn + j * 8 + g) * LDB + k];
      if constexpr (!CX) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[0][i][j][0], acc[0][i][j][1], af[i], bf[j]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double nai = -af[i].y;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            dmma884(acc[0][i][j][0], acc[0][i][j][1], af[i].x, bf[j].x);
            dmma884(acc[0][i][j][0], acc[0][i][j][1], nai, bf[j].y);
            dmma884(acc[1][i][j][0], acc[1][i][j][1], af[i].x, bf[j].y);
            dmma884(acc[1][i][j][0], acc[1][i][j][1], af[i].y, bf[j].x);
          }
        }
      }
    }
    __syncthreads();
  }
  // epilogue: C -= acc
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm + i * 8 + g;
    if (r >= mt) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = n0 + wn + j * 8 + q * 2 + h;
        if (c >= mt) continue;
        T* pc = F + (long long)(t0 + c) * ld + (t0 + r);
        if constexpr (!CX) {
          *pc -= acc[0][i][j][h];
        } else {
          T v = *pc;
          v.x -= acc[0][i][j][h];
          v.y -= acc[1][i][j][h];
          *pc = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tree solve, one level per launch, one CTA per (front, right-hand side)
// ------------------------------------------------------------------------------------------------
// forward (factornode.jl:77-82 fused with the L half of :89-99):
//   t = L11⁻¹·P·x[int];  x[bnd] -= L21·t;  x[int] = t
template <typename T>
__global__ void __launch_bounds__(256) k_solve_fwd(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                    const int* __restrict__ gidx, const int* __restrict__ rperm,
                                                    T* __restrict__ x, long long ldx, T* __restrict__ work,
                                                    long long wstride, long long ioff0, int f0) {
  const Front fr = fronts[f0 + blockIdx.x];
  const int n = fr.n, ni = fr.ni;
  if (ni == 0) return;
  const long long ld = fr.ld;
  const T* F = pool + fr.off;
  T* xr = x + (long long)blockIdx.y * ldx;
  T* w = work + (long long)blockIdx.y * wstride + (fr.ioff - ioff0);
  const int* gi = gidx + fr.ioff;
  const int* rp = rperm + fr.ioff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ T sw[32];
  for (int k = tid; k < n; k += 256) w[k] = xr[gi[k < ni ? rp[k] : k]];
  __syncthreads();
  for (int jb = 0; jb < ni; jb += 32) {
    const int bw = min(32, ni - jb);
    if (warp == 0) {
      T v = lane < bw ? w[jb + lane] : hs_zero<T>();
      for (int k = 0; k < bw - 1; ++k) {
        const T vk = hs_shfl(v, k);
        if (lane > k && lane < bw) v = hs_fnma(v, F[(long long)(jb + k) * ld + (jb + lane)], vk);
      }
      if (lane < bw) w[jb + lane] = v;
      sw[lane] = v;
    }
    __syncthreads();
    for (int r = jb + bw + tid; r < n; r += 256) {
      T acc = hs_zero<T>();
      for (int k = 0; k < bw; ++k) acc = hs_fma(acc, F[(long long)(jb + k) * ld + r], sw[k]);
      w[r] = hs_sub(w[r], acc);
    }
    __syncthreads();
  }
  for (int k = tid; k < n; k += 256) xr[gi[k]] = w[k];
}

// backward (U half of factornode.jl:89-99 fused with :83-88):  x[int] = U11⁻¹·(t − U12·x[bnd])
template <typename T>
__global__ void __launch_bounds__(256) k_solve_bwd(const Front* __restrict__ fronts, const T* __restrict__ pool,
                                                    const int* __restrict__ gidx, T* __restrict__ x, long long ldx,
                                                    T* __restrict__ work, long long wstride, long long ioff0,
                                                    int f0) {
  const Front fr = fronts[f0 + blockIdx.x];
  const int n = fr.n, ni = fr.ni;
  if (ni == 0) return;
  const long long ld = fr.ld;
  const T* F = pool + fr.off;
  T* xr = x + (long long)blockIdx.y * ldx;
  T* w = work + (long long)blockIdx.y * wstride + (fr.ioff - ioff0);
  const int* gi = gidx + fr.ioff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ T sw[32];
  for (int k = tid; k < n; k += 256) w[k] = xr[gi[k]];
  __syncthreads();
  for (int i = tid; i < ni; i += 256) {
    T acc = hs_zero<T>();
    for (int c = ni; c < n; ++c) acc = hs_fma(acc, F[(long long)c * ld + i], w[c]);
    w[i] = hs_sub(w[i], acc);
  }
  __syncthreads();
  const int nblk = (ni + 31) / 32;
  for (int b = nblk - 1; b >= 0; --b) {
    const int jb = b * 32;
    const int bw = min(32, ni - jb);
    if (warp == 0) {
      T v = lane < bw ? w[jb + lane] : hs_zero<T>();
      for (int k = bw - 1; k >= 0; --k) {
        if (lane == k) v = hs_mul(v, hs_recip(F[(long long)(jb + k) * ld + (jb + k)]));
        const T vk = hs_shfl(v, k);
        if (lane < k) v = hs_fnma(v, F[(long long)(jb + k) * ld + (jb + lane)], vk);
      }
      if (lane < bw) w[jb + lane] = v;
      sw[lane] = v;
    }
    __syncthreads();
    for (int r = tid; r < jb; r += 256) {
      T acc = hs_zero<T>();
      for (int k = 0; k < bw; ++k) acc = hs_fma(acc, F[(long long)(jb + k) * ld + r], sw[k]);
      w[r] = hs_sub(w[r], acc);
    }
    __syncthreads();
  }
  for (int k = tid; k < ni; k += 256) xr[gi[k]] = w[k];
}

// small helpers
template <typename T>
__global__ void k_copy_strided(const T* __restrict__ src, long long lds, T* __restrict__ dst, long long ldd, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[(long long)blockIdx.y * ldd + i] = src[(long long)blockIdx.y * lds + i];
}
