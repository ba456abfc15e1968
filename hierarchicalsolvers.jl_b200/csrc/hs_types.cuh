// Scalar types and device-side front descriptor shared by all kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// interleaved complex float64 (Julia ComplexF64 / numpy complex128 layout)
struct __align__(16) cplx {
  double x, y;
};

// One eliminated front (the device image of a reference FactorNode, factornode.jl:7-22).
// The front is ONE column-major n×n matrix, rows/cols ordered [int (ni); bnd (nb)]:
//
//        [ LU(A_ii)      U12 = L11⁻¹·P·A_ib ]      ni
//   F =  [ L21 = A_bi·U11⁻¹   S = A_bb − L21·U12 ]  nb
//
// so D, L = L21·L11⁻¹·P, R = U11⁻¹·U12 and S of the reference are all recoverable from it.
struct Front {
  long long off;   // element offset of F in the pool
  long long ioff;  // offset of this front's slice in the per-row int arrays (gidx, ipiv, rperm, cmap); length n
  int n, ni, ld;
  int parent;      // front id of the parent, -1 for the root
  int ni_l, nb_l;  // branch: #int / #bnd rows that come from the left child; leaf: ni_l = -1
  int flags;       // bit0: pseudo front (root Schur block, factornode.jl:72) — no assembly
  int split;       // > 0: pivoting is restricted to the two diagonal blocks [0, split) and [split, ni) of the pivot block —
                   // the 2×2 block elimination of blockfactor (blockmatrix.jl:115-120) — used when the pivot block has more
                   // rows than one panel cluster covers; 0: partial pivoting over the whole pivot block
};
// rows a pivot may be taken from while panel j0 is factored: [j0, hs_plim(fr, j0))
__host__ __device__ __forceinline__ int hs_plim(const Front& fr, int j0) { return (fr.split > 0 && j0 < fr.split) ? fr.split : fr.ni; }

template <typename T> struct hs_traits;
template <> struct hs_traits<double> { static constexpr bool is_complex = false; };
template <> struct hs_traits<cplx> { static constexpr bool is_complex = true; };

__host__ __device__ __forceinline__ double hs_zero(double*) { return 0.0; }
__host__ __device__ __forceinline__ cplx hs_zero(cplx*) { return cplx{0.0, 0.0}; }
template <typename T> __host__ __device__ __forceinline__ T hs_zero() { return hs_zero((T*)nullptr); }
template <typename T> __host__ __device__ __forceinline__ T hs_one();
template <> __host__ __device__ __forceinline__ double hs_one<double>() { return 1.0; }
template <> __host__ __device__ __forceinline__ cplx hs_one<cplx>() { return cplx{1.0, 0.0}; }

// |re| + |im| for complex, as LAPACK izamax/zgetf2 pick pivots (cabs1); |x| for real
__host__ __device__ __forceinline__ double hs_abs1(double a) { return fabs(a); }
__host__ __device__ __forceinline__ double hs_abs1(cplx a) { return fabs(a.x) + fabs(a.y); }

__host__ __device__ __forceinline__ double hs_mul(double a, double b) { return a * b; }
__host__ __device__ __forceinline__ cplx hs_mul(cplx a, cplx b) { return cplx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__host__ __device__ __forceinline__ double hs_add(double a, double b) { return a + b; }
__host__ __device__ __forceinline__ cplx hs_add(cplx a, cplx b) { return cplx{a.x + b.x, a.y + b.y}; }
__host__ __device__ __forceinline__ double hs_sub(double a, double b) { return a - b; }
__host__ __device__ __forceinline__ cplx hs_sub(cplx a, cplx b) { return cplx{a.x - b.x, a.y - b.y}; }
// acc - a*b
__host__ __device__ __forceinline__ double hs_fnma(double acc, double a, double b) { return fma(-a, b, acc); }
__host__ __device__ __forceinline__ cplx hs_fnma(cplx acc, cplx a, cplx b) {
  cplx r;
  r.x = fma(-a.x, b.x, acc.x);
  r.x = fma(a.y, b.y, r.x);
  r.y = fma(-a.x, b.y, acc.y);
  r.y = fma(-a.y, b.x, r.y);
  return r;
}
// acc + a*b
__host__ __device__ __forceinline__ double hs_fma(double acc, double a, double b) { return fma(a, b, acc); }
__host__ __device__ __forceinline__ cplx hs_fma(cplx acc, cplx a, cplx b) {
  cplx r;
  r.x = fma(a.x, b.x, acc.x);
  r.x = fma(-a.y, b.y, r.x);
  r.y = fma(a.x, b.y, acc.y);
  r.y = fma(a.y, b.x, r.y);
  return r;
}
__host__ __device__ __forceinline__ double hs_recip(double a) { return 1.0 / a; }
__host__ __device__ __forceinline__ cplx hs_recip(cplx a) {
  // Smith's algorithm, as LAPACK dladiv-style division of 1 by a
  if (fabs(a.x) >= fabs(a.y)) {
    double r = a.y / a.x, d = a.x + a.y * r;
    return cplx{1.0 / d, -r / d};
  } else {
    double r = a.x / a.y, d = a.y + a.x * r;
    return cplx{r / d, -1.0 / d};
  }
}
__host__ __device__ __forceinline__ bool hs_iszero(double a) { return a == 0.0; }
__host__ __device__ __forceinline__ bool hs_iszero(cplx a) { return a.x == 0.0 && a.y == 0.0; }

#ifdef __CUDACC__
// Reciprocal of a pivot for the elimination kernels, where every thread needs it on the per-column critical path:
// MUFU.RCP64H seed (≈20 bits) + two Newton steps instead of the ~30-instruction IEEE division sequence.  Within 1–2 ulp
// of 1/a; values outside the comfortable exponent range take the exact division.
__device__ __forceinline__ double hs_recip_pivot(double a) {
  const double aa = fabs(a);
  if (!(aa > 1e-280 && aa < 1e280)) return 1.0 / a;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ cplx hs_recip_pivot(cplx a) {
  // 1/a = conj(a)/|a|², scaled by the larger component as in Smith's algorithm to stay clear of overflow
  if (fabs(a.x) >= fabs(a.y)) {
    const double ix = hs_recip_pivot(a.x);
    const double r = a.y * ix, d = hs_recip_pivot(fma(a.y, r, a.x));
    return cplx{d, -r * d};
  } else {
    const double iy = hs_recip_pivot(a.y);
    const double r = a.x * iy, d = hs_recip_pivot(fma(a.x, r, a.y));
    return cplx{r * d, -d};
  }
}

// warp arg-max of a non-negative double (or -1 = "no candidate") with ties resolved to the smallest index, through
// three 32-bit REDUX operations instead of five rounds of 64-bit shuffles + compares.  Non-negative IEEE doubles order
// like their bit patterns; -1.0 is mapped to key 0 and can never win against a real candidate (|v| ≥ 0 maps to ≥ 1).
__device__ __forceinline__ void warp_argmax(double& val, int& idx) {
  const unsigned long long bits = val < 0.0 ? 0ull : (unsigned long long)__double_as_longlong(val) + 1ull;
  const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  const bool win = hi == mhi && lo == mlo;
  const unsigned widx = __reduce_min_sync(0xffffffffu, win ? (unsigned)idx : 0xffffffffu);
  const unsigned long long mb = ((unsigned long long)mhi << 32) | mlo;
  val = mb == 0ull ? -1.0 : __longlong_as_double((long long)(mb - 1ull));
  idx = (int)widx;
}
__device__ __forceinline__ double hs_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ cplx hs_shfl(cplx v, int src) {
  return cplx{__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)};
}
#endif
