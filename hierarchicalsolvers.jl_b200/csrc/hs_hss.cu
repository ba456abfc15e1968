// HSS storage of the Schur complements of compressed fronts — reference src/factorization.jl:102-110 and :228-249.
//
// The reference hands the matrix-free operator  S[perm,perm] = P·(Abb − (Abi·R.U)·R.Vᴴ)·Pᵀ  (`_schur_complement` :228,
// `_sample_schur!` :239, `_getindex_schur` :246) to `HssMatrices.randcompress_adaptive(Smap, cl, cl; kest, atol, rtol)`
// (:110) with cl = bisection_cluster((|int_loc|, n); leafsize) (:109).  HssMatrices.jl is not vendored; the algorithm
// restated here is the published one (Martinsson 2011, adaptive variant of Gorman et al. 2019) in exactly the form the test
// suite's CPU restatement (randcompress_adaptive) takes, which the parity tests run with the SAME host-supplied Gaussian matrices:
//
//   1. sketches  Sr = S·Ω,  Sc = Sᴴ·Ψ  with k = kest + 10 columns — matrix-free, tensor-core GEMMs (k_gemm descriptor mode):
//          Sr = Abb·Ωz − Z·(Ri·Ωz),     Scᴴ = Ψzᴴ·Abb − (Ψzᴴ·Z)·Ri,       Z = Abi·(Aii⁻¹Qi) from hs_compress.cu;
//      dense S is never formed.
//   2. leaves (bottom level of the cluster tree):  D = S[I,I] by entry evaluation, Sr_loc = Sr[I] − D·Ω[I],
//      Sc_loc = Sc[I] − Dᴴ·Ψ[I], row interpolative decompositions  Sr_loc ≈ U·Sr_loc[skel]  (Householder QR with column
//      pivoting of Sr_locᴴ, truncated by pqrfact's rule |R[k,k]| ≤ max(atol, rtol·|R[1,1]|)), likewise V.
//   3. branches, bottom up:  B12 = S[Iskel₁, Jskel₂], B21 = S[Iskel₂, Jskel₁] by entry evaluation at the skeleton index
//      sets, reduced samples  [Sr₁ − B12·Om₂; Sr₂ − B21·Om₁],  IDs of those give the translation operators R, W.
//   4. if any detected rank reaches k − 10 the front is sampled again with k + stepsize columns (same Ω, Ψ, more columns).
//
// Layout choices for the GPU.  All per-node sample blocks are kept TRANSPOSED (k × m, one sample row of S per column), so
// the columns the pivoted QR walks are contiguous; the reduction steps then read  Y ∓= X·Bᵀ  — plain column-major GEMMs for
// the DMMA kernel.  Children write their skeleton columns straight into their parent's concatenated block at a fixed
// offset (left child at 0, right child at cap(left)); the unused columns in between stay zero, are never chosen as pivots
// and get zero interpolation coefficients, so no offset depends on a rank and nothing has to be synchronised with the
// host inside a round.  All fronts of a tree level and all HSS nodes of one height run in the same launches.
//
// After the construction the generators are compacted into the level's store (true ranks), and the dense matrix the HSS
// form represents is written back into the front's dense slot: the parent assembles from the APPROXIMATED blocks, as the
// reference's HSS `_assemble_blocks` (:126-140) does.  The parent's low-rank Gauss transforms then start from this
// matrix's generators (hs_compress.cu, :184-209).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <vector>

#include "hs_fac.cuh"
#include "hs_kernels.cuh"

namespace {

inline long long up32(long long v) { return (v + 31) / 32 * 32; }
inline int even_up(int v) { return std::max(2, (v + 1) & ~1); }

__device__ __forceinline__ double habs2(double a) { return a * a; }
__device__ __forceinline__ double habs2(cplx a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ double hconj(double a) { return a; }
__device__ __forceinline__ cplx hconj(cplx a) { return cplx{a.x, -a.y}; }
// acc + conj(a)·b
__device__ __forceinline__ double hcjfma(double acc, double a, double b) { return fma(a, b, acc); }
__device__ __forceinline__ cplx hcjfma(cplx acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
  return acc;
}
__device__ __forceinline__ double hwsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ cplx hwsum(cplx v) { return cplx{hwsum(v.x), hwsum(v.y)}; }
__device__ __forceinline__ double hscal(double a, double s) { return a * s; }
__device__ __forceinline__ cplx hscal(cplx a, double s) { return cplx{a.x * s, a.y * s}; }
__device__ __forceinline__ double hfromreal(double r, double*) { return r; }
__device__ __forceinline__ cplx hfromreal(double r, cplx*) { return cplx{r, 0.0}; }
__device__ __forceinline__ double hmag(double a) { return fabs(a); }
__device__ __forceinline__ double hmag(cplx a) { return hypot(a.x, a.y); }

// ---- device records of one adaptive round ---------------------------------------------------------------------------
struct HFront {
  long long abb, z, ri;                                        // Abb(0,0), Z(0,0), Ri(0,0): element offsets from pool
  long long omz, psh, yu, cp, t1, t2, gst, gsct, gomt, gpst;   // sketch workspaces (see build_round)
  int ldf, ldz, ldri, nbld, kld, r2ld;
  int nb, r2, m, k;
  int perm;        // offset of this front's perm inside d_hperm
  int flag;        // ints[flag]: 1 when a detected rank saturated the sample count
  int pad0, pad1;
};
struct HNode {
  int front, lo, hi, left, right, parent, isright, leaf;
  int cap, mcat, capl, lde;
  int ldm0, ldmt, ldm1, ldm1t;
  int icat[2];     // ints offsets of the concatenated skeleton index lists (branches)
  int rk;          // ints[rk + side]: detected ranks
  int pad;
  long long ycat[2], xcat[2], wq[2], e[2], ec[2];
  long long m0, mt, mc, m1, m1t, m1c;
};

// ---- counter-based Gaussian generator (used when the caller supplies no sketch matrices) -----------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {  // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
template <typename T>
__global__ void k_sk_generate(T* __restrict__ out, long long rows, long long cols, long long c0, unsigned long long seed) {
  // entry (i, c) of matrix `w` (0 = Ω, 1 = Ψ) depends only on (seed, w, i, c): more columns can be appended later
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = rows * (cols - c0);
  if (e >= 2 * per) return;
  const int w = (int)(e / per);
  const long long q = e % per, i = q % rows, c = c0 + q / rows;
  const unsigned long long h1 = mix64(seed ^ mix64(((unsigned long long)c << 34) ^ ((unsigned long long)i << 1) ^ (unsigned long long)w));
  const unsigned long long h2 = mix64(h1);
  const double u1 = ((double)(h1 >> 11) + 1.0) * (1.0 / 9007199254740993.0);   // (0, 1)
  const double u2 = (double)(h2 >> 11) * (1.0 / 9007199254740992.0);
  const double g = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  out[(long long)w * rows * cols + c * rows + i] = hfromreal(g, (T*)nullptr);
}

// ---- 1. sketch preparation --------------------------------------------------------------------------------------------
// Ωz[perm[i], c] = Ω[i, c] (rows of S in the front's bnd order),  Ψzᴴ[c, perm[i]] = conj(Ψ[i, c]),
// Ωt[c, i] = Ω[i, c],  Ψt[c, i] = Ψ[i, c]  (the leaves' sample blocks of the test matrices, transposed)
template <typename T>
__global__ void __launch_bounds__(256) k_sk_fill(const HFront* __restrict__ fronts, T* __restrict__ pool, const int* __restrict__ hperm,
                                                  const T* __restrict__ om, const T* __restrict__ ps, long long sk_rows) {
  const HFront fr = fronts[blockIdx.x];
  const int i = blockIdx.y * 256 + threadIdx.x;
  if (i >= fr.m) return;
  const int p = hperm[fr.perm + i];
  T* omz = pool + fr.omz + p;
  T* psh = pool + fr.psh + (long long)p * fr.kld;
  T* gom = pool + fr.gomt + (long long)i * fr.kld;
  T* gps = pool + fr.gpst + (long long)i * fr.kld;
  for (int c = 0; c < fr.k; ++c) {
    const T a = om[(long long)c * sk_rows + i], b = ps[(long long)c * sk_rows + i];
    omz[(long long)c * fr.nbld] = a;
    psh[c] = hconj(b);
    gom[c] = a;
    gps[c] = b;
  }
}
// St[c, i] = (S·Ω)[perm[i], c],   Sct[c, i] = conj((Ψᴴ·S)[c, perm[i]]) = (Sᴴ·Ψ)[perm[i], c]
template <typename T>
__global__ void __launch_bounds__(256) k_sk_gather(const HFront* __restrict__ fronts, T* __restrict__ pool, const int* __restrict__ hperm) {
  const HFront fr = fronts[blockIdx.x];
  const int i = blockIdx.y * 256 + threadIdx.x;
  if (i >= fr.m) return;
  const int p = hperm[fr.perm + i];
  const T* yu = pool + fr.yu + p;
  const T* cp = pool + fr.cp + (long long)p * fr.kld;
  T* gst = pool + fr.gst + (long long)i * fr.kld;
  T* gsc = pool + fr.gsct + (long long)i * fr.kld;
  for (int c = 0; c < fr.k; ++c) {
    gst[c] = yu[(long long)c * fr.nbld];
    gsc[c] = hconj(cp[c]);
  }
}

// ---- 2. entry evaluation (`_getindex_schur`, factorization.jl:246-249) -------------------------------------------------
// leaf: D = S[I, I];  branch: B12 = S[Iskel(left), Jskel(right)] (which = 0), B21 = S[Iskel(right), Jskel(left)] (which = 1).
// Written three ways: M, Mᵀ and conj(M) — the reduction GEMMs need Mᵀ (row side) and conj(M) (column side).
template <typename T>
__global__ void __launch_bounds__(256) k_hss_entries(const HNode* __restrict__ nodes, const int* __restrict__ list,
                                                      const HFront* __restrict__ fronts, T* __restrict__ pool,
                                                      const int* __restrict__ ints, const int* __restrict__ hperm) {
  const HNode nd = nodes[list[blockIdx.x]];
  const int which = blockIdx.y;
  if (nd.leaf && which) return;
  const HFront fr = fronts[nd.front];
  int nr, nc, ro = 0, co = 0;
  if (nd.leaf) { nr = nc = nd.hi - nd.lo; }
  else {
    const HNode l = nodes[nd.left], r = nodes[nd.right];
    const int rl0 = ints[l.rk], rl1 = ints[l.rk + 1], rr0 = ints[r.rk], rr1 = ints[r.rk + 1];
    if (!which) { nr = rl0; nc = rr1; ro = 0; co = nd.capl; } else { nr = rr0; nc = rl1; ro = nd.capl; co = 0; }
  }
  T* M = pool + (which ? nd.m1 : nd.m0);
  T* Mt = pool + (which ? nd.m1t : nd.mt);
  T* Mc = pool + (which ? nd.m1c : nd.mc);
  const int ldm = which ? nd.ldm1 : nd.ldm0, ldt = which ? nd.ldm1t : nd.ldmt;
  const int* perm = hperm + fr.perm;
  const int* ri = ints + nd.icat[0] + ro;
  const int* ci = ints + nd.icat[1] + co;
  const T* Abb = pool + fr.abb;
  const T* Z = pool + fr.z;
  const T* Ri = pool + fr.ri;
  const long long total = (long long)nr * nc;
  for (long long e = threadIdx.x; e < total; e += 256) {
    const int a = (int)(e % nr), b = (int)(e / nr);
    const int i = nd.leaf ? nd.lo + a : ri[a], j = nd.leaf ? nd.lo + b : ci[b];
    const int pi = perm[i], pj = perm[j];
    T v = Abb[(long long)pj * fr.ldf + pi];
    const T* zr = Z + pi;
    const T* rc = Ri + (long long)pj * fr.ldri;
    for (int l = 0; l < fr.r2; ++l) v = hs_fnma(v, zr[(long long)l * fr.ldz], rc[l]);
    M[(long long)b * ldm + a] = v;
    Mt[(long long)a * ldt + b] = v;
    Mc[(long long)b * ldm + a] = hconj(v);
  }
}

// ---- 3. GEMM descriptors built on the device from the detected ranks ----------------------------------------------------
// (a) reduction of the sample blocks:   leaf   Y0 −= X1·Dᵀ,  Y1 −= X0·conj(D)
//     branch  Y0[:, left] −= X1[:, right]·B12ᵀ   Y0[:, right] −= X1[:, left]·B21ᵀ
//             Y1[:, left] −= X0[:, right]·conj(B21)   Y1[:, right] −= X0[:, left]·conj(B12)
__global__ void k_hss_desc_a(const HNode* __restrict__ nodes, const int* __restrict__ list, int nlist,
                             const HFront* __restrict__ fronts, const int* __restrict__ ints, GemmDesc* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nlist) return;
  const HNode nd = nodes[list[t]];
  const HFront fr = fronts[nd.front];
  GemmDesc d[4];
  for (int q = 0; q < 4; ++q) { d[q] = GemmDesc{}; d[q].sign = -1; }
  const long long kld = fr.kld;
  if (nd.parent >= 0 || !nd.leaf) {
    if (nd.leaf) {
      const int m = nd.hi - nd.lo;
      d[0].a = nd.xcat[1]; d[0].b = nd.mt; d[0].c = nd.ycat[0]; d[0].ldb = nd.ldmt; d[0].M = fr.k; d[0].N = m; d[0].K = m;
      d[1].a = nd.xcat[0]; d[1].b = nd.mc; d[1].c = nd.ycat[1]; d[1].ldb = nd.ldm0; d[1].M = fr.k; d[1].N = m; d[1].K = m;
    } else if (nd.parent >= 0) {
      const HNode l = nodes[nd.left], r = nodes[nd.right];
      const int ra0 = ints[l.rk], ra1 = ints[l.rk + 1], rb0 = ints[r.rk], rb1 = ints[r.rk + 1];
      const long long go = (long long)nd.capl * kld;
      d[0].a = nd.xcat[1] + go; d[0].b = nd.mt;  d[0].ldb = nd.ldmt;  d[0].c = nd.ycat[0];      d[0].M = fr.k; d[0].N = ra0; d[0].K = rb1;
      d[1].a = nd.xcat[1];      d[1].b = nd.m1t; d[1].ldb = nd.ldm1t; d[1].c = nd.ycat[0] + go; d[1].M = fr.k; d[1].N = rb0; d[1].K = ra1;
      d[2].a = nd.xcat[0] + go; d[2].b = nd.m1c; d[2].ldb = nd.ldm1;  d[2].c = nd.ycat[1];      d[2].M = fr.k; d[2].N = ra1; d[2].K = rb0;
      d[3].a = nd.xcat[0];      d[3].b = nd.mc;  d[3].ldb = nd.ldm0;  d[3].c = nd.ycat[1] + go; d[3].M = fr.k; d[3].N = rb1; d[3].K = ra0;
    }
  }
  for (int q = 0; q < 4; ++q) { d[q].lda = (int)kld; d[q].ldc = (int)kld; out[4 * t + q] = d[q]; }
}
// (b) the node's own test-matrix blocks in the new bases:  X0' = X0cat·conj(E0),  X1' = X1cat·conj(E1), written into the
//     parent's concatenated blocks
__global__ void k_hss_desc_b(const HNode* __restrict__ nodes, const int* __restrict__ list, int nlist,
                             const HFront* __restrict__ fronts, const int* __restrict__ ints, GemmDesc* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nlist) return;
  const HNode nd = nodes[list[t]];
  const HFront fr = fronts[nd.front];
  for (int s = 0; s < 2; ++s) {
    GemmDesc d{};
    d.sign = +1;
    d.lda = d.ldc = fr.kld;
    if (nd.parent >= 0) {
      const HNode p = nodes[nd.parent];
      if (p.parent >= 0) {   // the root needs no samples, only the skeleton index sets
        const long long off = (long long)(nd.isright ? p.capl : 0) * fr.kld;
        d.a = nd.xcat[s]; d.b = nd.ec[s]; d.ldb = nd.lde; d.c = p.xcat[s] + off;
        d.M = fr.k; d.N = ints[nd.rk + s]; d.K = nd.mcat;
      }
    }
    out[2 * t + s] = d;
  }
}

// ---- 4. row interpolative decomposition by Householder QR with column pivoting ------------------------------------------
// One CTA per (HSS node, side).  Y = the node's sample block, transposed: k × mcat, one column per (candidate skeleton) row
// of S.  QR with column pivoting of a scratch copy, truncated by pqrfact's rule; T = R11⁻¹·R12; interpolation matrix
// E[skel] = I, E[rest] = Tᵀ (= the reference's X[p[r:]] = Tᴴ for the un-conjugated transposed block, see header).
// Column norms are recomputed exactly while the reflector is applied (no downdating).  Outputs: rank, E, conj(E), and —
// into the PARENT's concatenated blocks — the skeleton index list and the skeleton columns of Y.
template <typename T>
__global__ void __launch_bounds__(256) k_hss_qrcp(const HNode* __restrict__ nodes, const int* __restrict__ list,
                                                   const HFront* __restrict__ fronts, T* __restrict__ pool, int* __restrict__ ints,
                                                   double atol, double rtol, int kcap, int mcap, int wcap) {
  const HNode nd = nodes[list[blockIdx.x]];
  if (nd.parent < 0) return;
  const int s = blockIdx.y;
  const HFront fr = fronts[nd.front];
  const int k = fr.k, mc = nd.mcat;
  const T* Y = pool + nd.ycat[s];
  extern __shared__ __align__(16) unsigned char sm_q[];
  T* u = reinterpret_cast<T*>(sm_q);
  double* nrm = reinterpret_cast<double*>(sm_q + (size_t)kcap * sizeof(T));
  int* cidx = reinterpret_cast<int*>(nrm + mcap);
  // the scratch copy the reflectors work on lives in shared memory whenever it fits (leaves and the lower HSS levels: a few
  // tens of KB) — every step of the column loop is then two shared-memory passes instead of two L2 round trips
  T* smW = reinterpret_cast<T*>(sm_q + (((size_t)kcap * sizeof(T) + (size_t)mcap * (sizeof(double) + sizeof(int)) + 15) & ~(size_t)15));
  const bool insm = (long long)fr.kld * mc <= (long long)wcap;
  T* W = insm ? smW : pool + nd.wq[s];
  const long long kld = fr.kld;
  const long long yld = fr.kld;
  __shared__ double s_v[8];
  __shared__ int s_i[8];
  __shared__ int s_rank;
  __shared__ double s_r11;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = warp; c < mc; c += 8) {
    const T* y = Y + (long long)c * yld;
    T* w = W + (long long)c * kld;
    double a = 0.0;
    for (int i = lane; i < k; i += 32) { const T v = y[i]; w[i] = v; a += habs2(v); }
    a = hwsum(a);
    if (lane == 0) { nrm[c] = a; cidx[c] = c; }
  }
  if (tid == 0) { s_rank = -1; s_r11 = 0.0; }
  __syncthreads();
  const int jmax = min(k, mc);
  int r = jmax;
  for (int j = 0; j < jmax; ++j) {
    // pivot: slot with the largest remaining norm, first one on ties
    double best = -1.0;
    int bp = 0x7fffffff;
    for (int c = j + tid; c < mc; c += 256) {
      const double v = nrm[c];
      if (v > best) { best = v; bp = c; }
    }
    warp_argmax(best, bp);
    if (lane == 0) { s_v[warp] = best; s_i[warp] = bp; }
    __syncthreads();
    best = s_v[0]; bp = s_i[0];
#pragma unroll
    for (int w8 = 1; w8 < 8; ++w8)
      if (s_v[w8] > best || (s_v[w8] == best && s_i[w8] < bp)) { best = s_v[w8]; bp = s_i[w8]; }
    const double rkk = sqrt(fmax(best, 0.0));
    const double r11 = j == 0 ? rkk : s_r11;
    const double ptol = fmax(atol, rtol * r11);
    if (!(best > 0.0) || !(rkk > ptol)) { r = j; break; }   // uniform: every thread sees the same values
    const int pc = cidx[bp], oldj = cidx[j];
    const double nj = nrm[j];
    __syncthreads();
    if (tid == 0) { cidx[j] = pc; cidx[bp] = oldj; nrm[bp] = nj; if (j == 0) s_r11 = rkk; }
    // Householder vector of the pivot column, rows j..k-1:  u = x − β·e1,  β = −phase(x0)·‖x‖,  H = I − 2·u·uᴴ/(uᴴu)
    T* xp = W + (long long)pc * kld;
    const T alpha = xp[j];
    const double am = hmag(alpha);
    const T phase = am > 0.0 ? hscal(alpha, 1.0 / am) : hs_one<T>();
    const T beta = hscal(phase, -rkk);
    for (int i = j + tid; i < k; i += 256) u[i - j] = i == j ? hs_sub(alpha, beta) : xp[i];
    const double uu = 2.0 * rkk * (rkk + am);
    __syncthreads();
    if (tid == 0) xp[j] = beta;   // R[j,j]
    const double f2 = 2.0 / uu;
    for (int c = j + 1 + warp; c < mc; c += 8) {
      T* y = W + (long long)cidx[c] * kld;
      T dot = hs_zero<T>();
      for (int i = j + lane; i < k; i += 32) dot = hcjfma(dot, u[i - j], y[i]);
      dot = hscal(hwsum(dot), f2);
      double a = 0.0;
      for (int i = j + lane; i < k; i += 32) {
        const T v = hs_fnma(y[i], u[i - j], dot);
        y[i] = v;
        if (i > j) a += habs2(v);
      }
      a = hwsum(a);
      if (lane == 0) nrm[c] = a;
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) {
    ints[nd.rk + s] = r;
    if (r >= k - 10) ints[fr.flag] = 1;   // the detected rank saturates the sample count (oracle: max(len) ≥ k − 10)
  }
  // T = R11⁻¹·R12 in place (top r entries of every non-skeleton column); one thread per column
  for (int c = r + tid; c < mc; c += 256) {
    T* t = W + (long long)cidx[c] * kld;
    for (int i = r - 1; i >= 0; --i) {
      T acc = t[i];
      for (int l = i + 1; l < r; ++l) acc = hs_fnma(acc, W[(long long)cidx[l] * kld + i], t[l]);
      t[i] = hs_mul(acc, hs_recip(W[(long long)cidx[i] * kld + i]));
    }
  }
  __syncthreads();
  T* E = pool + nd.e[s];
  T* Ec = pool + nd.ec[s];
  const long long lde = nd.lde;
  for (int q = tid; q < r; q += 256) {
    E[(long long)q * lde + cidx[q]] = hs_one<T>();
    Ec[(long long)q * lde + cidx[q]] = hs_one<T>();
  }
  for (long long e = tid; e < (long long)(mc - r) * r; e += 256) {
    const int q = (int)(e % r), c = r + (int)(e / r);
    const T v = W[(long long)cidx[c] * kld + q];
    E[(long long)q * lde + cidx[c]] = v;
    Ec[(long long)q * lde + cidx[c]] = hconj(v);
  }
  // skeleton index list and skeleton sample columns into the parent's concatenated blocks
  const HNode p = nodes[nd.parent];
  const int off = nd.isright ? p.capl : 0;
  for (int q = tid; q < r; q += 256) {
    const int c = cidx[q];
    ints[p.icat[s] + off + q] = nd.leaf ? nd.lo + c : ints[nd.icat[s] + c];
  }
  if (p.parent >= 0) {
    T* Yp = pool + p.ycat[s] + (long long)off * kld;
    for (int q = warp; q < r; q += 8) {
      const T* y = Y + (long long)cidx[q] * yld;
      T* yp = Yp + (long long)q * yld;
      for (int i = lane; i < k; i += 32) yp[i] = y[i];
    }
  }
}

// Warp-per-problem variant of the interpolative decomposition for the SMALL sample blocks — leaves and the lower HSS levels,
// where a block is a few dozen columns by a few dozen samples and there are tens of thousands of them per tree level: one
// warp per (HSS node, side), 4 per CTA, the scratch copy in shared memory with an odd pitch (lane l owns columns l, l+32,
// l+64, l+96: conflict-free), no block-wide barrier anywhere — the 256-thread kernel above spends most of its time in the
// three __syncthreads per pivot step.  No column swaps: a `done` mask marks the pivots, the pivot order is recorded.
// Same arithmetic, same truncation rule, same outputs as k_hss_qrcp.
constexpr int QW_WARPS = 4;
template <typename T>
__global__ void __launch_bounds__(QW_WARPS * 32) k_hss_qrcp_warp(const HNode* __restrict__ nodes, const int* __restrict__ list, int nitems,
                                                                  const HFront* __restrict__ fronts, T* __restrict__ pool,
                                                                  int* __restrict__ ints, double atol, double rtol, int wcapw, int kcap, int mcap) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * QW_WARPS + warp;
  if (item >= nitems) return;
  const HNode nd = nodes[list[item >> 1]];
  if (nd.parent < 0) return;
  const int s = item & 1;
  const HFront fr = fronts[nd.front];
  const int k = fr.k, mc = nd.mcat;
  const long long yld = fr.kld;
  const int pitch = k | 1;
  const T* Y = pool + nd.ycat[s];
  extern __shared__ __align__(16) unsigned char sm_w[];
  const size_t per_warp = (((size_t)(wcapw + kcap) * sizeof(T) + (size_t)mcap * sizeof(int) + 15) & ~(size_t)15);
  T* W = reinterpret_cast<T*>(sm_w + (size_t)warp * per_warp);
  T* u = W + wcapw;
  int* piv = reinterpret_cast<int*>(u + kcap);
  // scratch copy (coalesced over the rows of a column) and column norms (each lane its own columns)
  for (int c = 0; c < mc; ++c)
    for (int i = lane; i < k; i += 32) W[c * pitch + i] = Y[(long long)c * yld + i];
  __syncwarp();
  constexpr int CPL = 4;   // columns per lane (mc ≤ 128)
  double nrm[CPL];
  bool done[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const int c = lane + 32 * q;
    double a = 0.0;
    if (c < mc)
      for (int i = 0; i < k; ++i) a += habs2(W[c * pitch + i]);
    nrm[q] = c < mc ? a : -1.0;
    done[q] = c >= mc;
  }
  const int jmax = min(k, mc);
  int r = jmax;
  double r11 = 0.0;
  for (int j = 0; j < jmax; ++j) {
    double best = -1.0;
    int bp = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < CPL; ++q)
      if (!done[q] && nrm[q] > best) { best = nrm[q]; bp = lane + 32 * q; }
    warp_argmax(best, bp);
    const double rkk = sqrt(fmax(best, 0.0));
    if (j == 0) r11 = rkk;
    const double ptol = fmax(atol, rtol * r11);
    if (!(best > 0.0) || !(rkk > ptol)) { r = j; break; }
    const int pc = bp;
    const T alpha = W[pc * pitch + j];
    const double am = hmag(alpha);
    const T phase = am > 0.0 ? hscal(alpha, 1.0 / am) : hs_one<T>();
    const T beta = hscal(phase, -rkk);
    for (int i = j + lane; i < k; i += 32) u[i - j] = i == j ? hs_sub(alpha, beta) : W[pc * pitch + i];
    __syncwarp();
    const double f2 = 2.0 / (2.0 * rkk * (rkk + am));
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const int c = lane + 32 * q;
      if (c == pc) { done[q] = true; nrm[q] = -1.0; }
      if (done[q]) continue;
      T* y = W + c * pitch;
      T dot = hs_zero<T>();
      for (int i = j; i < k; ++i) dot = hcjfma(dot, u[i - j], y[i]);
      dot = hscal(dot, f2);
      double a = 0.0;
      for (int i = j; i < k; ++i) {
        const T v = hs_fnma(y[i], u[i - j], dot);
        y[i] = v;
        if (i > j) a += habs2(v);
      }
      nrm[q] = a;
    }
    if (lane == 0) { piv[j] = pc; W[pc * pitch + j] = beta; }   // R[j, j]
    __syncwarp();
  }
  __syncwarp();
  if (lane == 0) {
    ints[nd.rk + s] = r;
    if (r >= k - 10) ints[fr.flag] = 1;
  }
  // T = R11⁻¹·R12 in place, each lane its own non-pivot columns
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const int c = lane + 32 * q;
    if (c >= mc) continue;
    bool isp = false;
    for (int l = 0; l < r; ++l) isp = isp || piv[l] == c;
    done[q] = isp;
    if (isp) continue;
    T* t = W + c * pitch;
    for (int i = r - 1; i >= 0; --i) {
      T acc = t[i];
      for (int l = i + 1; l < r; ++l) acc = hs_fnma(acc, W[piv[l] * pitch + i], t[l]);
      t[i] = hs_mul(acc, hs_recip(W[piv[i] * pitch + i]));
    }
  }
  __syncwarp();
  T* E = pool + nd.e[s];
  T* Ec = pool + nd.ec[s];
  const long long lde = nd.lde;
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const int c = lane + 32 * q;
    if (c >= mc || done[q]) continue;
    for (int l = 0; l < r; ++l) {
      const T v = W[c * pitch + l];
      E[(long long)l * lde + c] = v;
      Ec[(long long)l * lde + c] = hconj(v);
    }
  }
  for (int l = lane; l < r; l += 32) {
    E[(long long)l * lde + piv[l]] = hs_one<T>();
    Ec[(long long)l * lde + piv[l]] = hs_one<T>();
  }
  const HNode p = nodes[nd.parent];
  const int off = nd.isright ? p.capl : 0;
  for (int l = lane; l < r; l += 32) {
    const int c = piv[l];
    ints[p.icat[s] + off + l] = nd.leaf ? nd.lo + c : ints[nd.icat[s] + c];
  }
  if (p.parent >= 0) {
    T* Yp = pool + p.ycat[s] + (long long)off * yld;
    for (int l = 0; l < r; ++l) {
      const T* y = Y + (long long)piv[l] * yld;
      T* yp = Yp + (long long)l * yld;
      for (int i = lane; i < k; i += 32) yp[i] = y[i];
    }
  }
}

// ---- generic batched block copy (compaction of the generators, diagonal blocks of the expansion) ----------------------
template <typename T>
__global__ void __launch_bounds__(256) k_copy_desc(const CopyDesc* __restrict__ descs, T* __restrict__ pool) {
  const CopyDesc d = descs[blockIdx.x];
  const T* src = pool + d.src;
  T* dst = pool + d.dst;
  const long long total = (long long)d.rows * d.cols;
  for (long long e = (long long)blockIdx.y * 256 + threadIdx.x; e < total; e += (long long)gridDim.y * 256) {
    const int i = (int)(e % d.rows), j = (int)(e / d.rows);
    const int si = i < d.gap_at ? i : i + d.gap_skip;
    const T v = src[(long long)j * d.lds + si];
    if (d.mode == 0) dst[(long long)j * d.ldd + i] = v;
    else dst[(long long)i * d.ldd + j] = d.mode == 1 ? hconj(v) : v;
  }
}
// slot[perm[i], perm[j]] = P[i, j]: the dense matrix the HSS form represents goes back into the front's S block
struct ScatterDesc { long long src, dst; int lds, ldd, m, perm; };
template <typename T>
__global__ void __launch_bounds__(256) k_hss_scatter(const ScatterDesc* __restrict__ descs, T* __restrict__ pool, const int* __restrict__ hperm) {
  const ScatterDesc d = descs[blockIdx.x];
  const int* perm = hperm + d.perm;
  const T* src = pool + d.src;
  T* dst = pool + d.dst;
  for (int j = blockIdx.y; j < d.m; j += gridDim.y) {
    const long long cj = (long long)perm[j] * d.ldd;
    for (int i = threadIdx.x; i < d.m; i += 256) dst[cj + perm[i]] = src[(long long)j * d.lds + i];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host-side wall-clock breakdown of the construction (HS_PLAN_TIMING=1): the stream is synchronised at every tick
struct HssClock {
  bool on = getenv("HS_PLAN_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  void start(cudaStream_t st) { if (on) { cudaStreamSynchronize(st); t0 = std::chrono::steady_clock::now(); } }
  void tick(int slot, cudaStream_t st) {
    if (!on) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    acc[slot] += std::chrono::duration<double, std::milli>(t1 - t0).count();
    t0 = t1;
  }
  void report() {
    if (!on) return;
    fprintf(stderr, "[hss] make_round %.1f  sketch %.1f  entries+desc %.1f  node gemms %.1f  qrcp %.1f  readback %.1f  store %.1f  expand %.1f ms\n",
            acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], acc[7]);
    for (double& a : acc) a = 0;
  }
};
static HssClock g_hclk;

struct Bump {
  long long off = 0;
  long long take(long long n) { const long long o = off; off += up32(std::max<long long>(n, 1)); return o; }
};

// recycled device scratch (hs_fac::dev_cache): smallest cached block that fits and is not more than 4x too large
void* cache_take(hs_fac* f, size_t bytes, size_t* got) {
  bytes = std::max<size_t>(bytes, 256);
  int best = -1;
  for (int i = 0; i < (int)f->dev_cache.size(); ++i) {
    const size_t c = f->dev_cache[i].second;
    if (c >= bytes && c <= 4 * bytes + (1u << 20) && (best < 0 || c < f->dev_cache[best].second)) best = i;
  }
  if (best >= 0) {
    void* p = f->dev_cache[best].first;
    *got = f->dev_cache[best].second;
    f->dev_cache.erase(f->dev_cache.begin() + best);
    return p;
  }
  const size_t want = bytes + bytes / 4;   // head room: the next round / level usually asks for a little more
  void* p = nullptr;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    // out of memory with blocks of the wrong size parked in the cache: release them and try once more
    cudaGetLastError();
    cudaStreamSynchronize(f->ctx->stream);
    for (auto& b : f->dev_cache) cudaFree(b.first);
    f->dev_cache.clear();
    CUDA_OK(cudaMalloc(&p, want));
  }
  *got = want;
  return p;
}
void cache_give(hs_fac* f, void* p, size_t bytes) { if (p) f->dev_cache.emplace_back(p, bytes); }

template <typename V> struct DevVec {   // small helper: device copy of a host vector (scratch from the cache)
  hs_fac* f;
  V* d = nullptr;
  size_t bytes = 0;
  explicit DevVec(hs_fac* f_) : f(f_) {}
  DevVec(const DevVec&) = delete;
  ~DevVec() { cache_give(f, d, bytes); }
  void upload(const std::vector<V>& h, cudaStream_t st) {
    if (h.size() * sizeof(V) > bytes) { cache_give(f, d, bytes); d = nullptr; bytes = 0; d = (V*)cache_take(f, h.size() * sizeof(V), &bytes); }
    if (!h.empty()) CUDA_OK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice, st));
  }
};

struct Round {
  hs_fac* f = nullptr;
  void* ws = nullptr;
  size_t ws_bytes = 0, ints_bytes = 0;
  int* ints = nullptr;
  int nints = 0;
  std::vector<int> hidx;     // index into f->hss of every front of this round
  std::vector<int> nbase;    // first HNode of every front
  std::vector<HFront> fr;
  std::vector<HNode> nd;
  std::vector<int> h_ints;   // ranks / flags read back
  ~Round() { if (f) { cache_give(f, ws, ws_bytes); cache_give(f, ints, ints_bytes); } }
};

template <typename T> long long rel(const hs_fac* f, const void* p) {
  return (long long)(((const char*)p - (const char*)f->pool) / (long long)sizeof(T));
}

// sketch matrices on the device: host-supplied (uploaded by hs_hss_plan) or generated, with at least `kneed` columns
template <typename T> void ensure_sketch(hs_fac* f, long long rows_need, long long kneed) {
  if (f->opts.sketch_omega) {
    if (kneed > f->sk_cols || rows_need > f->sk_rows)
      throw hs_error(HS_EARG, "hs_factor: the supplied sketch matrices are " + std::to_string(f->sk_rows) + "x" + std::to_string(f->sk_cols) +
                                  ", the randomized HSS construction needs " + std::to_string(rows_need) + "x" + std::to_string(kneed));
    return;
  }
  if (kneed <= f->sk_cols) return;
  cudaStream_t st = f->ctx->stream;
  const long long cols = std::max<long long>(kneed + 32, f->sk_cols * 3 / 2);
  T* nb = nullptr;
  CUDA_OK(cudaMalloc((void**)&nb, (size_t)(2 * f->sk_rows * cols) * sizeof(T)));
  // regenerate everything (values depend only on (seed, matrix, row, column))
  const long long total = 2 * f->sk_rows * cols;
  k_sk_generate<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(nb, f->sk_rows, cols, 0, f->opts.sketch_seed);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(f->d_sk);
  f->d_sk = nb;
  f->sk_cols = cols;
  ++f->stats.launches_factor;
}

template <typename T> Round* make_round(hs_fac* f, const std::vector<int>& act) {
  std::unique_ptr<Round> R(new Round());
  R->f = f;
  Bump B;
  int nints = 0;
  auto take_int = [&](int n) { const int o = nints; nints += std::max(n, 1); return o; };
  for (int hi : act) {
    const HssFront& H = f->hss[hi];
    const CompFront& cf = f->comp[H.comp];
    const Front& fd = f->fronts[cf.fi];
    HFront F{};
    F.nb = cf.nb; F.r2 = cf.r2; F.m = H.m; F.k = H.k;
    F.ldf = fd.ld; F.ldz = even_up(cf.nb); F.ldri = cf.ri_ld;
    F.nbld = even_up(cf.nb); F.kld = even_up(H.k); F.r2ld = even_up(std::max(cf.r2, 1));
    F.abb = fd.off + (long long)cf.ni * fd.ld + cf.ni;
    F.z = cf.zoff;
    F.ri = cf.ri;
    F.omz = B.take((long long)F.nbld * F.k); F.psh = B.take((long long)F.kld * F.nb);
    F.yu = B.take((long long)F.nbld * F.k);  F.cp = B.take((long long)F.kld * F.nb);
    F.t1 = B.take((long long)F.r2ld * F.k);  F.t2 = B.take((long long)F.kld * std::max(cf.r2, 1));
    F.gst = B.take((long long)F.kld * F.m);  F.gsct = B.take((long long)F.kld * F.m);
    F.gomt = B.take((long long)F.kld * F.m); F.gpst = B.take((long long)F.kld * F.m);
    F.flag = take_int(1);
    F.perm = H.perm_off;
    const int fi = (int)R->fr.size();
    R->hidx.push_back(hi);
    const int base = (int)R->nd.size();
    R->nbase.push_back(base);
    const int nn = (int)H.tree.size();
    R->nd.resize(base + nn);
    // caps bottom-up (children have larger pre-order numbers than their parent)
    for (int t = nn - 1; t >= 0; --t) {
      const HssTreeNode& tn = H.tree[t];
      HNode& N = R->nd[base + t];
      N = HNode{};
      N.front = fi; N.lo = tn.lo; N.hi = tn.hi; N.isright = tn.isright; N.leaf = tn.left < 0;
      N.left = tn.left < 0 ? -1 : base + tn.left; N.right = tn.right < 0 ? -1 : base + tn.right;
      N.parent = tn.parent < 0 ? -1 : base + tn.parent;
      if (N.leaf) { N.mcat = tn.hi - tn.lo; N.capl = 0; }
      else { N.capl = R->nd[N.left].cap; N.mcat = N.capl + R->nd[N.right].cap; }
      N.cap = std::min(N.mcat, H.k);
      N.rk = take_int(2);
    }
    for (int t = 0; t < nn; ++t) {
      HNode& N = R->nd[base + t];
      const int m = N.hi - N.lo;
      if (N.leaf) {
        N.ycat[0] = F.gst + (long long)N.lo * F.kld;  N.ycat[1] = F.gsct + (long long)N.lo * F.kld;
        N.xcat[0] = F.gpst + (long long)N.lo * F.kld; N.xcat[1] = F.gomt + (long long)N.lo * F.kld;
        N.ldm0 = even_up(m); N.ldmt = even_up(m);
        N.m0 = B.take((long long)N.ldm0 * m); N.mt = B.take((long long)N.ldmt * m); N.mc = B.take((long long)N.ldm0 * m);
        N.icat[0] = N.icat[1] = 0;
      } else {
        const int capr = N.mcat - N.capl;
        for (int s = 0; s < 2; ++s) {
          N.ycat[s] = B.take((long long)F.kld * N.mcat);
          N.xcat[s] = B.take((long long)F.kld * N.mcat);
          N.icat[s] = take_int(N.mcat);
        }
        N.ldm0 = even_up(N.capl); N.ldmt = even_up(capr); N.ldm1 = even_up(capr); N.ldm1t = even_up(N.capl);
        N.m0 = B.take((long long)N.ldm0 * capr); N.mt = B.take((long long)N.ldmt * N.capl); N.mc = B.take((long long)N.ldm0 * capr);
        N.m1 = B.take((long long)N.ldm1 * N.capl); N.m1t = B.take((long long)N.ldm1t * capr); N.m1c = B.take((long long)N.ldm1 * N.capl);
      }
      if (N.parent >= 0) {
        N.lde = even_up(N.mcat);
        for (int s = 0; s < 2; ++s) {
          N.wq[s] = B.take((long long)F.kld * N.mcat);
          N.e[s] = B.take((long long)N.lde * N.cap);
          N.ec[s] = B.take((long long)N.lde * N.cap);
        }
      }
    }
    R->fr.push_back(F);
  }
  cudaStream_t st = f->ctx->stream;
  R->ws = cache_take(f, (size_t)std::max<long long>(B.off, 32) * sizeof(T), &R->ws_bytes);
  R->ints = (int*)cache_take(f, (size_t)std::max(nints, 1) * sizeof(int), &R->ints_bytes);
  R->nints = nints;
  CUDA_OK(cudaMemsetAsync(R->ws, 0, (size_t)std::max<long long>(B.off, 32) * sizeof(T), st));
  CUDA_OK(cudaMemsetAsync(R->ints, 0, (size_t)std::max(nints, 1) * sizeof(int), st));
  const long long base = rel<T>(f, R->ws);
  for (HFront& F : R->fr) {
    F.omz += base; F.psh += base; F.yu += base; F.cp += base; F.t1 += base; F.t2 += base;
    F.gst += base; F.gsct += base; F.gomt += base; F.gpst += base;
  }
  // leaves point into their front's sample matrices (already relocated through F above — redo them), the rest is relocated here
  for (size_t q = 0; q < R->fr.size(); ++q) {
    const HFront& F = R->fr[q];
    const int b0 = R->nbase[q], b1 = q + 1 < R->fr.size() ? R->nbase[q + 1] : (int)R->nd.size();
    for (int t = b0; t < b1; ++t) {
      HNode& N = R->nd[t];
      if (N.leaf) {
        N.ycat[0] = F.gst + (long long)N.lo * F.kld;  N.ycat[1] = F.gsct + (long long)N.lo * F.kld;
        N.xcat[0] = F.gpst + (long long)N.lo * F.kld; N.xcat[1] = F.gomt + (long long)N.lo * F.kld;
      } else {
        for (int s = 0; s < 2; ++s) { N.ycat[s] += base; N.xcat[s] += base; }
        N.m1 += base; N.m1t += base; N.m1c += base;
      }
      N.m0 += base; N.mt += base; N.mc += base;
      if (N.parent >= 0) for (int s = 0; s < 2; ++s) { N.wq[s] += base; N.e[s] += base; N.ec[s] += base; }
    }
  }
  return R.release();
}

template <typename T> void run_round(hs_fac* f, Round& R) {
  cudaStream_t st = f->ctx->stream;
  const int nf = (int)R.fr.size();
  int max_m = 0, max_k = 0, max_nb = 0, max_r2 = 0, max_height = 0, max_mcat = 0, max_cap = 0;
  long long rows_need = 0;
  for (size_t q = 0; q < R.fr.size(); ++q) {
    const HFront& F = R.fr[q];
    max_m = std::max(max_m, F.m); max_k = std::max(max_k, F.k); max_nb = std::max(max_nb, F.nb); max_r2 = std::max(max_r2, F.r2);
    rows_need = std::max<long long>(rows_need, F.m);
    for (const HssTreeNode& tn : f->hss[R.hidx[q]].tree) max_height = std::max(max_height, tn.height);
  }
  for (const HNode& N : R.nd) { max_mcat = std::max(max_mcat, N.mcat); max_cap = std::max(max_cap, N.cap); }
  ensure_sketch<T>(f, rows_need, max_k);
  g_hclk.start(st);
  DevVec<HFront> dfr(f);
  DevVec<HNode> dnd(f);
  dfr.upload(R.fr, st);
  dnd.upload(R.nd, st);
  T* pool = (T*)f->pool;
  const T* om = (const T*)f->d_sk;
  const T* ps = om + f->sk_rows * f->sk_cols;
  // ---- sketches (factorization.jl:239-244), matrix-free --------------------------------------------------------------
  k_sk_fill<T><<<dim3(nf, (max_m + 255) / 256), 256, 0, st>>>(dfr.d, pool, f->d_hperm, om, ps, f->sk_rows);
  CUDA_OK(cudaGetLastError());
  ++f->stats.launches_factor;
  {
    std::vector<GemmDesc> gd;
    auto item = [&](long long a, int lda, long long b, int ldb, long long c, int ldc, int M, int N, int K, int sign) {
      GemmDesc d{};
      d.a = a; d.b = b; d.c = c; d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.M = M; d.N = N; d.K = K; d.sign = sign;
      gd.push_back(d);
    };
    // stage 1: T1 = Ri·Ωz, Yu = Abb·Ωz, T2 = Ψzᴴ·Z, Cp = Ψzᴴ·Abb      stage 2: Yu −= Z·T1, Cp −= T2·Ri
    for (const HFront& F : R.fr) {
      item(F.ri, F.ldri, F.omz, F.nbld, F.t1, F.r2ld, F.r2, F.k, F.nb, +1);
      item(F.abb, F.ldf, F.omz, F.nbld, F.yu, F.nbld, F.nb, F.k, F.nb, +1);
      item(F.psh, F.kld, F.z, F.ldz, F.t2, F.kld, F.k, F.r2, F.nb, +1);
      item(F.psh, F.kld, F.abb, F.ldf, F.cp, F.kld, F.k, F.nb, F.nb, +1);
    }
    const int n1 = (int)gd.size();
    for (const HFront& F : R.fr) {
      item(F.z, F.ldz, F.t1, F.r2ld, F.yu, F.nbld, F.nb, F.k, F.r2, -1);
      item(F.t2, F.kld, F.ri, F.ldri, F.cp, F.kld, F.k, F.nb, F.r2, -1);
    }
    const double cx = f->dtype == HS_C64 ? 4.0 : 1.0;
    for (const GemmDesc& d : gd) f->stats.sketch_flops += cx * 2.0 * d.M * (double)d.N * d.K;
    DevVec<GemmDesc> dgd(f);
    dgd.upload(gd, st);
    const int mm = std::max(max_nb, max_k), nn = std::max(max_nb, max_k);
    hs_gen_gemm(f, dgd.d, n1, mm, nn);
    hs_gen_gemm(f, dgd.d + n1, (int)gd.size() - n1, mm, nn);
    CUDA_OK(cudaStreamSynchronize(st));   // gd / dgd go out of scope
  }
  k_sk_gather<T><<<dim3(nf, (max_m + 255) / 256), 256, 0, st>>>(dfr.d, pool, f->d_hperm);
  CUDA_OK(cudaGetLastError());
  ++f->stats.launches_factor;
  g_hclk.tick(1, st);
  // ---- HSS nodes by height ------------------------------------------------------------------------------------------------
  std::vector<std::vector<int>> byh(max_height + 1);
  for (size_t q = 0; q < R.fr.size(); ++q) {
    const HssFront& H = f->hss[R.hidx[q]];
    for (size_t t = 0; t < H.tree.size(); ++t) byh[H.tree[t].height].push_back(R.nbase[q] + (int)t);
  }
  std::vector<int> flat, off(max_height + 2, 0);
  for (int h = 0; h <= max_height; ++h) { off[h] = (int)flat.size(); flat.insert(flat.end(), byh[h].begin(), byh[h].end()); }
  off[max_height + 1] = (int)flat.size();
  DevVec<int> dlist(f);
  dlist.upload(flat, st);
  size_t maxl = 1;
  for (auto& v : byh) maxl = std::max(maxl, v.size());
  size_t dg_bytes = 0;
  GemmDesc* dg = (GemmDesc*)cache_take(f, maxl * 4 * sizeof(GemmDesc), &dg_bytes);
  struct GiveBack { hs_fac* f; void* p; size_t b; ~GiveBack() { cache_give(f, p, b); } } dg_guard{f, dg, dg_bytes};
  const int kcap = even_up(max_k), mcap = max_mcat + 2;
  const size_t smem0 = (((size_t)kcap * sizeof(T) + (size_t)mcap * (sizeof(double) + sizeof(int)) + 15) & ~(size_t)15);
  if (smem0 > 200 * 1024) { throw hs_error(HS_ESIZE, "randomized HSS construction: sample count / block size exceed the shared-memory budget of the pivoted QR"); }
  const double atol = f->opts.atol, rtol = f->opts.rtol;   // factorization.jl:110 passes atol, rtol unhalved
  {
    for (int h = 0; h <= max_height; ++h) {
      const int nl = (int)byh[h].size();
      if (nl == 0) continue;
      const int* lst = dlist.d + off[h];
      k_hss_entries<T><<<dim3(nl, 2), 256, 0, st>>>(dnd.d, lst, dfr.d, pool, R.ints, f->d_hperm);
      k_hss_desc_a<<<(nl + 127) / 128, 128, 0, st>>>(dnd.d, lst, nl, dfr.d, R.ints, dg);
      CUDA_OK(cudaGetLastError());
      f->stats.launches_factor += 2;
      g_hclk.tick(2, st);
      hs_gen_gemm(f, dg, 4 * nl, max_k, std::max(max_cap, max_mcat));
      g_hclk.tick(3, st);
      {
        // small sample blocks (≤ 128 columns, scratch ≤ 40 KB) go to the warp-per-problem kernel, 4 per CTA; the rest to
        // the CTA-per-problem kernel with the scratch in shared memory whenever it fits next to the bookkeeping
        const long long budget_w = (long long)(40 * 1024) / (long long)sizeof(T);
        std::vector<int> small, large;
        long long wmax = 0, wmax_w = 0;
        int kmax_w = 2, mmax_w = 2;
        for (int t : byh[h]) {
          const HNode& N = R.nd[t];
          if (N.parent < 0) continue;
          const int kk = R.fr[N.front].k;
          const long long need_w = (long long)(kk | 1) * N.mcat;
          if (N.mcat <= 128 && need_w + kk <= budget_w) {
            small.push_back(t); wmax_w = std::max(wmax_w, need_w); kmax_w = std::max(kmax_w, kk); mmax_w = std::max(mmax_w, N.mcat);
          } else {
            large.push_back(t); wmax = std::max(wmax, (long long)R.fr[N.front].kld * N.mcat);
          }
        }
        std::vector<int> both(small);
        both.insert(both.end(), large.begin(), large.end());
        DevVec<int> dsplit(f);
        dsplit.upload(both, st);
        if (!small.empty()) {
          const int wcapw = (int)((wmax_w + 1) & ~1ll), kcw = even_up(kmax_w), mcw = (mmax_w + 3) & ~3;
          const size_t per_warp = (((size_t)(wcapw + kcw) * sizeof(T) + (size_t)mcw * sizeof(int) + 15) & ~(size_t)15);
          const int nitems = 2 * (int)small.size();
          k_hss_qrcp_warp<T><<<(nitems + QW_WARPS - 1) / QW_WARPS, QW_WARPS * 32, per_warp * QW_WARPS, st>>>(
              dnd.d, dsplit.d, nitems, dfr.d, pool, R.ints, atol, rtol, wcapw, kcw, mcw);
        }
        if (!large.empty()) {
          const long long budget = ((long long)200 * 1024 - (long long)smem0) / (long long)sizeof(T);
          const int wcap = (int)std::max<long long>(0, std::min(wmax, budget));
          const size_t smem = smem0 + (size_t)wcap * sizeof(T);
          k_hss_qrcp<T><<<dim3((unsigned)large.size(), 2), 256, smem, st>>>(dnd.d, dsplit.d + small.size(), dfr.d, pool, R.ints, atol, rtol, kcap, mcap, wcap);
        }
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaStreamSynchronize(st));   // `dsplit` goes back to the cache
      }
      g_hclk.tick(4, st);
      k_hss_desc_b<<<(nl + 127) / 128, 128, 0, st>>>(dnd.d, lst, nl, dfr.d, R.ints, dg);
      CUDA_OK(cudaGetLastError());
      f->stats.launches_factor += 2;
      hs_gen_gemm(f, dg, 2 * nl, max_k, max_cap);
      g_hclk.tick(3, st);
    }
    R.h_ints.resize(std::max(R.nints, 1));
    CUDA_OK(cudaMemcpyAsync(R.h_ints.data(), R.ints, (size_t)std::max(R.nints, 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    g_hclk.tick(5, st);
  }
}

// ---- expansion: the dense matrix a stored HSS form represents ----------------------------------------------------------
// to_slot: write it (un-permuted) into the dense slot of the front and keep the parent's generator blocks in the store;
// otherwise P (m×m, S[perm,perm] order) of the single front `hids[0]` is copied to `host_out`.
template <typename T> void expand(hs_fac* f, const std::vector<int>& hids, bool to_slot, T* host_out) {
  if (hids.empty()) return;
  cudaStream_t st = f->ctx->stream;
  Bump B;
  // assembled (nested) bases of every node: Û (m×r0) and V̂ᴴ (r1×m); `wsU` / `wsV`: offset is into the workspace (not yet relocated)
  struct Asm { long long U = 0, VH = 0; int ldU = 0, ldVH = 0; bool wsU = false, wsV = false, have = false; };
  struct TT { long long a = 0, b = 0; int lda = 0, ldb = 0; bool ws = false; };   // Ta = Û_a·B12, Tb = Û_b·B21 of a branch
  std::vector<std::vector<Asm>> as(hids.size());
  std::vector<std::vector<TT>> tt(hids.size());
  std::vector<long long> Poff(hids.size());
  int max_height = 0, max_m = 0;
  for (size_t q = 0; q < hids.size(); ++q) {
    const HssFront& H = f->hss[hids[q]];
    const int nn = (int)H.tree.size();
    as[q].resize(nn); tt[q].resize(nn);
    Poff[q] = B.take((long long)even_up(H.m) * H.m);
    max_m = std::max(max_m, H.m);
    for (int t = 0; t < nn; ++t) {
      const HssTreeNode& tn = H.tree[t];
      const HssStored& S = H.st[t];
      max_height = std::max(max_height, tn.height);
      const int m = tn.hi - tn.lo;
      Asm& A = as[q][t];
      if (tn.left < 0) {
        A.U = H.sbase + S.U; A.ldU = S.ldU; A.VH = H.sbase + S.VH; A.ldVH = S.ldVH; A.have = true;
      } else if (tn.parent >= 0) {
        A.have = true;
        A.ldU = even_up(m); A.U = B.take((long long)A.ldU * std::max(S.r0, 1)); A.wsU = true;
        // the assembled column bases of the two depth-1 nodes are what the parent's Gauss transforms read: they live in the store
        if (to_slot && tn.parent == 0) { A.VH = H.sbase + (tn.isright ? H.vhb : H.vha); A.ldVH = tn.isright ? H.ld_vhb : H.ld_vha; }
        else { A.ldVH = even_up(std::max(S.r1, 1)); A.VH = B.take((long long)A.ldVH * m); A.wsV = true; }
      }
      if (tn.left >= 0) {
        const HssTreeNode &a = H.tree[tn.left], &b = H.tree[tn.right];
        const HssStored &Sa = H.st[tn.left], &Sb = H.st[tn.right];
        TT& X = tt[q][t];
        if (to_slot && t == 0) { X.a = H.sbase + H.ta; X.lda = H.ld_ta; X.b = H.sbase + H.tb; X.ldb = H.ld_tb; }   // kept for the parent front
        else {
          X.ws = true;
          X.lda = even_up(a.hi - a.lo); X.a = B.take((long long)X.lda * std::max(Sb.r1, 1));
          X.ldb = even_up(b.hi - b.lo); X.b = B.take((long long)X.ldb * std::max(Sa.r1, 1));
        }
      }
    }
  }
  size_t ws_bytes = 0;
  void* ws = cache_take(f, (size_t)std::max<long long>(B.off, 32) * sizeof(T), &ws_bytes);
  struct GiveBack { hs_fac* f; void* p; size_t b; ~GiveBack() { cache_give(f, p, b); } } fr_{f, ws, ws_bytes};
  CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)std::max<long long>(B.off, 32) * sizeof(T), st));
  const long long base = rel<T>(f, ws);
  std::vector<CopyDesc> cds;
  std::vector<std::vector<GemmDesc>> gh(max_height + 1);
  std::vector<GemmDesc> gT, gP;
  auto item = [&](std::vector<GemmDesc>& v, long long a, int lda, long long b, int ldb, long long c, int ldc, int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return;
    GemmDesc d{};
    d.a = a; d.b = b; d.c = c; d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.M = M; d.N = N; d.K = K; d.sign = +1;
    v.push_back(d);
  };
  std::vector<ScatterDesc> sds;
  for (size_t q = 0; q < hids.size(); ++q) {
    HssFront& H = f->hss[hids[q]];
    const int nn = (int)H.tree.size();
    const long long P = base + Poff[q];
    const int ldP = even_up(H.m);
    for (int t = 0; t < nn; ++t) {   // relocate workspace offsets
      Asm& A = as[q][t];
      if (A.wsU) A.U += base;
      if (A.wsV) A.VH += base;
      if (tt[q][t].ws) { tt[q][t].a += base; tt[q][t].b += base; }
    }
    for (int t = 0; t < nn; ++t) {
      const HssTreeNode& tn = H.tree[t];
      const HssStored& S = H.st[t];
      const int m = tn.hi - tn.lo;
      if (tn.left < 0) {
        CopyDesc c{};
        c.src = H.sbase + S.D; c.lds = S.ldD; c.dst = P + (long long)tn.lo * ldP + tn.lo; c.ldd = ldP; c.rows = m; c.cols = m; c.gap_at = m; c.mode = 0;
        cds.push_back(c);
        continue;
      }
      const HssTreeNode &a = H.tree[tn.left], &b = H.tree[tn.right];
      const HssStored &Sa = H.st[tn.left], &Sb = H.st[tn.right];
      const Asm &Aa = as[q][tn.left], &Ab = as[q][tn.right];
      const int ma = a.hi - a.lo, mb = b.hi - b.lo;
      if (tn.parent >= 0) {
        // Û = [Û_a·R1; Û_b·R2],  V̂ᴴ = [W1ᴴ·V̂ᴴ_a, W2ᴴ·V̂ᴴ_b]
        const Asm& At = as[q][t];
        const long long R = H.sbase + S.R, WH = H.sbase + S.WH;
        item(gh[tn.height], Aa.U, Aa.ldU, R, S.ldR, At.U, At.ldU, ma, S.r0, Sa.r0);
        item(gh[tn.height], Ab.U, Ab.ldU, R + Sa.r0, S.ldR, At.U + ma, At.ldU, mb, S.r0, Sb.r0);
        item(gh[tn.height], WH, S.ldWH, Aa.VH, Aa.ldVH, At.VH, At.ldVH, S.r1, ma, Sa.r1);
        item(gh[tn.height], WH + (long long)Sa.r1 * S.ldWH, S.ldWH, Ab.VH, Ab.ldVH, At.VH + (long long)ma * At.ldVH, At.ldVH, S.r1, mb, Sb.r1);
      }
      // off-diagonal blocks  P[Ia, Jb] = (Û_a·B12)·V̂ᴴ_b,  P[Ib, Ja] = (Û_b·B21)·V̂ᴴ_a
      const TT& X = tt[q][t];
      item(gT, Aa.U, Aa.ldU, H.sbase + S.B12, S.ldB12, X.a, X.lda, ma, Sb.r1, Sa.r0);
      item(gT, Ab.U, Ab.ldU, H.sbase + S.B21, S.ldB21, X.b, X.ldb, mb, Sa.r1, Sb.r0);
      item(gP, X.a, X.lda, Ab.VH, Ab.ldVH, P + (long long)b.lo * ldP + a.lo, ldP, ma, mb, Sb.r1);
      item(gP, X.b, X.ldb, Aa.VH, Aa.ldVH, P + (long long)a.lo * ldP + b.lo, ldP, mb, ma, Sa.r1);
    }
    if (to_slot) {
      const CompFront& cf = f->comp[H.comp];
      const Front& fd = f->fronts[cf.fi];
      ScatterDesc s{};
      s.src = P; s.lds = ldP; s.dst = fd.off + (long long)cf.ni * fd.ld + cf.ni; s.ldd = fd.ld; s.m = H.m;
      s.perm = H.perm_off;
      sds.push_back(s);
    }
  }
  T* pool = (T*)f->pool;
  auto run_gemms = [&](std::vector<GemmDesc>& v) {
    if (v.empty()) return;
    int mM = 0, mN = 0;
    for (const GemmDesc& d : v) { mM = std::max(mM, d.M); mN = std::max(mN, d.N); }
    DevVec<GemmDesc> dv(f);
    dv.upload(v, st);
    hs_gen_gemm(f, dv.d, (int)v.size(), mM, mN);
    CUDA_OK(cudaStreamSynchronize(st));
  };
  if (!cds.empty()) {
    DevVec<CopyDesc> dc(f);
    dc.upload(cds, st);
    k_copy_desc<T><<<dim3((unsigned)cds.size(), 4), 256, 0, st>>>(dc.d, pool);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(st));
    ++f->stats.launches_factor;
  }
  for (int h = 1; h <= max_height; ++h) run_gemms(gh[h]);
  run_gemms(gT);
  run_gemms(gP);
  if (to_slot) {
    DevVec<ScatterDesc> ds(f);
    ds.upload(sds, st);
    k_hss_scatter<T><<<dim3((unsigned)sds.size(), std::min(max_m, 256)), 256, 0, st>>>(ds.d, pool, f->d_hperm);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(st));
    ++f->stats.launches_factor;
  } else {
    const HssFront& H = f->hss[hids[0]];
    CUDA_OK(cudaMemcpy2DAsync(host_out, (size_t)H.m * sizeof(T), pool + base + Poff[0], (size_t)even_up(H.m) * sizeof(T),
                              (size_t)H.m * sizeof(T), H.m, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
}

template <typename T> void build_impl(hs_fac* f, CompLevel& C) {
  cudaStream_t st = f->ctx->stream;
  std::vector<int> all, act;
  for (int c = C.c0; c < C.c1; ++c)
    if (f->comp[c].hss >= 0) all.push_back(f->comp[c].hss);
  if (all.empty()) return;
  for (int hi : all) {
    HssFront& H = f->hss[hi];
    const CompFront& cf = f->comp[H.comp];
    // kest < 0 ⇒ ceil(0.5·rank(L)) (factorization.jl:102-104); randcompress_adaptive oversamples by 10
    const long long kest = f->opts.kest < 0 ? (cf.r1 + 1) / 2 : f->opts.kest;
    H.k = (int)kest + 10;
    H.rounds = 0; H.done = false; H.hssrank = 0;
    for (HssStored& S : H.st) S = HssStored{};
  }
  act = all;
  std::vector<std::unique_ptr<Round>> rounds;
  std::vector<std::pair<int, int>> where(f->hss.size(), {-1, -1});   // (round, slot) holding each front's final data
  const int max_rounds = 12;
  while (!act.empty()) {
    g_hclk.start(st);
    rounds.emplace_back(make_round<T>(f, act));
    g_hclk.tick(0, st);
    Round& R = *rounds.back();
    run_round<T>(f, R);
    ++f->stats.hss_rounds;
    std::vector<int> next;
    for (size_t q = 0; q < R.fr.size(); ++q) {
      HssFront& H = f->hss[R.hidx[q]];
      const bool sat = R.h_ints[R.fr[q].flag] != 0;
      ++H.rounds;
      if (!sat || H.k >= H.m + 10 || H.rounds >= max_rounds) {
        H.done = true;
        where[R.hidx[q]] = {(int)rounds.size() - 1, (int)q};
        for (size_t t = 0; t < H.tree.size(); ++t) {
          const HNode& N = R.nd[R.nbase[q] + t];
          if (N.parent >= 0) { H.st[t].r0 = R.h_ints[N.rk]; H.st[t].r1 = R.h_ints[N.rk + 1]; }
        }
        if (f->opts.verbose)
          fprintf(stderr, "hsolve: HSS Schur complement of front %d: m = %d, k = %d, rounds = %d%s\n", f->comp[H.comp].fi, H.m, H.k, H.rounds,
                  sat ? " (still saturated)" : "");
      } else {
        H.k += (int)std::max<int64_t>(f->opts.stepsize, 1);
        next.push_back(R.hidx[q]);
      }
    }
    act.swap(next);
  }
  // ---- compact store -----------------------------------------------------------------------------------------------------
  g_hclk.start(st);
  Bump B;
  for (int hi : all) {
    HssFront& H = f->hss[hi];
    const int nn = (int)H.tree.size();
    for (int t = 0; t < nn; ++t) {
      const HssTreeNode& tn = H.tree[t];
      HssStored& S = H.st[t];
      const int m = tn.hi - tn.lo;
      if (tn.left < 0) {
        S.ldD = even_up(m); S.D = B.take((long long)S.ldD * m);
        S.ldU = even_up(m); S.U = B.take((long long)S.ldU * std::max(S.r0, 1));
        S.ldVH = even_up(std::max(S.r1, 1)); S.VH = B.take((long long)S.ldVH * m);
      } else {
        const HssStored &Sa = H.st[tn.left], &Sb = H.st[tn.right];
        if (tn.parent >= 0) {
          S.ldR = even_up(std::max(Sa.r0 + Sb.r0, 1)); S.R = B.take((long long)S.ldR * std::max(S.r0, 1));
          S.ldWH = even_up(std::max(S.r1, 1)); S.WH = B.take((long long)S.ldWH * std::max(Sa.r1 + Sb.r1, 1));
        }
        S.ldB12 = even_up(std::max(Sa.r0, 1)); S.B12 = B.take((long long)S.ldB12 * std::max(Sb.r1, 1));
        S.ldB21 = even_up(std::max(Sb.r0, 1)); S.B21 = B.take((long long)S.ldB21 * std::max(Sa.r1, 1));
        H.hssrank = std::max(H.hssrank, std::max(std::max(Sa.r0, Sb.r1), std::max(Sb.r0, Sa.r1)));
      }
    }
    // generator blocks for the parent front (:129-137)
    const HssTreeNode &a = H.tree[H.tree[0].left], &b = H.tree[H.tree[0].right];
    const HssStored &Sa = H.st[H.tree[0].left], &Sb = H.st[H.tree[0].right];
    const int ma = a.hi - a.lo, mb = b.hi - b.lo;
    H.ra1 = Sa.r1; H.rb1 = Sb.r1;
    H.ld_ta = even_up(ma); H.ta = B.take((long long)H.ld_ta * std::max(Sb.r1, 1));
    H.ld_tb = even_up(mb); H.tb = B.take((long long)H.ld_tb * std::max(Sa.r1, 1));
    if (a.left < 0) { H.vha = Sa.VH; H.ld_vha = Sa.ldVH; } else { H.ld_vha = even_up(std::max(Sa.r1, 1)); H.vha = B.take((long long)H.ld_vha * ma); }
    if (b.left < 0) { H.vhb = Sb.VH; H.ld_vhb = Sb.ldVH; } else { H.ld_vhb = even_up(std::max(Sb.r1, 1)); H.vhb = B.take((long long)H.ld_vhb * mb); }
    f->stats.hss_maxrank = std::max<int64_t>(f->stats.hss_maxrank, H.hssrank);
    f->stats.maxrank = std::max<int64_t>(f->stats.maxrank, H.hssrank);
  }
  const size_t need = (size_t)std::max<long long>(B.off, 32) * sizeof(T);
  if (need > C.hss_store_bytes) {
    cudaFree(C.hss_store);
    C.hss_store = nullptr; C.hss_store_bytes = 0;
    CUDA_OK(cudaMalloc(&C.hss_store, need));
    C.hss_store_bytes = need;
  }
  CUDA_OK(cudaMemsetAsync(C.hss_store, 0, need, st));
  const long long sbase = rel<T>(f, C.hss_store);
  std::vector<CopyDesc> cds;
  auto cpy = [&](long long src, int lds, long long dst, int ldd, int rows, int cols, int gap_at, int gap_skip, int mode) {
    if (rows <= 0 || cols <= 0) return;
    CopyDesc c{};
    c.src = src; c.lds = lds; c.dst = dst; c.ldd = ldd; c.rows = rows; c.cols = cols; c.gap_at = gap_at; c.gap_skip = gap_skip; c.mode = mode;
    cds.push_back(c);
  };
  for (int hi : all) {
    HssFront& H = f->hss[hi];
    const Round& R = *rounds[where[hi].first];
    const int q = where[hi].second;
    const int nn = (int)H.tree.size();
    H.sbase = sbase;
    for (int t = 0; t < nn; ++t) {
      const HssTreeNode& tn = H.tree[t];
      const HssStored& S = H.st[t];
      const HNode& N = R.nd[R.nbase[q] + t];
      const int m = tn.hi - tn.lo;
      if (tn.left < 0) {
        cpy(N.m0, N.ldm0, sbase + S.D, S.ldD, m, m, m, 0, 0);
        cpy(N.e[0], N.lde, sbase + S.U, S.ldU, m, S.r0, m, 0, 0);
        cpy(N.e[1], N.lde, sbase + S.VH, S.ldVH, m, S.r1, m, 0, 1);
      } else {
        const HssStored &Sa = H.st[tn.left], &Sb = H.st[tn.right];
        if (tn.parent >= 0) {
          cpy(N.e[0], N.lde, sbase + S.R, S.ldR, Sa.r0 + Sb.r0, S.r0, Sa.r0, N.capl - Sa.r0, 0);
          cpy(N.e[1], N.lde, sbase + S.WH, S.ldWH, Sa.r1 + Sb.r1, S.r1, Sa.r1, N.capl - Sa.r1, 1);
        }
        cpy(N.m0, N.ldm0, sbase + S.B12, S.ldB12, Sa.r0, Sb.r1, Sa.r0, 0, 0);
        cpy(N.m1, N.ldm1, sbase + S.B21, S.ldB21, Sb.r0, Sa.r1, Sb.r0, 0, 0);
      }
    }
    f->stats.hss_nodes += nn;
  }
  if (!cds.empty()) {
    DevVec<CopyDesc> dc(f);
    dc.upload(cds, st);
    k_copy_desc<T><<<dim3((unsigned)cds.size(), 2), 256, 0, st>>>(dc.d, (T*)f->pool);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(st));
    ++f->stats.launches_factor;
  }
  rounds.clear();   // the sample workspaces are no longer needed
  f->stats.hss_bytes += (double)need;
  g_hclk.tick(6, st);
  // ---- the matrix the HSS form represents goes back into the dense slots; the parent assembles from it (:126-140) ----
  expand<T>(f, all, true, nullptr);
  g_hclk.tick(7, st);
  if (C.li + 1 >= (int)f->levels.size() || &C == &f->clevels.back()) g_hclk.report();
}

}  // namespace

void hs_hss_setup() {
  CUDA_OK(cudaFuncSetAttribute(k_hss_qrcp<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_hss_qrcp<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_hss_qrcp_warp<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(cudaFuncSetAttribute(k_hss_qrcp_warp<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

// cluster tree of `bisection_cluster((n1, m); leafsize)` (factorization.jl:109), pre-order
static void hss_cluster(std::vector<HssTreeNode>& tree, int m, int n1, int leafsize) {
  tree.clear();
  std::function<int(int, int, int, int, bool)> rec = [&](int lo, int hi, int parent, int isright, bool force) -> int {
    const int id = (int)tree.size();
    tree.push_back(HssTreeNode{});
    tree[id].lo = lo; tree[id].hi = hi; tree[id].parent = parent; tree[id].isright = isright;
    int mid = -1;
    if (force) mid = n1;
    else if (hi - lo > leafsize) mid = lo + (hi - lo + 1) / 2;
    if (mid >= 0) {
      const int l = rec(lo, mid, id, 0, false);
      const int r = rec(mid, hi, id, 1, false);
      tree[id].left = l; tree[id].right = r;
      tree[id].height = 1 + std::max(tree[l].height, tree[r].height);
    }
    return id;
  };
  rec(0, m, -1, 0, n1 > 0 && n1 < m);
}

void hs_hss_plan(hs_fac* f) {
  f->hss.clear();
  for (CompFront& cf : f->comp) cf.hss = -1;
  if (!f->opts.hss || f->comp.empty()) return;
  std::vector<int64_t> front2node(f->fronts.size(), -1);
  for (int64_t k = 0; k < f->nnodes; ++k) front2node[f->node2front[k]] = k;
  long long max_m = 0;
  std::vector<int> hperm;
  for (size_t c = 0; c < f->comp.size(); ++c) {
    CompFront& cf = f->comp[c];
    const int64_t node = front2node[cf.fi];
    HssFront H;
    H.comp = (int)c;
    for (int64_t q = f->iloc_ptr[node]; q < f->iloc_ptr[node + 1]; ++q) H.perm.push_back(f->iloc_idx[q]);
    H.n1 = (int)H.perm.size();
    for (int64_t q = f->bloc_ptr[node]; q < f->bloc_ptr[node + 1]; ++q) H.perm.push_back(f->bloc_idx[q]);
    H.m = (int)H.perm.size();
    if (H.m == 0) continue;
    hss_cluster(H.tree, H.m, H.n1, (int)f->opts.leafsize);
    if (H.tree.size() <= 1) continue;   // a single dense block: S stays dense (an HssMatrix leaf holds exactly that)
    H.st.assign(H.tree.size(), HssStored{});
    cf.hss = (int)f->hss.size();
    max_m = std::max<long long>(max_m, H.m);
    f->hss.push_back(std::move(H));
  }
  if (f->hss.empty()) return;
  for (HssFront& H : f->hss) { H.perm_off = (int)hperm.size(); hperm.insert(hperm.end(), H.perm.begin(), H.perm.end()); }
  cudaStream_t st = f->ctx->stream;
  CUDA_OK(cudaMalloc((void**)&f->d_hperm, std::max<size_t>(hperm.size(), 1) * sizeof(int)));
  CUDA_OK(cudaMemcpyAsync(f->d_hperm, hperm.data(), hperm.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  if (f->opts.sketch_omega) {
    if (!f->opts.sketch_psi || f->opts.sketch_rows < max_m || f->opts.sketch_cols < 1)
      throw hs_error(HS_EARG, "hs_factor: sketch matrices need at least " + std::to_string(max_m) + " rows (largest compressed boundary)");
    f->sk_rows = f->opts.sketch_rows; f->sk_cols = f->opts.sketch_cols;
    const size_t one = (size_t)f->sk_rows * f->sk_cols * f->esz;
    CUDA_OK(cudaMalloc(&f->d_sk, 2 * one));
    CUDA_OK(cudaMemcpyAsync(f->d_sk, f->opts.sketch_omega, one, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync((char*)f->d_sk + one, f->opts.sketch_psi, one, cudaMemcpyHostToDevice, st));
  } else {
    f->sk_rows = max_m; f->sk_cols = 0;   // generated on first use (ensure_sketch)
  }
  CUDA_OK(cudaStreamSynchronize(st));
  // the caller's sketch buffers are not referenced after this call
  f->opts.sketch_psi = nullptr;
  if (f->opts.sketch_omega) f->opts.sketch_omega = (const void*)f->d_sk;   // keeps "host-supplied" distinguishable from "generate"
}

void hs_hss_copy(hs_fac* f, const std::vector<CopyDesc>& blocks) {
  if (blocks.empty()) return;
  cudaStream_t st = f->ctx->stream;
  DevVec<CopyDesc> dc(f);
  dc.upload(blocks, st);
  if (f->dtype == HS_F64) k_copy_desc<double><<<dim3((unsigned)blocks.size(), 4), 256, 0, st>>>(dc.d, (double*)f->pool);
  else k_copy_desc<cplx><<<dim3((unsigned)blocks.size(), 4), 256, 0, st>>>(dc.d, (cplx*)f->pool);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(st));
  ++f->stats.launches_factor;
}

void hs_hss_build(hs_fac* f, CompLevel& C) {
  if (f->hss.empty()) return;
  if (f->dtype == HS_F64) build_impl<double>(f, C); else build_impl<cplx>(f, C);
}

void hs_hss_dense(hs_fac* f, int hi, void* out_host) {
  std::vector<int> one{hi};
  if (f->dtype == HS_F64) expand<double>(f, one, false, (double*)out_host); else expand<cplx>(f, one, false, (cplx*)out_host);
}
