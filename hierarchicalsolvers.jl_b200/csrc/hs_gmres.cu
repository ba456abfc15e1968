// Device-resident restarted GMRES with the factorization as right preconditioner.
//
// Takes over the driver loop of the reference, `gmres(A, b; Pr=F, reltol, restart, maxiter)`
// (test/rungmres.jl:47-48; IterativeSolvers.jl 0.9.0 semantics: x0 = 0, right preconditioning through
// ldiv!(y, Pr, x), modified Gram-Schmidt, Givens least squares, one solution update per restart cycle).
// All n-vectors stay in HBM; only the Hessenberg scalars cross to the host.
#include <cuda_runtime.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

#include <chrono>
#include <complex>
#include <cstdio>

#include "hs_fac.cuh"

namespace {

constexpr int RED_BLOCKS = 592;  // 4 × 148 SMs
constexpr int RED_THREADS = 256;

// y = A·x, CSR, one warp per row group of 32 rows is overkill for ≤7 nnz/row stencils: thread per row
template <typename T>
__global__ void k_spmv_csr(long long n, const long long* __restrict__ ptr, const long long* __restrict__ col,
                           const T* __restrict__ val, const T* __restrict__ x, T* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T acc = hs_zero<T>();
  for (long long p = ptr[i]; p < ptr[i + 1]; ++p) acc = hs_fma(acc, val[p], x[col[p]]);
  y[i] = acc;
}

__device__ __forceinline__ cplx conj_mul(double a, double b) { return cplx{a * b, 0.0}; }
__device__ __forceinline__ cplx conj_mul(cplx a, cplx b) {  // conj(a)·b
  return cplx{a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x};
}

// deterministic two-stage dot product  out = Σ conj(a_i)·b_i
template <typename T>
__global__ void __launch_bounds__(RED_THREADS) k_dot_partial(long long n, const T* __restrict__ a, const T* __restrict__ b,
                                                              cplx* __restrict__ part) {
  __shared__ cplx s[RED_THREADS];
  cplx acc{0.0, 0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const cplx t = conj_mul(a[i], b[i]);
    acc.x += t.x; acc.y += t.y;
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s[threadIdx.x].x += s[threadIdx.x + o].x; s[threadIdx.x].y += s[threadIdx.x + o].y; }
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}

__global__ void __launch_bounds__(RED_THREADS) k_dot_final(int nb, const cplx* __restrict__ part, cplx* __restrict__ out) {
  __shared__ cplx s[RED_THREADS];
  cplx acc{0.0, 0.0};
  for (int i = threadIdx.x; i < nb; i += RED_THREADS) { acc.x += part[i].x; acc.y += part[i].y; }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s[threadIdx.x].x += s[threadIdx.x + o].x; s[threadIdx.x].y += s[threadIdx.x + o].y; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0];
}

__device__ __forceinline__ double scal_of(cplx a, double*) { return a.x; }
__device__ __forceinline__ cplx scal_of(cplx a, cplx*) { return a; }

// y += alpha·x   /   y = alpha·x
template <typename T>
__global__ void k_axpy(long long n, cplx alpha, const T* __restrict__ x, T* __restrict__ y, int assign) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T a = scal_of(alpha, (T*)nullptr);
  y[i] = assign ? hs_mul(a, x[i]) : hs_fma(y[i], a, x[i]);
}

// CSC → CSR on the device
__global__ void k_count_rows(long long nnz, const long long* __restrict__ rowval, unsigned long long* __restrict__ cnt) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) atomicAdd(&cnt[rowval[p]], 1ull);
}
// the CSR image keeps, per entry, the position of that entry in the CSC value array (`csrc`): new values of the same
// pattern (hs_refactor) are then one gather away
__global__ void k_fill_csr(long long n, const long long* __restrict__ colptr, const long long* __restrict__ rowval,
                           const long long* __restrict__ rptr, unsigned long long* __restrict__ fill,
                           long long* __restrict__ ccol, long long* __restrict__ csrc) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  for (long long p = colptr[j]; p < colptr[j + 1]; ++p) {
    const long long i = rowval[p];
    const long long q = rptr[i] + (long long)atomicAdd(&fill[i], 1ull);
    ccol[q] = j;
    csrc[q] = p;
  }
}
// rows were filled in a racy order: sort each row by column so the mat-vec is bitwise reproducible
__global__ void k_sort_rows(long long n, const long long* __restrict__ rptr, long long* __restrict__ ccol, long long* __restrict__ csrc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long a = rptr[i], b = rptr[i + 1];
  for (long long p = a + 1; p < b; ++p) {
    const long long c = ccol[p];
    const long long v = csrc[p];
    long long q = p - 1;
    while (q >= a && ccol[q] > c) { ccol[q + 1] = ccol[q]; csrc[q + 1] = csrc[q]; --q; }
    ccol[q + 1] = c;
    csrc[q + 1] = v;
  }
}
template <typename T>
__global__ void k_gather_vals(long long nnz, const long long* __restrict__ csrc, const T* __restrict__ nzval, T* __restrict__ cval) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nnz) cval[q] = nzval[csrc[q]];
}

// CSR image of the matrix the factorization holds: pattern + source map built once, values gathered again whenever
// hs_refactor installed new ones (f->csr_stale)
template <typename T> void ensure_csr(hs_fac* f) {
  cudaStream_t st = f->ctx->stream;
  const long long n = f->n, nnz = f->nnz;
  if (!f->d_csr_ptr) {
    unsigned long long* cnt = nullptr;
    CUDA_OK(cudaMalloc((void**)&f->d_csr_ptr, (size_t)(n + 1) * sizeof(long long)));
    CUDA_OK(cudaMalloc((void**)&f->d_csr_col, std::max<size_t>(nnz, 1) * sizeof(long long)));
    CUDA_OK(cudaMalloc((void**)&f->d_csr_src, std::max<size_t>(nnz, 1) * sizeof(long long)));
    CUDA_OK(cudaMalloc(&f->d_csr_val, std::max<size_t>(nnz, 1) * sizeof(T)));
    CUDA_OK(cudaMalloc((void**)&cnt, (size_t)(n + 1) * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)(n + 1) * sizeof(unsigned long long), st));
    if (nnz) k_count_rows<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, f->d_rowval, cnt);
    thrust::exclusive_scan(thrust::cuda::par.on(st), (long long*)cnt, (long long*)cnt + n + 1, f->d_csr_ptr);
    CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)(n + 1) * sizeof(unsigned long long), st));
    k_fill_csr<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, f->d_colptr, f->d_rowval, f->d_csr_ptr, cnt, f->d_csr_col, f->d_csr_src);
    k_sort_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, f->d_csr_ptr, f->d_csr_col, f->d_csr_src);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(cnt);
    f->csr_stale = true;
  }
  if (f->csr_stale) {
    if (nnz) k_gather_vals<T><<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, f->d_csr_src, (const T*)f->d_nzval, (T*)f->d_csr_val);
    CUDA_OK(cudaGetLastError());
    ++f->ctx->launches;
    f->csr_stale = false;
  }
}

using zc = std::complex<double>;

template <typename T>
void gmres_impl(hs_ctx* ctx, hs_fac* f, long long n, const long long* rptr, const long long* ccol, const T* cval,
                const void* b_host, void* x_host, double reltol, int64_t restart, int64_t maxiter, double* resnorm,
                int64_t* niter, int32_t* converged, int on_device) {
  cudaStream_t st = ctx->stream;
  const bool timing = getenv("HS_PLAN_TIMING") != nullptr;
  auto tg0 = std::chrono::steady_clock::now();
  auto tick = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[gmres] %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - tg0).count());
    tg0 = t1;
  };
  T *V = nullptr, *w = nullptr, *z = nullptr, *x = nullptr, *bdev = nullptr, *Zk = nullptr;
  cplx *part = nullptr, *dres = nullptr;
  // one grow-only workspace per context: cudaMalloc/cudaFree of the ~1 GB Krylov basis per call would dominate
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t szv = up((size_t)n * sizeof(T));
  // The solution update x += Pr⁻¹·(V·y) equals Σ y_j·z_j with z_j = Pr⁻¹·v_j, which the Arnoldi step already computed:
  // the first ZK of them are kept, so a solve that converges within ZK iterations of a cycle (a direct-solver
  // preconditioner needs one) skips the second preconditioner application of IterativeSolvers' update.
  const int64_t ZK = std::min<int64_t>(restart, 4);
  const size_t need = szv * (restart + 1) + (4 + ZK) * szv + up(RED_BLOCKS * sizeof(cplx)) + 256;
  if (ctx->gm_bytes < need) {
    cudaFree(ctx->gm_buf);
    ctx->gm_buf = nullptr; ctx->gm_bytes = 0;
    CUDA_OK(cudaMalloc(&ctx->gm_buf, need));
    ctx->gm_bytes = need;
  }
  {
    char* p = (char*)ctx->gm_buf;
    V = (T*)p; p += szv * (restart + 1);
    w = (T*)p; p += szv;
    z = (T*)p; p += szv;
    x = (T*)p; p += szv;
    bdev = (T*)p; p += szv;
    Zk = (T*)p; p += szv * ZK;
    part = (cplx*)p; p += up(RED_BLOCKS * sizeof(cplx));
    dres = (cplx*)p;
  }
  tick("workspace");
  const unsigned gb = (unsigned)((n + 255) / 256);
  auto dot = [&](const T* a, const T* bb) -> zc {
    ctx->launches += 2;
    k_dot_partial<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, a, bb, part);
    k_dot_final<<<1, RED_THREADS, 0, st>>>(RED_BLOCKS, part, dres);
    cplx h;
    CUDA_OK(cudaMemcpyAsync(&h, dres, sizeof(cplx), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return zc(h.x, h.y);
  };
  auto axpy = [&](zc alpha, const T* xx, T* yy, int assign) {
    ++ctx->launches;
    k_axpy<T><<<gb, 256, 0, st>>>(n, cplx{alpha.real(), alpha.imag()}, xx, yy, assign);
  };
  auto precond = [&](const T* in, T* out) {
    if (f) {
      int32_t rc = hs_solve(f, 1, in, n, out, n, 1);
      if (rc != HS_OK) throw hs_error(rc, hs_last_error());
    } else {
      CUDA_OK(cudaMemcpyAsync(out, in, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    }
  };
  auto matvec = [&](const T* in, T* out) { ++ctx->launches; k_spmv_csr<T><<<gb, 256, 0, st>>>(n, rptr, ccol, cval, in, out); };

  CUDA_OK(cudaMemcpyAsync(bdev, b_host, (size_t)n * sizeof(T), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(T), st));
  // r0 = b (x0 = 0)
  CUDA_OK(cudaMemcpyAsync(w, bdev, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
  double beta = std::sqrt(dot(w, w).real());
  tick("b to device + norm");
  const double tol = reltol * beta;
  double resid = beta;
  int64_t it = 0;
  std::vector<zc> H((restart + 1) * restart), cs(restart), sn(restart), g(restart + 1);
  auto Hm = [&](int64_t i, int64_t j) -> zc& { return H[j * (restart + 1) + i]; };
  while (it < maxiter && resid > tol) {
    std::fill(H.begin(), H.end(), zc(0));
    std::fill(g.begin(), g.end(), zc(0));
    axpy(zc(1.0 / beta), w, V, 1);  // V[:,0] = r/β
    g[0] = beta;
    int64_t k = 0;
    while (k < restart && it < maxiter && resid > tol) {
      T* vk = V + (size_t)k * n;
      T* vk1 = V + (size_t)(k + 1) * n;
      T* zk = k < ZK ? (T*)((char*)Zk + (size_t)k * szv) : z;
      precond(vk, zk);
      matvec(zk, vk1);
      for (int64_t j = 0; j <= k; ++j) {  // modified Gram-Schmidt
        const T* vj = V + (size_t)j * n;
        const zc h = dot(vj, vk1);
        Hm(j, k) = h;
        axpy(-h, vj, vk1, 0);
      }
      const double hn = std::sqrt(dot(vk1, vk1).real());
      Hm(k + 1, k) = hn;
      if (hn != 0.0) axpy(zc(1.0 / hn), vk1, vk1, 1);
      for (int64_t j = 0; j < k; ++j) {
        const zc t = cs[j] * Hm(j, k) + sn[j] * Hm(j + 1, k);
        Hm(j + 1, k) = -std::conj(sn[j]) * Hm(j, k) + cs[j] * Hm(j + 1, k);
        Hm(j, k) = t;
      }
      const zc a = Hm(k, k), c = Hm(k + 1, k);
      const double den = std::sqrt(std::norm(a) + std::norm(c));
      if (den == 0.0) { cs[k] = 1.0; sn[k] = 0.0; }
      else if (std::abs(a) == 0.0) { cs[k] = 0.0; sn[k] = 1.0; }
      else { cs[k] = std::abs(a) / den; sn[k] = (a / std::abs(a)) * std::conj(c) / den; }
      Hm(k, k) = cs[k] * a + sn[k] * c;
      Hm(k + 1, k) = 0.0;
      g[k + 1] = -std::conj(sn[k]) * g[k];
      g[k] = cs[k] * g[k];
      resid = std::abs(g[k + 1]);
      resnorm[it] = resid;
      ++k; ++it;
    }
    // y = H(0:k,0:k)⁻¹ g,  x += Pr⁻¹ (V y)
    std::vector<zc> y(k);
    for (int64_t i = k - 1; i >= 0; --i) {
      zc s = g[i];
      for (int64_t j = i + 1; j < k; ++j) s -= Hm(i, j) * y[j];
      y[i] = s / Hm(i, i);
    }
    if (k <= ZK) {
      for (int64_t j = 0; j < k; ++j) axpy(y[j], (const T*)((const char*)Zk + (size_t)j * szv), x, 0);
    } else {
      CUDA_OK(cudaMemsetAsync(w, 0, (size_t)n * sizeof(T), st));
      for (int64_t j = 0; j < k; ++j) axpy(y[j], V + (size_t)j * n, w, 0);
      precond(w, z);
      axpy(zc(1.0), z, x, 0);
    }
    if (it < maxiter && resid > tol) {  // restart: r = b − A x
      matvec(x, w);
      axpy(zc(-1.0), bdev, w, 0);
      axpy(zc(-1.0), w, w, 1);
      beta = std::sqrt(dot(w, w).real());
      resid = beta;
    }
  }
  CUDA_OK(cudaGetLastError());
  tick("iterations");
  CUDA_OK(cudaMemcpyAsync(x_host, x, (size_t)n * sizeof(T), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  tick("x to host");
  *niter = it;
  *converged = resid <= tol;
}

template <typename T>
void gmres_entry(hs_ctx* ctx, long long n, const int64_t* colptr, const int64_t* rowval, const void* nzval, int32_t base,
                 hs_fac* f, const void* b, void* x, double reltol, int64_t restart, int64_t maxiter, double* resnorm,
                 int64_t* niter, int32_t* converged, int on_device) {
  cudaStream_t st = ctx->stream;
  long long *rptr = nullptr, *ccol = nullptr;
  T* cval = nullptr;
  bool own = false;
  struct Tmp { std::vector<void*> p; ~Tmp() { for (void* q : p) cudaFree(q); } } tmp;
  if (!colptr) {
    if (!f) throw hs_error(HS_EARG, "hs_gmres: no matrix given and no factorization to take it from");
    ensure_csr<T>(f);
    rptr = f->d_csr_ptr; ccol = f->d_csr_col; cval = (T*)f->d_csr_val;
  } else {
    // host CSC → CSR on the host (A given explicitly, e.g. a matrix that differs from the factored one)
    const long long nnz = colptr[n] - base;
    std::vector<long long> hp(n + 1, 0), hc(nnz);
    std::vector<T> hv(nnz);
    const T* vin = (const T*)nzval;
    for (long long p = 0; p < nnz; ++p) ++hp[rowval[p] - base + 1];
    for (long long i = 0; i < n; ++i) hp[i + 1] += hp[i];
    std::vector<long long> pos(hp.begin(), hp.end() - 1);
    for (long long j = 0; j < n; ++j)
      for (long long p = colptr[j] - base; p < colptr[j + 1] - base; ++p) {
        const long long q = pos[rowval[p] - base]++;
        hc[q] = j; hv[q] = vin[p];
      }
    CUDA_OK(cudaMalloc((void**)&rptr, (size_t)(n + 1) * sizeof(long long))); tmp.p.push_back(rptr);
    CUDA_OK(cudaMalloc((void**)&ccol, std::max<size_t>(nnz, 1) * sizeof(long long))); tmp.p.push_back(ccol);
    CUDA_OK(cudaMalloc((void**)&cval, std::max<size_t>(nnz, 1) * sizeof(T))); tmp.p.push_back(cval);
    CUDA_OK(cudaMemcpyAsync(rptr, hp.data(), (size_t)(n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(ccol, hc.data(), (size_t)nnz * sizeof(long long), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(cval, hv.data(), (size_t)nnz * sizeof(T), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));
    own = true;
  }
  (void)own;
  if (getenv("HS_PLAN_TIMING")) { cudaStreamSynchronize(st); fprintf(stderr, "[gmres] matrix ready (%s)\n", colptr ? "host CSC given" : "taken from the factorization"); }
  gmres_impl<T>(ctx, f, n, rptr, ccol, cval, b, x, reltol, restart, maxiter, resnorm, niter, converged, on_device);
}

}  // namespace

// y = A·x on the device with the matrix the factorization holds (CSR image built on first use, as hs_gmres does)
template <typename T> static void spmv_impl(hs_fac* f, const void* x, void* y) {
  ensure_csr<T>(f);
  const unsigned gb = (unsigned)((f->n + 255) / 256);
  k_spmv_csr<T><<<gb, 256, 0, f->ctx->stream>>>(f->n, f->d_csr_ptr, f->d_csr_col, (const T*)f->d_csr_val, (const T*)x, (T*)y);
  CUDA_OK(cudaGetLastError());
  ++f->ctx->launches;
}

extern "C" int32_t hs_spmv(hs_fac* f, const void* x, void* y) {
  HS_TRY_BEGIN
  if (!f || !x || !y) return hs_fail(HS_EARG, "hs_spmv: null argument");
  if (x == y) return hs_fail(HS_EARG, "hs_spmv: x and y must not alias");
  CUDA_OK(cudaSetDevice(f->ctx->device));
  if (f->dtype == HS_F64) spmv_impl<double>(f, x, y); else spmv_impl<cplx>(f, x, y);
  return HS_OK;
  HS_TRY_END
}

// order-independent checksum of the bit patterns of the matrix values a factorization holds (wrapping sum and XOR of the
// 64-bit words): lets a host binding decide cheaply whether a matrix it is handed is the one already resident in HBM
__global__ void k_checksum(const unsigned long long* __restrict__ w, long long n, unsigned long long* __restrict__ out) {
  unsigned long long s = 0, x = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long v = w[i];
    s += v; x ^= v;
  }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); x ^= __shfl_xor_sync(0xffffffffu, x, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, s); atomicXor(out + 1, x); }
}

extern "C" int32_t hs_matrix_checksum(hs_fac* f, uint64_t* sum_out, uint64_t* xor_out) {
  HS_TRY_BEGIN
  if (!f || !sum_out || !xor_out) return hs_fail(HS_EARG, "hs_matrix_checksum: null argument");
  CUDA_OK(cudaSetDevice(f->ctx->device));
  cudaStream_t st = f->ctx->stream;
  unsigned long long* d = nullptr;
  CUDA_OK(cudaMalloc((void**)&d, 2 * sizeof(unsigned long long)));
  CUDA_OK(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), st));
  const long long words = f->nnz * (long long)(f->esz / 8);
  if (words) k_checksum<<<592, 256, 0, st>>>((const unsigned long long*)f->d_nzval, words, d);
  unsigned long long h[2];
  cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  CUDA_OK(e);
  ++f->ctx->launches;
  *sum_out = h[0]; *xor_out = h[1];
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_gmres(hs_ctx* ctx, hs_dtype dtype, int64_t n, const int64_t* colptr, const int64_t* rowval,
                            const void* nzval, int32_t index_base, hs_fac* fac, const void* b, void* x, double reltol,
                            int64_t restart, int64_t maxiter, double* resnorm, int64_t* niter, int32_t* converged,
                            int32_t on_device) {
  HS_TRY_BEGIN
  if (!ctx || !b || !x || !resnorm || !niter || !converged) return hs_fail(HS_EARG, "hs_gmres: null argument");
  if (restart < 1 || maxiter < 0 || n <= 0) return hs_fail(HS_EARG, "hs_gmres: bad restart/maxiter/n");
  if (fac && (fac->n != n || fac->dtype != dtype)) return hs_fail(HS_EDIM, "hs_gmres: factorization does not match the system");
  if (colptr && (!rowval || !nzval)) return hs_fail(HS_EARG, "hs_gmres: incomplete matrix");
  CUDA_OK(cudaSetDevice(ctx->device));
  if (dtype == HS_F64)
    gmres_entry<double>(ctx, n, colptr, rowval, nzval, index_base, fac, b, x, reltol, restart, maxiter, resnorm, niter, converged, on_device);
  else
    gmres_entry<cplx>(ctx, n, colptr, rowval, nzval, index_base, fac, b, x, reltol, restart, maxiter, resnorm, niter, converged, on_device);
  return HS_OK;
  HS_TRY_END
}
