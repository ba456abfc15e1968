// Per-iteration cost of (a) cluster.sync(), (b) an all-to-all mailbox exchange signalled with st.async + mbarrier
// complete_tx, for clusters of 2..16 CTAs of 256 threads.   nvcc -arch=sm_100a -O3 -o cluster_sync cluster_sync.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(256, 1) k_sync(int iters, long long* out) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double box[2][16];
  const int C = cluster.num_blocks(), r = cluster.block_rank();
  double acc = 0;
  cluster.sync();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int par = it & 1;
    if (threadIdx.x < C) { double* dst = cluster.map_shared_rank(&box[par][r], threadIdx.x); *dst = it + r; }
    cluster.sync();
    if (threadIdx.x < C) acc += box[par][threadIdx.x];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)acc; }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}

__global__ void __launch_bounds__(256, 1) k_mbar(int iters, int with_bar, long long* out) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double box[2][16];
  __shared__ __align__(8) unsigned long long mbar[2];
  const int C = cluster.num_blocks(), r = cluster.block_rank();
  if (threadIdx.x == 0) {
    for (int p = 0; p < 2; ++p) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[p])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  cluster.sync();
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int par = it & 1, phase = (it >> 1) & 1;
    if (with_bar) __syncthreads();
    if (threadIdx.x < C) {
      const uint32_t dst = mapa(smem_u32(&box[par][r]), threadIdx.x), mb = mapa(smem_u32(&mbar[par]), threadIdx.x);
      const double v = it + r;
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst),
                   "l"(__double_as_longlong(v)), "r"(mb) : "memory");
    }
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[par])), "r"(8 * C) : "memory");
    {
      uint32_t ok = 0;
      const uint32_t mb = smem_u32(&mbar[par]);
      while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(mb), "r"(phase) : "memory");
    }
    if (threadIdx.x < C) acc += box[par][threadIdx.x];
  }
  long long t1 = clock64();
  cluster.sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)acc; }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 20000;
  cudaFuncSetAttribute(k_sync, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(k_mbar, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int C : {2, 4, 8, 16}) {
    for (int mode = 0; mode < 3; ++mode) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(C); cfg.blockDim = dim3(256);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaError_t e = mode == 0 ? cudaLaunchKernelEx(&cfg, k_sync, iters, d) : cudaLaunchKernelEx(&cfg, k_mbar, iters, mode - 1, d);
      if (e != cudaSuccess) { printf("C=%d launch: %s\n", C, cudaGetErrorString(e)); continue; }
      e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("cluster %2d  %-34s %8.1f cycles/iter  (%s, chk %lld)\n", C,
             mode == 0 ? "push + cluster.sync" : mode == 1 ? "st.async + mbarrier" : "syncthreads + st.async + mbarrier",
             (double)h[0] / iters, cudaGetErrorString(e), h[1]);
    }
  }
  return 0;
}
