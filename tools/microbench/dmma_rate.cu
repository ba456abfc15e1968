// DMMA.8x8x4 issue-rate microbenchmark: TFLOP/s as a function of warps per SM and independent accumulator chains.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC> __global__ void k(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC> void run(int warps, int ctas_per_sm) {
  int dev_sms = 148;
  double* out; cudaMalloc(&out, sizeof(double) * dev_sms * ctas_per_sm * warps * 32);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NACC><<<dev_sms * ctas_per_sm, warps * 32>>>(out, 100, 1.0, 2.0);
  cudaEventRecord(e0);
  k<NACC><<<dev_sms * ctas_per_sm, warps * 32>>>(out, iters, 1.0, 2.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops = 2.0 * 256 * NACC * (double)iters * warps * ctas_per_sm * dev_sms;
  printf("NACC=%2d warps/CTA=%2d CTAs/SM=%d : %.2f TFLOP/s  (%.1f cycles per DMMA per SMSP at 1.965GHz)\n", NACC, warps, ctas_per_sm, flops / ms / 1e9,
         ms * 1e-3 * 1.965e9 / ((double)NACC * iters * warps * ctas_per_sm / 4.0));
  cudaFree(out);
}
int main() {
  run<1>(4, 1); run<2>(4, 1); run<4>(4, 1); run<8>(4, 1); run<16>(4, 1); run<32>(4, 1);
  run<4>(8, 1); run<8>(8, 1); run<16>(8, 1); run<32>(8, 1);
  run<4>(16, 1); run<8>(16, 1); run<16>(16, 1);
  run<8>(8, 2); run<8>(32, 1);
  return 0;
}
