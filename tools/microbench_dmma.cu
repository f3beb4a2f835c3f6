// DMMA (mma.sync.m8n8k4.f64) throughput on B200 -- decides whether the 8x8 Gram accumulation of
// the GN kernel can live in tensor-core fragments (4 registers) instead of 54 registers.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(double* out, double s, int iters) {
  double acc[NACC][2];
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = s + threadIdx.x, b = s * 0.5 + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma(acc[i][0], acc[i][1], a, b);
  }
  double r = 0;
  for (int i = 0; i < NACC; ++i) r += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NACC>
void run(int warps_per_block) {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
  const int iters = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NACC><<<148 * 2, warps_per_block * 32>>>(out, 1.0, iters);
  cudaEventRecord(e0);
  k<NACC><<<148 * 2, warps_per_block * 32>>>(out, 1.0, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double mmas = 148.0 * 2 * warps_per_block * iters * 8 * NACC;
  printf("NACC=%d warps/blk=%d: %.3f ms, %.1f G DMMA/s, %.2f TFLOP/s fp64, %.2f cycles/DMMA/SM @1965MHz\n", NACC,
         warps_per_block, ms, mmas / ms / 1e6, mmas * 512 / ms / 1e9, 1965e3 * ms / (mmas / 148));
  cudaFree(out);
}
int main() {
  run<1>(8); run<2>(8); run<4>(8); run<1>(16); run<4>(16);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
