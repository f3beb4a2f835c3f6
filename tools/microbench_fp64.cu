// Throughput microbenchmark of the fp64 / conversion instructions the GN kernel leans on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_fp64 tools/microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template <int OP>
__global__ void k(double* out, float seedf, double seedd, int seedi) {
  float f0 = seedf + threadIdx.x, f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
  double d0 = seedd + threadIdx.x, d1 = d0 + 1., d2 = d0 + 2., d3 = d0 + 3.;
  int i0 = seedi + threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) { d0 = fma(d0, d1, d2); d1 = fma(d1, d2, d3); d2 = fma(d2, d3, d0); d3 = fma(d3, d0, d1); }
      if (OP == 1) { d0 += (double)f0; d1 += (double)f1; d2 += (double)f2; d3 += (double)f3; f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; }  // F2F.F64.F32 + DADD + FADD
      if (OP == 2) { f0 += (float)d0; f1 += (float)d1; f2 += (float)d2; f3 += (float)d3; d0 += 1.0; d1 += 1.0; d2 += 1.0; d3 += 1.0; }      // F2F.F32.F64 + FADD + DADD
      if (OP == 3) { d0 += (double)i0; d1 += (double)i1; d2 += (double)i2; d3 += (double)i3; i0 += 3; i1 += 3; i2 += 3; i3 += 3; }          // I2F.F64 + DADD
      if (OP == 4) { d0 += 1.5; d1 += 1.5; d2 += 1.5; d3 += 1.5; }                                                                          // DADD only
      if (OP == 5) { f0 = fmaf(f0, f1, f2); f1 = fmaf(f1, f2, f3); f2 = fmaf(f2, f3, f0); f3 = fmaf(f3, f0, f1); }                          // FFMA
      if (OP == 6) { f0 = __fdiv_rn(f0, f1); f1 = __fdiv_rn(f1, f2); f2 = __fdiv_rn(f2, f3); f3 = __fdiv_rn(f3, f0); }                      // IEEE fdiv
      if (OP == 7) {  // f32->f64 by bit manipulation (normal numbers)
        unsigned b0 = __float_as_uint(f0), b1 = __float_as_uint(f1);
        d0 += __hiloint2double((b0 & 0x80000000u) | (((b0 >> 3) & 0x0FFFFFFFu) + 0x38000000u), b0 << 29);
        d1 += __hiloint2double((b1 & 0x80000000u) | (((b1 >> 3) & 0x0FFFFFFFu) + 0x38000000u), b1 << 29);
        f0 += 1.f; f1 += 1.f;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3 + f0 + f1 + f2 + f3 + i0 + i1 + i2 + i3;
}
template <int OP>
void run(const char* name, double ops_per_iter_thread) {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<148 * 8, 256>>>(out, 1.f, 1.0, 1);
  cudaEventRecord(a);
  k<OP><<<148 * 8, 256>>>(out, 1.f, 1.0, 1);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double total = 148.0 * 8 * 256 * ITER * 8 * ops_per_iter_thread;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %8.3f ms  %8.2f Gop/s  = %6.1f ops/clk/SM (at %d MHz nominal)\n", name, ms, total / ms / 1e6,
         total / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("DFMA", 4);
  run<4>("DADD", 4);
  run<1>("F2F.F64.F32 (+DADD+FADD)", 4);
  run<2>("F2F.F32.F64 (+FADD+DADD)", 4);
  run<3>("I2F.F64.S32 (+DADD+IADD)", 4);
  run<5>("FFMA", 4);
  run<6>("FDIV.RN (IEEE)", 4);
  run<7>("f32->f64 bit trick (+DADD)", 2);
  return 0;
}
