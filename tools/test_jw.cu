// packed vs scalar evaluation of the Jw entries (Tracker.cpp:455-467)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__global__ void k(unsigned long long* bad, float* ex) {
  uint32_t s = 4242u + blockIdx.x * 7919u + threadIdx.x * 104729u;
  const float fx = 42.8575f, fy = 42.8525f;
  for (int it = 0; it < 2000; ++it) {
    float x2 = (rng(s) >> 8) * (80.0f / 16777216.0f), y2 = (rng(s) >> 8) * (64.0f / 16777216.0f);
    float iz = 0.9f + (rng(s) >> 8) * (0.2f / 16777216.0f);
    // scalar
    const float fxx2 = __fmul_rn(fx, x2), fyx2 = __fmul_rn(fy, x2), fyy2 = __fmul_rn(fy, y2);
    float r[10];
    r[0] = __fmul_rn(fx, iz);
    r[1] = -__fmul_rn(__fmul_rn(fxx2, iz), iz);
    r[2] = -__fmul_rn(__fmul_rn(__fmul_rn(fxx2, y2), iz), iz);
    r[3] = __fmul_rn(fx, __fadd_rn(1.0f, __fmul_rn(__fmul_rn(__fmul_rn(x2, x2), iz), iz)));
    r[4] = __fmul_rn(__fmul_rn(-fx, y2), iz);
    r[5] = __fmul_rn(fy, iz);
    r[6] = -__fmul_rn(__fmul_rn(fyy2, iz), iz);
    r[7] = -__fmul_rn(fy, __fadd_rn(1.0f, __fmul_rn(__fmul_rn(__fmul_rn(y2, y2), iz), iz)));
    r[8] = __fmul_rn(__fmul_rn(__fmul_rn(fyx2, y2), iz), iz);
    r[9] = __fmul_rn(fyx2, iz);
    // packed
    const float2 fxy = make_float2(fx, fy), xy2 = make_float2(x2, y2), iz2 = make_float2(iz, iz);
    const float2 p1 = __fmul2_rn(fxy, xy2);
    const float2 w00_11 = __fmul2_rn(fxy, iz2);
    const float2 p4 = __fmul2_rn(__fmul2_rn(p1, iz2), iz2);
    const float2 t1 = __fmul2_rn(make_float2(-fx, fy), make_float2(y2, x2));
    const float2 w05_15 = __fmul2_rn(t1, iz2);
    const float2 q3 = __fmul2_rn(__fmul2_rn(__fmul2_rn(make_float2(p1.x, t1.y), make_float2(y2, y2)), iz2), iz2);
    const float2 s3 = __fmul2_rn(__fmul2_rn(__fmul2_rn(xy2, xy2), iz2), iz2);
    const float2 s5 = __fmul2_rn(fxy, make_float2(__fadd_rn(1.0f, s3.x), __fadd_rn(1.0f, s3.y)));
    float p[10] = {w00_11.x, -p4.x, -q3.x, s5.x, w05_15.x, w00_11.y, -p4.y, -s5.y, q3.y, w05_15.y};
    for (int i = 0; i < 10; ++i)
      if (__float_as_uint(p[i]) != __float_as_uint(r[i])) {
        unsigned long long n = atomicAdd(bad + i, 1ull);
        if (n == 0) { ex[i*5]=x2; ex[i*5+1]=y2; ex[i*5+2]=iz; ex[i*5+3]=p[i]; ex[i*5+4]=r[i]; }
      }
  }
}
int main() {
  unsigned long long* bad; float* ex; cudaMallocManaged(&bad, 10*8); cudaMallocManaged(&ex, 10*5*4);
  for (int i = 0; i < 10; ++i) bad[i] = 0;
  k<<<148 * 4, 256>>>(bad, ex); cudaDeviceSynchronize();
  for (int i = 0; i < 10; ++i) printf("w[%d]: mismatches %llu  (x2=%.9g y2=%.9g iz=%.9g packed=%.9g scalar=%.9g)\n", i, bad[i], ex[i*5], ex[i*5+1], ex[i*5+2], ex[i*5+3], ex[i*5+4]);
  return 0;
}
