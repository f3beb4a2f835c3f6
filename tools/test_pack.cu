#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__global__ void k(unsigned long long* bad, float* ex) {
  uint32_t s = 99991u + blockIdx.x * 7919u + threadIdx.x * 104729u;
  unsigned long long nb = 0;
  for (int it = 0; it < 20000; ++it) {
    // operands spanning the full exponent range incl. denormal results
    float a = __uint_as_float((rng(s) & 0x807FFFFFu) | ((rng(s) % 254u + 1u) << 23));
    float b = __uint_as_float((rng(s) & 0x807FFFFFu) | ((rng(s) % 254u + 1u) << 23));
    float c = __uint_as_float((rng(s) & 0x807FFFFFu) | ((rng(s) % 254u + 1u) << 23));
    if (it % 7 == 0) a = __uint_as_float(rng(s) & 0x807FFFFFu);  // denormal input
    float2 m = __fmul2_rn(make_float2(a, b), make_float2(c, a));
    float2 ad = __fadd2_rn(make_float2(a, b), make_float2(c, a));
    float2 f = __ffma2_rn(make_float2(a, b), make_float2(c, a), make_float2(b, c));
    float r0 = __fmul_rn(a, c), r1 = __fmul_rn(b, a), r2 = __fadd_rn(a, c), r3 = __fadd_rn(b, a);
    float r4 = __fmaf_rn(a, c, b), r5 = __fmaf_rn(b, a, c);
    auto ne = [](float x, float y) { return __float_as_uint(x) != __float_as_uint(y) && !(x != x && y != y); };
    if (ne(m.x, r0) || ne(m.y, r1) || ne(ad.x, r2) || ne(ad.y, r3) || ne(f.x, r4) || ne(f.y, r5)) {
      if (atomicAdd(bad + 1, 1ull) < 6) { int i = atomicAdd((int*)(bad + 2), 1); if (i < 6) { ex[i*5]=a; ex[i*5+1]=c; ex[i*5+2]=m.x; ex[i*5+3]=r0; ex[i*5+4]=b; } }
      ++nb;
    }
  }
  atomicAdd(bad, nb);
}
int main() {
  unsigned long long* bad; float* ex; cudaMallocManaged(&bad, 64); cudaMallocManaged(&ex, 6*5*4);
  bad[0] = bad[1] = bad[2] = 0;
  k<<<148 * 4, 256>>>(bad, ex); cudaDeviceSynchronize();
  printf("packed vs scalar mismatches %llu of %llu (%s)\n", bad[0], 148ull * 4 * 256 * 20000, cudaGetErrorString(cudaGetLastError()));
  for (int i = 0; i < 6 && i < (int)bad[2]; ++i) printf("  a=%.9g c=%.9g packed=%.9g scalar=%.9g b=%.9g\n", ex[i*5], ex[i*5+1], ex[i*5+2], ex[i*5+3], ex[i*5+4]);
  return 0;
}
