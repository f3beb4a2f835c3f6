// uwt_frame_fused.cu -- K1+K2 fused: the image pyramid AND the gradient images of every level
// from ONE read of the new frame.
//
// Restates the pyramid loop of System::AddFrame (/root/reference/src/System.cpp:246-251,
// cv::resize x0.5 == (a + b + c + d + 2) >> 2) and Tracker::ApplyGradient
// (src/Tracker.cpp:1127-1143: Scharr x / y as CV_16S with BORDER_REFLECT_101,
// convertScaleAbs, addWeighted(.5, .5) with ties to even) in one kernel:
//
//   * a CTA owns a 128 x 128 tile of level 0.  The tile and a 16-pixel halo (160 x 160 bytes)
//     are staged in shared memory by ONE 2-D tensor copy (cp.async.bulk.tensor.3d on a
//     CUtensorMap of the source frames; SASS: UTMALDG) that completes on an mbarrier;
//     out-of-image parts of the box are zero-filled by the copy engine;
//   * the level-l region the tile needs (tile >> l plus a (16 >> l)-pixel halo: 80^2, 40^2, 20^2,
//     10^2 bytes) is reduced from the level above it while everything stays in shared memory: a
//     1-pixel halo at level 4 is 16 pixels of level 0, which is why the box carries 16;
//   * BORDER_REFLECT_101 is a property of each level's own image border (the level-1 pixel at
//     x = -1 is level-1 pixel 1, not a reduction of reflected level-0 pixels): border tiles write
//     it into every level's halo cells, after which the stencils have no edge cases;
//   * per level: the interior of the region is stored as the pyramid image, the stencil runs on
//     packed 4-pixel groups (five dp4a per pixel, I2IP saturating packs, SWAR ties-to-even
//     blend) and the u8 gradient image is stored; its sum (the candidate threshold needs the
//     mean, Tracker.cpp:1325-1327) is an integer atomic per level, and the last CTA of a frame
//     turns the five sums into the five integer thresholds.
// The int16 gradientX_/gradientY_ planes are not stored (the candidate kernel rebuilds them for
// the selected pixels, uwt_get_gradients for read-back), exactly as with the separate kernels.
#include <cuda.h>

#include <mutex>

#include "uwt_internal.cuh"

namespace uwt {

namespace {

constexpr int kFT = 128;              // level-0 tile
constexpr int kFH = 16;               // level-0 halo = 1 pixel of level 4
constexpr int kF0 = kFT + 2 * kFH;    // 160: region size at level 0
constexpr int kFusedMaxLevels = 5;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// 3-D tiled tensor copy global -> shared (x, y, frame); SASS: UTMALDG.3D
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}

// dp4a with unsigned pixel bytes and signed stencil weights (SASS: IDP.4A.U8.S8)
__device__ __forceinline__ int dp4a_us(uint32_t pix, uint32_t wgt, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pix), "r"(wgt), "r"(acc));
  return d;
}
// min(v, 255) of four non-negative ints packed into 4 bytes, a in the lowest (SASS: 2 x
// I2IP.U8.S32.SAT).  cvt.pack d, x, y, z = (z << 16) | (sat(x) << 8) | sat(y).
__device__ __forceinline__ uint32_t pack_sat_u8x4(int a, int b, int c, int d) {
  uint32_t hi, r;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
  return r;
}

// The four 3-pixel windows of a 4-pixel group: window i = bytes [sx-1+i, sx+2+i] of the row
// (sx a multiple of 4).
struct RowWin {
  uint32_t w[4];
};
__device__ __forceinline__ RowWin row_windows(const uint8_t* srow, int sx) {
  const uint32_t l = *reinterpret_cast<const uint32_t*>(srow + sx - 4);
  const uint32_t c = *reinterpret_cast<const uint32_t*>(srow + sx);
  const uint32_t r = *reinterpret_cast<const uint32_t*>(srow + sx + 4);
  RowWin o;
  o.w[0] = __funnelshift_r(l, c, 24);
  o.w[1] = c;
  o.w[2] = __funnelshift_r(c, r, 8);
  o.w[3] = __funnelshift_r(c, r, 16);
  return o;
}

// Gradient image value of 4 adjacent pixels (Tracker.cpp:1133-1142) from their row windows.
template <bool kSobel>
__device__ __forceinline__ uint32_t gradient4(const RowWin& top, const RowWin& mid,
                                              const RowWin& bot) {
  constexpr StencilWeights sw = kSobel
                                    ? StencilWeights{0x000100FFu, 0x000200FEu, 0x00010201u, 0x00FFFEFFu}
                                    : StencilWeights{0x000300FDu, 0x000A00F6u, 0x00030A03u, 0x00FDF6FDu};
  int vx[4], vy[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // Tracker.cpp:1133-1134: Scharr x / y, CV_16S
    vx[i] = dp4a_us(top.w[i], sw.d, dp4a_us(mid.w[i], sw.dm, dp4a_us(bot.w[i], sw.d, 0)));
    vy[i] = dp4a_us(bot.w[i], sw.sp, dp4a_us(top.w[i], sw.sm, 0));
  }
  // Tracker.cpp:1139-1140: convertScaleAbs -> min(|v|, 255), packed 4 x u8
  const uint32_t ax = pack_sat_u8x4(abs(vx[0]), abs(vx[1]), abs(vx[2]), abs(vx[3]));
  const uint32_t ay = pack_sat_u8x4(abs(vy[0]), abs(vy[1]), abs(vy[2]), abs(vy[3]));
  // Tracker.cpp:1142: addWeighted(.5, .5) = (ax + ay) / 2, ties to even, on 4 bytes at once:
  // floor average, plus one where the sum is odd and the floor is odd
  const uint32_t x_or = ax ^ ay;
  const uint32_t fl = (ax & ay) + ((x_or >> 1) & 0x7F7F7F7Fu);
  return fl + (x_or & fl & 0x01010101u);
}

// One pixel, scalar (levels whose interior does not start on a 4-byte boundary: 3 and 4).
template <bool kSobel>
__device__ __forceinline__ uint32_t gradient1(const uint8_t* c, int pitch) {
  const int a00 = c[-pitch - 1], a01 = c[-pitch], a02 = c[-pitch + 1];
  const int a10 = c[-1], a12 = c[1];
  const int a20 = c[pitch - 1], a21 = c[pitch], a22 = c[pitch + 1];
  constexpr int k0 = kSobel ? 1 : 3, k1 = kSobel ? 2 : 10;
  const int vx = k0 * (a02 - a00) + k1 * (a12 - a10) + k0 * (a22 - a20);
  const int vy = k0 * (a20 - a00) + k1 * (a21 - a01) + k0 * (a22 - a02);
  const int ax = min(abs(vx), 255), ay = min(abs(vy), 255);
  const int s = ax + ay;
  return (uint32_t)((s >> 1) + ((s & 1) & ((s >> 1) & 1)));
}

// 2x2 reduction of a shared-memory region into the next level's region, 4 outputs per item
// (System.cpp:246-251: (a + b + c + d + 2) >> 2), two outputs per register: the even and the odd
// bytes of a word are spread into 16-bit lanes with one byte permute each, so a word's two
// horizontal pair sums are one add, the vertical sum and the rounding constant a 3-input add,
// and no lane can carry into its neighbour (4 * 255 + 2 < 2^16).
__device__ __forceinline__ uint32_t pair_sums(uint32_t x) {
  return __byte_perm(x, 0u, 0x4240) + __byte_perm(x, 0u, 0x4341);  // (b0 + b1) | (b2 + b3) << 16
}
template <int kDSize>
__device__ __forceinline__ void down4(const uint8_t* src, uint8_t* dst, int t) {
  constexpr int kSPitch = 2 * kDSize, kGroups = kDSize / 4;  // kDSize is a multiple of 4
  for (int i = t; i < kDSize * kGroups; i += 256) {
    const int r = i / kGroups, c = (i % kGroups) * 4;
    const uint2 a = *reinterpret_cast<const uint2*>(src + (2 * r) * kSPitch + 2 * c);
    const uint2 b = *reinterpret_cast<const uint2*>(src + (2 * r + 1) * kSPitch + 2 * c);
    const uint32_t s0 = pair_sums(a.x) + pair_sums(b.x) + 0x00020002u;
    const uint32_t s1 = pair_sums(a.y) + pair_sums(b.y) + 0x00020002u;
    // outputs 0, 1 sit in bytes 0, 2 of (s0 >> 2), outputs 2, 3 in bytes 0, 2 of (s1 >> 2)
    *reinterpret_cast<uint32_t*>(dst + r * kDSize + c) = __byte_perm(s0 >> 2, s1 >> 2, 0x6420);
  }
}

// BORDER_REFLECT_101 into the halo cells of one level's region: x = -1 -> 1, x = w -> w - 2
// (columns first, every row of the region), then y = -1 -> 1, y = h -> h - 2 (rows, including
// the patched columns).  `ox`, `oy`: image coordinates of region cell (0, 0).
__device__ __forceinline__ void patch_columns(uint8_t* reg, int pitch, int size, int ox, int w,
                                              int t) {
  const int cl = -1 - ox, cr = w - ox;  // region columns of image x = -1 and x = w
  const bool left = cl >= 0, right = cr < size;
  if (!(left || right)) return;
  for (int r = t; r < size; r += 256) {
    uint8_t* row = reg + r * pitch;
    if (left) row[cl] = row[cl + 2];
    if (right) row[cr] = row[cr - 2];
  }
}
__device__ __forceinline__ void patch_rows(uint8_t* reg, int pitch, int size, int oy, int h,
                                           int t) {
  const int rt = -1 - oy, rb = h - oy;
  const bool top = rt >= 0, bottom = rb < size;
  if (!(top || bottom)) return;
  for (int c = t; c < size; c += 256) {
    if (top) reg[rt * pitch + c] = reg[(rt + 2) * pitch + c];
    if (bottom) reg[rb * pitch + c] = reg[(rb - 2) * pitch + c];
  }
}

struct FusedShared {
  alignas(128) uint8_t s0[kF0 * kF0];          // 160 x 160, written by the tensor copy
  alignas(16) uint8_t s1[(kF0 / 2) * (kF0 / 2)];
  alignas(16) uint8_t s2[(kF0 / 4) * (kF0 / 4)];
  alignas(16) uint8_t s3[(kF0 / 8) * (kF0 / 8)];
  alignas(16) uint8_t s4[(kF0 / 16) * (kF0 / 16) + 12];
  alignas(8) uint64_t bar;
  unsigned int sums[kFusedMaxLevels];
  int is_last;
};

// Stencil + stores of one level whose interior starts on a 4-byte boundary (levels 0, 1, 2).
// A thread owns one 4-pixel group and walks kRows rows down with a sliding 3-row window.
template <int kLevel, bool kSobel>
__device__ __forceinline__ uint32_t level_pass_packed(const uint8_t* reg, const LevelGeom& L,
                                                      uint8_t* img_plane, uint8_t* g_plane,
                                                      int x0l, int y0l, int t, bool store_img) {
  constexpr int kSize = kF0 >> kLevel, kOff = kFH >> kLevel, kTile = kFT >> kLevel;
  constexpr int kGroups = kTile / 4;            // 32, 16, 8
  constexpr int kBands = 256 / kGroups;         // 8, 16, 32
  constexpr int kRows = kTile / kBands;         // 16, 4, 1
  const int grp = t % kGroups, band = t / kGroups;
  const int xg = x0l + grp * 4, yb = y0l + band * kRows;
  uint32_t gsum = 0;
  if (xg >= L.w || yb >= L.h) return 0;
  const int nvalid = min(4, L.w - xg);
  const uint32_t vmask = nvalid == 4 ? 0xFFFFFFFFu : ((1u << (8 * nvalid)) - 1u);
  const int sx = kOff + grp * 4;
  const uint8_t* trow = reg + (kOff + band * kRows - 1) * kSize;  // region row of image row yb-1
  RowWin top = row_windows(trow, sx);
  RowWin mid = row_windows(trow + kSize, sx);
  // running store addresses (one 64-bit add per row instead of a multiply-add per store); rows of
  // the band below the image bottom are cut off by the trip count
  const size_t o0 = (size_t)yb * L.pitch + xg;
  uint8_t* gp = g_plane + o0;
  uint8_t* ip = img_plane + o0;
  const int rows = min(kRows, L.h - yb);
#pragma unroll
  for (int j = 0; j < kRows; ++j) {
    if (j >= rows) break;
    const RowWin bot = row_windows(trow + (j + 2) * kSize, sx);
    const uint32_t gq = gradient4<kSobel>(top, mid, bot) & vmask;
    gsum = __dp4a(gq, 0x01010101u, gsum);
    // the row pitch is a multiple of 16 and xg of 4: a 4-byte store never leaves the row; bytes
    // beyond the image width land in the pitch padding, which must stay zero -> masked
    *reinterpret_cast<uint32_t*>(gp) = gq;
    if (store_img) *reinterpret_cast<uint32_t*>(ip) = mid.w[1] & vmask;
    gp += L.pitch;
    ip += L.pitch;
    top = mid;
    mid = bot;
  }
  return gsum;
}

template <int kLevel, bool kSobel>
__device__ __forceinline__ uint32_t level_pass_scalar(const uint8_t* reg, const LevelGeom& L,
                                                      uint8_t* img_plane, uint8_t* g_plane,
                                                      int x0l, int y0l, int t) {
  constexpr int kSize = kF0 >> kLevel, kOff = kFH >> kLevel, kTile = kFT >> kLevel;
  uint32_t gsum = 0;
  for (int i = t; i < kTile * kTile; i += 256) {
    const int r = i / kTile, c = i % kTile;
    const int x = x0l + c, y = y0l + r;
    if (x < L.w && y < L.h) {
      const uint8_t* cell = reg + (kOff + r) * kSize + kOff + c;
      const uint32_t g = gradient1<kSobel>(cell, kSize);
      gsum += g;
      const size_t o = (size_t)y * L.pitch + x;
      g_plane[o] = (uint8_t)g;
      img_plane[o] = *cell;
    }
  }
  return gsum;
}

template <bool kSobel>
__global__ void __launch_bounds__(256)
frame_fused_kernel(const __grid_constant__ Geom geom, const Pools pools,
                   const int* __restrict__ slots, const __grid_constant__ CUtensorMap src_map) {
  __shared__ FusedShared sh;  // 34 KB static; s0 is 128-byte aligned for the tensor copy
  const int t = threadIdx.x, lane = t & 31;
  const int slot = slots[blockIdx.z];
  const int x0 = blockIdx.x * kFT, y0 = blockIdx.y * kFT;
  const int levels = geom.levels;

  if (t == 0) {
    mbar_init(&sh.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < kFusedMaxLevels) sh.sums[t] = 0u;
  __syncthreads();
  if (t == 0) {
    mbar_expect_tx(&sh.bar, (uint32_t)(kF0 * kF0));
    tma_load_3d(sh.s0, &src_map, x0 - kFH, y0 - kFH, (int)blockIdx.z, &sh.bar);
  }
  mbar_wait(&sh.bar, 0);

  // ---- pyramid regions, all in shared memory (System.cpp:246-251) ----
  if (levels > 1) down4<kF0 / 2>(sh.s0, sh.s1, t);
  __syncthreads();
  if (levels > 2) down4<kF0 / 4>(sh.s1, sh.s2, t);
  __syncthreads();
  if (levels > 3) down4<kF0 / 8>(sh.s2, sh.s3, t);
  __syncthreads();
  if (levels > 4) {
    constexpr int kS = kF0 / 16, kP = kF0 / 8;
    if (t < kS * kS) {
      const int r = t / kS, c = t % kS;
      const uint8_t* s = sh.s3 + (2 * r) * kP + 2 * c;
      sh.s4[r * kS + c] = (uint8_t)((s[0] + s[1] + s[kP] + s[kP + 1] + 2) >> 2);
    }
  }
  // ---- BORDER_REFLECT_101 of every level's own border (only border tiles do any work) ----
  uint8_t* const regs[kFusedMaxLevels] = {sh.s0, sh.s1, sh.s2, sh.s3, sh.s4};
  const bool border = x0 == 0 || y0 == 0 || x0 + kFT >= geom.lv[0].w || y0 + kFT >= geom.lv[0].h;
  if (border) {
    __syncthreads();
#pragma unroll
    for (int l = 0; l < kFusedMaxLevels; ++l)
      if (l < levels)
        patch_columns(regs[l], kF0 >> l, kF0 >> l, (x0 >> l) - (kFH >> l), geom.lv[l].w, t);
    __syncthreads();
#pragma unroll
    for (int l = 0; l < kFusedMaxLevels; ++l)
      if (l < levels)
        patch_rows(regs[l], kF0 >> l, kF0 >> l, (y0 >> l) - (kFH >> l), geom.lv[l].h, t);
  }
  __syncthreads();

  // ---- per level: pyramid image store, stencil, gradient image store, sum ----
  uint8_t* const img = pools.img + (size_t)slot * geom.plane_elems;
  uint8_t* const gpl = pools.g + (size_t)slot * geom.plane_elems;
  uint32_t s[kFusedMaxLevels] = {0, 0, 0, 0, 0};
  {
    const LevelGeom& L = geom.lv[0];
    s[0] = level_pass_packed<0, kSobel>(sh.s0, L, img + L.plane_off, gpl + L.plane_off, x0, y0, t,
                                        true);
  }
  if (levels > 1) {
    const LevelGeom& L = geom.lv[1];
    s[1] = level_pass_packed<1, kSobel>(sh.s1, L, img + L.plane_off, gpl + L.plane_off, x0 >> 1,
                                        y0 >> 1, t, true);
  }
  if (levels > 2) {
    const LevelGeom& L = geom.lv[2];
    s[2] = level_pass_packed<2, kSobel>(sh.s2, L, img + L.plane_off, gpl + L.plane_off, x0 >> 2,
                                        y0 >> 2, t, true);
  }
  if (levels > 3) {
    const LevelGeom& L = geom.lv[3];
    s[3] = level_pass_scalar<3, kSobel>(sh.s3, L, img + L.plane_off, gpl + L.plane_off, x0 >> 3,
                                        y0 >> 3, t);
  }
  if (levels > 4) {
    const LevelGeom& L = geom.lv[4];
    s[4] = level_pass_scalar<4, kSobel>(sh.s4, L, img + L.plane_off, gpl + L.plane_off, x0 >> 4,
                                        y0 >> 4, t);
  }
  // ---- sums of the gradient images (integers: exact and order-free) ----
#pragma unroll
  for (int l = 0; l < kFusedMaxLevels; ++l) {
    const uint32_t v = __reduce_add_sync(0xffffffffu, s[l]);
    if (lane == 0 && v) atomicAdd(&sh.sums[l], v);
  }
  __syncthreads();
  unsigned long long* gsum = pools.gsum + (size_t)slot * kMaxLevels;
  if (t < levels && sh.sums[t]) atomicAdd(&gsum[t], (unsigned long long)sh.sums[t]);
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const uint32_t ctas = gridDim.x * gridDim.y;
    const uint32_t ticket = atomicAdd(&pools.ticket[(size_t)slot * kMaxLevels], 1u);
    sh.is_last = (ticket == ctas - 1u);
  }
  __syncthreads();
  if (sh.is_last && t < levels) {
    __threadfence();
    const LevelGeom& L = geom.lv[t];
    const unsigned long long S = __ldcg(&gsum[t]);
    // Tracker.cpp:1325-1329: thres = mean + GRADIENT_THRESHOLD (float); 8-bit threshold
    // compares against floor(thres)  (ARITHMETIC.md U6)
    const double mean = (double)S / (double)((long long)L.w * L.h);
    const float thres = (float)(mean + geom.gradient_threshold);
    pools.ithr[(size_t)slot * kMaxLevels + t] = (int)floorf(thres);
    gsum[t] = 0ull;  // re-arm for the next frame
    if (t == 0) pools.ticket[(size_t)slot * kMaxLevels] = 0u;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

// (x, y, frame) map of u8 planes with a box_w x box_h x 1 box, out-of-range parts read as zero.
bool encode_u8_map3d(CUtensorMap* out, const uint8_t* base, uint64_t w, uint64_t h, uint64_t n,
                     uint64_t row_stride, uint64_t frame_stride, uint32_t box_w, uint32_t box_h) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  if ((((uintptr_t)base) | row_stride | frame_stride) & 15) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[2] = {(cuuint64_t)row_stride, (cuuint64_t)frame_stride};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides,
             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool frame_fused_supported(const Geom& g) { return g.levels <= kFusedMaxLevels; }

// Returns the number of kernels launched (1), -1 on a launch error, or -2 when this source cannot
// be described by a tensor map (unaligned base / strides): the caller then runs the separate
// pyramid + gradient kernels, which take any layout.
int launch_frame_fused(const Geom& g, const Pools& p, int n, const int* d_slots,
                       const uint8_t* src, size_t row_stride, size_t frame_stride,
                       cudaStream_t st) {
  if (!frame_fused_supported(g) || !p.gsum) return -2;
  const LevelGeom& L0 = g.lv[0];
  if (n == 1 || frame_stride == 0) frame_stride = row_stride * (size_t)L0.h;
  CUtensorMap map;
  if (!encode_u8_map3d(&map, src, L0.w, L0.h, n, row_stride, frame_stride, kF0, kF0)) return -2;
  const bool sobel = g.gradient_op == UWT_GRADIENT_SOBEL;
  dim3 grid((L0.w + kFT - 1) / kFT, (L0.h + kFT - 1) / kFT, n);
  if (sobel)
    frame_fused_kernel<true><<<grid, 256, 0, st>>>(g, p, d_slots, map);
  else
    frame_fused_kernel<false><<<grid, 256, 0, st>>>(g, p, d_slots, map);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace uwt
