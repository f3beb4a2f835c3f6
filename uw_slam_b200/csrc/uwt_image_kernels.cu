// uwt_image_kernels.cu -- K1 (pyramid) and K2 (Scharr gradients + gradient image).
//
// K1 restates the pyramid loop of System::AddFrame (/root/reference/src/System.cpp:246-251):
//     cv::resize(img[l-1], img[l], Size(), 0.5, 0.5)  ==  (a + b + c + d + 2) >> 2
// for ALL levels in one pass over level 0 (one 64x64 tile per CTA, reduced in shared memory).
//
// K2 restates Tracker::ApplyGradient (src/Tracker.cpp:1127-1143) for all levels in one
// launch: Scharr x / y as int16 with BORDER_REFLECT_101, min(|.|,255) and the 0.5/0.5 blend
// with ties-to-even.  Only the gradient image (u8) is stored: the int16 gradientX_/gradientY_
// values the tracker needs are recomputed from the image for the selected pixels by the
// candidate kernel (bit-identical, same stencil), and full planes are written only when a
// caller reads them back (uwt_get_gradients).  Tiles are staged in shared memory with their halo by the bulk-copy
// engine (cp.async.bulk + mbarrier), one bulk copy per tile row.  The per-level sum of the
// gradient image (needed for the candidate threshold, Tracker.cpp:1325-1327) is reduced in
// the same kernel; the last CTA of a level turns it into the integer threshold.
#include "uwt_internal.cuh"

namespace uwt {

// ----------------------------------------------------------------------------------------
// K1: pyramid
// ----------------------------------------------------------------------------------------
// cv::remap(.., INTER_LINEAR) with fixed-point maps (CameraModel.cpp:101-103, System.cpp:232-234):
// map1 = integer source (x, y), map2 = 5-bit fractions (fy << 5 | fx); integer weights
// (32-a)(32-b)*32 of 2^15, rounded by (v + 2^14) >> 15; samples outside the source read the
// constant border 0 (cv::BORDER_CONSTANT).  Bit-identical to OpenCV's remapBilinear for 8U.
__device__ __forceinline__ uint32_t remap_pixel(const uint8_t* __restrict__ src, size_t row_stride,
                                                int in_w, int in_h, short2 m, uint32_t frac) {
  const int sx = m.x, sy = m.y;
  const int a = frac & 31, b = (frac >> 5) & 31;
  const bool x0 = sx >= 0 && sx < in_w, x1 = sx + 1 >= 0 && sx + 1 < in_w;
  const bool y0 = sy >= 0 && sy < in_h, y1 = sy + 1 >= 0 && sy + 1 < in_h;
  const uint8_t* p = src + (ptrdiff_t)sy * (ptrdiff_t)row_stride + sx;
  const int v00 = (x0 && y0) ? __ldg(p) : 0;
  const int v01 = (x1 && y0) ? __ldg(p + 1) : 0;
  const int v10 = (x0 && y1) ? __ldg(p + row_stride) : 0;
  const int v11 = (x1 && y1) ? __ldg(p + row_stride + 1) : 0;
  const int v = v00 * ((32 - a) * (32 - b) * 32) + v01 * (a * (32 - b) * 32) +
                v10 * ((32 - a) * b * 32) + v11 * (a * b * 32);
  return (uint32_t)min(max((v + (1 << 14)) >> 15, 0), 255);
}

__global__ void __launch_bounds__(256)
pyramid_kernel(const __grid_constant__ Geom geom, const Pools pools, const int* __restrict__ slots,
               const uint8_t* __restrict__ src, size_t row_stride, size_t frame_stride,
               int src_is_slot, int src_aligned, const RemapArgs rm) {
  __shared__ __align__(16) uint8_t s0[kPyrTile][kPyrTile];
  __shared__ __align__(16) uint8_t s1[32][32];
  __shared__ uint8_t s2[16][16];
  __shared__ uint8_t s3[8][8];
  __shared__ uint8_t s4[4][4];
  __shared__ uint8_t s5[2][2];

  const int t = threadIdx.x;
  const int slot = slots[blockIdx.z];
  const int x0 = blockIdx.x * kPyrTile, y0 = blockIdx.y * kPyrTile;
  const LevelGeom& L0 = geom.lv[0];
  uint8_t* plane = pools.img + (size_t)slot * geom.plane_elems;

  if (rm.map1) {
    // ---- level 0 = remap + ROI crop of the distorted source frame (System.cpp:232-235) ----
    // a warp walks 32 consecutive pixels of a row: map reads and source gathers stay coalesced
    const uint8_t* frame = src + (size_t)blockIdx.z * frame_stride;
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
      const int r = j * 4 + (t >> 6), c = t & 63;
      const int gx = x0 + c, gy = y0 + r;
      uint32_t v = 0;
      if (gx < L0.w && gy < L0.h) {
        const size_t mi = (size_t)(gy + rm.roi_y) * rm.map_w + (gx + rm.roi_x);
        v = remap_pixel(frame, row_stride, rm.in_w, rm.in_h, __ldg(&rm.map1[mi]),
                        __ldg(&rm.map2[mi]));
      }
      s0[r][c] = (uint8_t)v;
    }
    __syncthreads();
    const int r = t >> 2, c = (t & 3) * 16;
    const int gx = x0 + c, gy = y0 + r;
    if (gx < L0.w && gy < L0.h)
      *reinterpret_cast<uint4*>(plane + L0.plane_off + (size_t)gy * L0.pitch + gx) =
          *reinterpret_cast<const uint4*>(&s0[r][c]);
  } else {
    // ---- level 0 tile: 256 threads x 16 bytes ----
    const int r = t >> 2, c = (t & 3) * 16;
    const int gx = x0 + c, gy = y0 + r;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (gx < L0.w && gy < L0.h) {  // w is a multiple of 16: a 16-byte group is all in or out
      const uint8_t* sp =
          src_is_slot ? plane + L0.plane_off + (size_t)gy * L0.pitch + gx
                      : src + (size_t)blockIdx.z * frame_stride + (size_t)gy * row_stride + gx;
      if (src_aligned) {
        v = *reinterpret_cast<const uint4*>(sp);
      } else {
        uint32_t w4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w4[i] = sp[4 * i] | (sp[4 * i + 1] << 8) | (sp[4 * i + 2] << 16) | (sp[4 * i + 3] << 24);
        v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
      if (!src_is_slot)
        *reinterpret_cast<uint4*>(plane + L0.plane_off + (size_t)gy * L0.pitch + gx) = v;
    }
    *reinterpret_cast<uint4*>(&s0[r][c]) = v;
  }
  __syncthreads();

  // ---- level 1: 32 x 32, four outputs per thread ----
  if (geom.levels > 1) {
    const LevelGeom& L = geom.lv[1];
    const int r = t >> 3, c = (t & 7) * 4;
    const uint2 a = *reinterpret_cast<const uint2*>(&s0[2 * r][2 * c]);
    const uint2 b = *reinterpret_cast<const uint2*>(&s0[2 * r + 1][2 * c]);
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t aw = (i < 2) ? a.x : a.y, bw = (i < 2) ? b.x : b.y;
      const int sh = (i & 1) * 16;
      const uint32_t s = ((aw >> sh) & 0xFF) + ((aw >> (sh + 8)) & 0xFF) + ((bw >> sh) & 0xFF) +
                         ((bw >> (sh + 8)) & 0xFF) + 2;
      out |= (s >> 2) << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(&s1[r][c]) = out;
    const int gx = (x0 >> 1) + c, gy = (y0 >> 1) + r;
    if (gx < L.w && gy < L.h)  // w1 is a multiple of 8
      *reinterpret_cast<uint32_t*>(plane + L.plane_off + (size_t)gy * L.pitch + gx) = out;
  }
  __syncthreads();

  // ---- levels 2..6: one output per thread ----
#define UWT_DOWN(LVL, SRC, DST, DIM)                                                         \
  if (geom.levels > LVL) {                                                                   \
    if (t < DIM * DIM) {                                                                     \
      const LevelGeom& L = geom.lv[LVL];                                                     \
      const int r = t / DIM, c = t % DIM;                                                    \
      const uint32_t s = SRC[2 * r][2 * c] + SRC[2 * r][2 * c + 1] + SRC[2 * r + 1][2 * c] + \
                         SRC[2 * r + 1][2 * c + 1] + 2;                                      \
      const uint8_t o = (uint8_t)(s >> 2);                                                   \
      DST[r][c] = o;                                                                         \
      const int gx = (x0 >> LVL) + c, gy = (y0 >> LVL) + r;                                  \
      if (gx < L.w && gy < L.h) plane[L.plane_off + (size_t)gy * L.pitch + gx] = o;          \
    }                                                                                        \
    __syncthreads();                                                                         \
  }
  UWT_DOWN(2, s1, s2, 16)
  UWT_DOWN(3, s2, s3, 8)
  UWT_DOWN(4, s3, s4, 4)
  UWT_DOWN(5, s4, s5, 2)
#undef UWT_DOWN
  if (geom.levels > 6 && t == 0) {
    const LevelGeom& L = geom.lv[6];
    const uint32_t s = s5[0][0] + s5[0][1] + s5[1][0] + s5[1][1] + 2;
    const int gx = x0 >> 6, gy = y0 >> 6;
    if (gx < L.w && gy < L.h) plane[L.plane_off + (size_t)gy * L.pitch + gx] = (uint8_t)(s >> 2);
  }
}

int launch_pyramid(const Geom& g, const Pools& p, int n, const int* d_slots, const uint8_t* src,
                   size_t row_stride, size_t frame_stride, bool src_is_slot, cudaStream_t st,
                   const RemapArgs& rm) {
  const LevelGeom& L0 = g.lv[0];
  dim3 grid((L0.w + kPyrTile - 1) / kPyrTile, (L0.h + kPyrTile - 1) / kPyrTile, n);
  const int aligned =
      src_is_slot || ((((uintptr_t)src) | row_stride | frame_stride) & 15) == 0 ? 1 : 0;
  pyramid_kernel<<<grid, 256, 0, st>>>(g, p, d_slots, src, row_stride, frame_stride,
                                       src_is_slot ? 1 : 0, aligned, rm);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Depth pyramid (System.cpp:248-250): cv::resize(depths_[l-1], depths_[l], Size(), 0.5, 0.5) on
// CV_16U = mean of the 2x2 block rounded half-to-even.  One launch per level; a thread makes two
// horizontally adjacent outputs from two 8-byte loads.
__global__ void __launch_bounds__(256)
depth_down_kernel(const __grid_constant__ Geom geom, const Pools pools,
                  const int* __restrict__ slots, int lvl) {
  const LevelGeom& S = geom.lv[lvl - 1];
  const LevelGeom& D = geom.lv[lvl];
  const int slot = slots[blockIdx.z];
  const int x = (blockIdx.x * 64 + (threadIdx.x & 63)) * 2, y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= D.w || y >= D.h) return;
  const uint16_t* src = pools.dep + (size_t)slot * geom.plane_elems + S.plane_off;
  uint16_t* dst = pools.dep + (size_t)slot * geom.plane_elems + D.plane_off;
  // pitch is a multiple of 16 elements and x is even: 8-byte aligned loads of 4 source pixels
  const ushort4 a = *reinterpret_cast<const ushort4*>(src + (size_t)(2 * y) * S.pitch + 2 * x);
  const ushort4 b = *reinterpret_cast<const ushort4*>(src + (size_t)(2 * y + 1) * S.pitch + 2 * x);
  auto mean4 = [](unsigned s) {
    unsigned q = s >> 2;
    const unsigned rem = s & 3u;
    return q + ((rem == 3u || (rem == 2u && (q & 1u))) ? 1u : 0u);
  };
  const unsigned o0 = mean4((unsigned)a.x + a.y + b.x + b.y);
  const unsigned o1 = mean4((unsigned)a.z + a.w + b.z + b.w);
  uint16_t* o = dst + (size_t)y * D.pitch + x;
  if (x + 1 < D.w)
    *reinterpret_cast<uint32_t*>(o) = o0 | (o1 << 16);
  else
    *o = (uint16_t)o0;
}

// Level 0 of the depth pyramid for n slots from frames in DEVICE memory (the caller's, or the
// upload staging buffer): frame i at src + i * frame_stride bytes (0 = every slot gets the same
// frame), rows row_stride bytes apart.  One launch instead of one 2-D copy per slot.
__global__ void __launch_bounds__(256)
depth_import_kernel(const __grid_constant__ Geom geom, const Pools pools,
                    const int* __restrict__ slots, const uint8_t* __restrict__ src,
                    size_t row_stride, size_t frame_stride, int vec) {
  const LevelGeom& L = geom.lv[0];
  const int slot = slots[blockIdx.z];
  const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int x = (blockIdx.x * 64 + (threadIdx.x & 63)) * 8;
  if (x >= L.w || y >= L.h) return;
  const uint8_t* s = src + (size_t)blockIdx.z * frame_stride + (size_t)y * row_stride + 2 * (size_t)x;
  uint16_t* d = pools.dep + (size_t)slot * geom.plane_elems + L.plane_off + (size_t)y * L.pitch + x;
  if (vec) {  // the width is a multiple of 16 pixels: a group of 8 is all in or all out
    *reinterpret_cast<uint4*>(d) = __ldg(reinterpret_cast<const uint4*>(s));
  } else {
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(s);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = s16[i];
  }
}

int launch_depth_import(const Geom& g, const Pools& p, int n, const int* d_slots,
                        const uint8_t* d_src, size_t row_stride, size_t frame_stride,
                        cudaStream_t st) {
  const LevelGeom& L = g.lv[0];
  const int vec = ((((uintptr_t)d_src) | row_stride | frame_stride) & 15) == 0 ? 1 : 0;
  dim3 grid((L.w + 511) / 512, (L.h + 3) / 4, n);
  depth_import_kernel<<<grid, 256, 0, st>>>(g, p, d_slots, d_src, row_stride, frame_stride, vec);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_depth_pyramid(const Geom& g, const Pools& p, int n, const int* d_slots,
                         cudaStream_t st) {
  int k = 0;
  for (int l = 1; l < g.levels; ++l) {
    dim3 grid((g.lv[l].w + 127) / 128, (g.lv[l].h + 3) / 4, n);
    depth_down_kernel<<<grid, 256, 0, st>>>(g, p, d_slots, l);
    if (cudaGetLastError() != cudaSuccess) return -1;
    ++k;
  }
  return k;
}

// CameraModel::Undistort (CameraModel.cpp:101-103): plain remap of one image, no pyramid.
__global__ void __launch_bounds__(256)
remap_kernel(const uint8_t* __restrict__ src, size_t row_stride, int in_w, int in_h,
             const short2* __restrict__ map1, const uint16_t* __restrict__ map2, int out_w,
             int out_h, uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= out_w || y >= out_h) return;
  const size_t mi = (size_t)y * out_w + x;
  dst[mi] = (uint8_t)remap_pixel(src, row_stride, in_w, in_h, __ldg(&map1[mi]), __ldg(&map2[mi]));
}

int launch_remap(const uint8_t* d_src, size_t row_stride, int in_w, int in_h, const short2* map1,
                 const uint16_t* map2, int out_w, int out_h, uint8_t* d_dst, cudaStream_t st) {
  dim3 grid((out_w + 63) / 64, (out_h + 3) / 4);
  remap_kernel<<<grid, 256, 0, st>>>(d_src, row_stride, in_w, in_h, map1, map2, out_w, out_h, d_dst);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ----------------------------------------------------------------------------------------
// K2: gradients
// ----------------------------------------------------------------------------------------
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
// TMA bulk copy global -> shared (SASS: UBLKCP), completion counted on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

constexpr int kHaloX = 16;                               // 16-byte aligned halo
constexpr int kTileRowBytes = kGradTileW + 2 * kHaloX;   // 160
constexpr int kTileRows = kGradTileH + 2;

// dp4a with unsigned pixel bytes and signed stencil weights (SASS: IDP.4A.U8.S8)
__device__ __forceinline__ int dp4a_us(uint32_t pix, uint32_t wgt, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pix), "r"(wgt), "r"(acc));
  return d;
}

// The four 3-pixel windows of a 4-pixel group: window i = bytes [sx-1+i, sx+2+i] of the row.
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

template <bool kPlanes, bool kSobel>
__global__ void __launch_bounds__(256)
gradient_kernel(const __grid_constant__ Geom geom, const Pools pools,
                const int* __restrict__ slots, int tile_begin, int16_t* __restrict__ gx_out,
                int16_t* __restrict__ gy_out) {
  // compile-time stencil rows: Scharr (reference) or Sobel (north-star option)
  constexpr StencilWeights sw = kSobel
                                    ? StencilWeights{0x000100FFu, 0x000200FEu, 0x00010201u, 0x00FFFEFFu}
                                    : StencilWeights{0x000300FDu, 0x000A00F6u, 0x00030A03u, 0x00FDF6FDu};
  __shared__ __align__(128) uint8_t tile[kTileRows][kTileRowBytes];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t warp_sums[8];
  __shared__ int is_last;

  const int t = threadIdx.x;
  const int slot = slots[blockIdx.y];
  // locate (level, tile) from the flattened tile index
  int lvl = 0, tidx = blockIdx.x + tile_begin;
  while (lvl + 1 < geom.levels && tidx >= geom.lv[lvl + 1].tile_off) ++lvl;
  // tile_off is cumulative and increasing with the level; find the last level whose
  // tile_off <= tidx
  tidx -= geom.lv[lvl].tile_off;
  const LevelGeom& L = geom.lv[lvl];
  const int tx = tidx % L.tiles_x, ty = tidx / L.tiles_x;
  const int x0 = tx * kGradTileW, y0 = ty * kGradTileH;
  const size_t plane_base = (size_t)slot * geom.plane_elems + L.plane_off;
  const uint8_t* img = pools.img + plane_base;

  // ---- stage the tile + halo: rows [y0-1, y0+H], bytes [x0-16, x0+W+16) clipped ----
  const int row_lo = max(y0 - 1, 0), row_hi = min(y0 + kGradTileH, L.h - 1);  // inclusive
  const int col_lo = max(x0 - kHaloX, 0), col_hi = min(x0 + kGradTileW + kHaloX, L.pitch);
  const uint32_t row_bytes = (uint32_t)(col_hi - col_lo);
  if (t == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (t == 0) mbar_expect_tx(&bar, row_bytes * (uint32_t)(row_hi - row_lo + 1));
  if (t < kTileRows) {
    const int gy = y0 - 1 + t;
    if (gy >= row_lo && gy <= row_hi)
      bulk_g2s(&tile[t][col_lo - (x0 - kHaloX)], img + (size_t)gy * L.pitch + col_lo, row_bytes,
               &bar);
  }
  mbar_wait(&bar, 0);

  // ---- BORDER_REFLECT_101 written into the halo once, so the stencil loop has no edge cases:
  //      x = -1 -> 1, x = w -> w - 2 (columns first), then y = -1 -> 1, y = h -> h - 2 (rows,
  //      including the patched columns).  Only tiles on the image border take these branches.
  const bool edge_l = (x0 == 0), edge_r = (x0 + kGradTileW >= L.w);
  const bool edge_t = (y0 == 0), edge_b = (y0 + kGradTileH >= L.h);
  if (edge_l || edge_r) {
    if (t < kTileRows) {
      if (edge_l) tile[t][kHaloX - 1] = tile[t][kHaloX + 1];
      if (edge_r) {
        const int cx = kHaloX + (L.w - x0);
        tile[t][cx] = tile[t][cx - 2];
      }
    }
    __syncthreads();
  }
  if (edge_t || edge_b) {
    if (t < kTileRowBytes / 4) {
      uint32_t* rows = reinterpret_cast<uint32_t*>(&tile[0][0]);
      constexpr int kRowWords = kTileRowBytes / 4;
      if (edge_t) rows[t] = rows[2 * kRowWords + t];
      if (edge_b) {
        const int ry = L.h - (y0 - 1);  // tile row of image row h
        rows[ry * kRowWords + t] = rows[(ry - 2) * kRowWords + t];
      }
    }
    __syncthreads();
  }

  // ---- compute: warp = kGradTileH / 8 rows, lane = 4 pixels; 5 dp4a per pixel (3 gx, 2 gy) ----
  const int lane = t & 31, wy = t >> 5;
  const int xg = x0 + lane * 4;
  const int sx = kHaloX + lane * 4;
  uint32_t gsum = 0;
  constexpr int kRowsPerWarp = kGradTileH / 8;
  const int ybase = y0 + wy * kRowsPerWarp;
  if (xg < L.w && ybase < L.h) {
    const int nvalid = min(4, L.w - xg);
    const uint32_t vmask = nvalid == 4 ? 0xFFFFFFFFu : ((1u << (8 * nvalid)) - 1u);
    const uint8_t* trow = &tile[wy * kRowsPerWarp][0];  // tile row of image row ybase - 1
    RowWin top = row_windows(trow, sx);
    RowWin mid = row_windows(trow + kTileRowBytes, sx);
#pragma unroll
    for (int j = 0; j < kRowsPerWarp; ++j) {
      const int y = ybase + j;
      if (y >= L.h) break;
      const RowWin bot = row_windows(trow + (j + 2) * kTileRowBytes, sx);
      int vx[4], vy[4];
      uint32_t ax = 0, ay = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // Tracker.cpp:1133-1134: Scharr x / y, CV_16S
        vx[i] = dp4a_us(top.w[i], sw.d, dp4a_us(mid.w[i], sw.dm, dp4a_us(bot.w[i], sw.d, 0)));
        vy[i] = dp4a_us(bot.w[i], sw.sp, dp4a_us(top.w[i], sw.sm, 0));
        // Tracker.cpp:1139-1140: convertScaleAbs -> min(|v|, 255), packed 4 x u8
        ax |= (uint32_t)min(abs(vx[i]), 255) << (8 * i);
        ay |= (uint32_t)min(abs(vy[i]), 255) << (8 * i);
      }
      // Tracker.cpp:1142: addWeighted(.5, .5) = (ax + ay) / 2, ties to even, on 4 bytes at once:
      // floor average, plus one where the sum is odd and the floor is odd
      const uint32_t x_or = ax ^ ay;
      const uint32_t fl = (ax & ay) + ((x_or >> 1) & 0x7F7F7F7Fu);
      const uint32_t gq = fl + (x_or & fl & 0x01010101u);
      gsum = __dp4a(gq & vmask, 0x01010101u, gsum);
      const size_t o = plane_base + (size_t)y * L.pitch + xg;
      if (nvalid == 4) {
        *reinterpret_cast<uint32_t*>(pools.g + o) = gq;
      } else {
        for (int i = 0; i < nvalid; ++i) pools.g[o + i] = (uint8_t)(gq >> (8 * i));
      }
      if (kPlanes) {
        // gradientX_/gradientY_ planes are materialised only on request (read-back): the
        // tracker itself consumes the gradients through the packed candidate records
        const size_t po = (size_t)L.plane_off + (size_t)y * L.pitch + xg;
        for (int i = 0; i < nvalid; ++i) {
          gx_out[po + i] = (int16_t)vx[i];
          gy_out[po + i] = (int16_t)vy[i];
        }
      }
      top = mid;
      mid = bot;
    }
  }

  // ---- per-tile sum of g, then the last CTA of the level makes the threshold ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
  if (lane == 0) warp_sums[wy] = gsum;
  __syncthreads();
  const int ntiles = L.tiles_x * L.tiles_y;
  uint32_t* gpart = pools.gpart + (size_t)slot * geom.tile_elems + L.tile_off;
  if (t == 0) {
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += warp_sums[i];
    gpart[tidx] = s;
    __threadfence();
    const uint32_t ticket = atomicAdd(&pools.ticket[(size_t)slot * kMaxLevels + lvl], 1u);
    is_last = (ticket == (uint32_t)ntiles - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // fixed-order (deterministic) sum of the per-tile partials; integers, so exact
    unsigned long long s = 0;
    for (int i = t; i < ntiles; i += 256) s += __ldcg(&gpart[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ unsigned long long wsum[8];
    if (lane == 0) wsum[wy] = s;
    __syncthreads();
    if (t == 0) {
      unsigned long long S = 0;
      for (int i = 0; i < 8; ++i) S += wsum[i];
      // Tracker.cpp:1325-1329: thres = mean + GRADIENT_THRESHOLD (float); 8-bit threshold
      // compares against floor(thres)  (ARITHMETIC.md U6)
      const double mean = (double)S / (double)((long long)L.w * L.h);
      const float thres = (float)(mean + geom.gradient_threshold);
      pools.ithr[(size_t)slot * kMaxLevels + lvl] = (int)floorf(thres);
      pools.ticket[(size_t)slot * kMaxLevels + lvl] = 0;  // re-arm for the next frame
    }
  }
}

int launch_gradient(const Geom& g, const Pools& p, int n, const int* d_slots, cudaStream_t st,
                    const LevelRange& lr, int16_t* gx_out, int16_t* gy_out) {
  dim3 grid(lr.tile_count, n);
  const bool sobel = g.gradient_op == UWT_GRADIENT_SOBEL;
  if (gx_out && gy_out) {  // read-back of the int16 planes (not on the hot path)
    if (sobel)
      gradient_kernel<true, true><<<grid, 256, 0, st>>>(g, p, d_slots, lr.tile_begin, gx_out, gy_out);
    else
      gradient_kernel<true, false><<<grid, 256, 0, st>>>(g, p, d_slots, lr.tile_begin, gx_out, gy_out);
  } else if (sobel) {
    gradient_kernel<false, true><<<grid, 256, 0, st>>>(g, p, d_slots, lr.tile_begin, nullptr, nullptr);
  } else {
    gradient_kernel<false, false><<<grid, 256, 0, st>>>(g, p, d_slots, lr.tile_begin, nullptr, nullptr);
  }
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace uwt
