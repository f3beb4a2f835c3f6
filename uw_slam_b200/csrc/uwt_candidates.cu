// uwt_candidates.cu -- K3: high-gradient candidate selection in the reference's point order.
//
// Restates Tracker::ObtainCandidatePoints (/root/reference/src/Tracker.cpp:1314-1357, mono
// branch): threshold the gradient image at floor(mean + GRADIENT_THRESHOLD) and enumerate
// the selected pixels with x as the OUTER and y as the INNER loop (Tracker.cpp:1334-1335).
//
// Stream compaction in that (column-major) order over row-major images:
//   count   : one CTA per (128-column strip, 64-row segment): the tile of the gradient
//             image is read row-wise (coalesced 128 B per warp request) into shared memory,
//             then every warp walks whole columns of the tile (lane = row, padded pitch =
//             conflict-free), thresholds, ballots and popcounts -> per-(column, segment)
//             counts
//   scan    : one CTA per (slot, level): exclusive prefix sum over (column major, segment
//             minor) -> start offset of every (column, segment) run; total = N_l
//   scatter : same tiles as count; a warp handles one column x 32 rows per step, so the
//             selected points of a column are appended with ONE coalesced store per step
//             (ballot + popc prefix): (x, y) -- and, on the levels EstimatePose optimises,
//             the packed 8-byte record (x, y, I1, gx, gy) the Gauss-Newton kernel streams;
//             I1 and the Scharr gx, gy of a selected pixel come from an image tile with a
//             one-pixel halo in shared memory (no int16 gradient planes are read).
#include <string.h>

#include <algorithm>
#include <mutex>

#include "uwt_internal.cuh"

namespace uwt {

struct TileItem {
  int lvl, strip, seg;
};

__device__ __forceinline__ TileItem locate_item(const Geom& geom, int item) {
  TileItem w;
  w.lvl = w.strip = w.seg = 0;
  for (int l = 0; l < geom.levels; ++l) {
    const int cnt = geom.lv[l].nstrip * geom.lv[l].nseg;
    if (item < cnt) {
      w.lvl = l;
      w.strip = item % geom.lv[l].nstrip;
      w.seg = item / geom.lv[l].nstrip;
      return w;
    }
    item -= cnt;
  }
  return w;
}

constexpr int kTilePitch8 = kStripW + 4;    // bytes per u8 tile row: 33 words -> no conflicts

// 4-byte asynchronous global -> shared copy (LDGSTS); `valid == false` writes zeros.  The count
// and scatter kernels are persistent: while the tile of one work item is processed, the copies
// of the next item are already in flight into the other shared-memory buffer.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Loads the (kSegRows x kStripW) u8 tile at (x0, y0) row-wise into shared memory.
__device__ __forceinline__ void load_tile_u8(const uint8_t* __restrict__ plane, const LevelGeom& L,
                                             int x0, int y0, uint8_t* tile, int t) {
  const int lane = t & 31, wid = t >> 5;
  const int gx = x0 + lane * 4;
#pragma unroll
  for (int j = 0; j < kSegRows / 8; ++j) {
    const int row = wid + 8 * j, gy = y0 + row;
    const bool ok = gy < L.h && gx < L.pitch;
    cp_async4(tile + row * kTilePitch8 + lane * 4, ok ? plane + (size_t)gy * L.pitch + gx : plane,
              ok);
  }
}

// Image tile with a one-pixel halo for the Scharr stencil: rows [y0-1, y0+kSegRows], byte
// columns [x0-4, x0+kStripW+4) (4-byte aligned words); pitch 35 words -> column walks are
// bank-conflict free.
constexpr int kImgTileRows = kSegRows + 2;
constexpr int kImgTileWords = kStripW / 4 + 2;
constexpr int kImgTilePitch = (kImgTileWords + 1) * 4;  // 140 bytes

__device__ __forceinline__ void load_img_tile_halo(const uint8_t* __restrict__ plane,
                                                   const LevelGeom& L, int x0, int y0,
                                                   uint8_t* tile, int t) {
  for (int i = t; i < kImgTileRows * kImgTileWords; i += 256) {
    const int row = i / kImgTileWords, j = i % kImgTileWords;
    const int gy = y0 - 1 + row, gx = x0 - 4 + 4 * j;
    const bool ok = gy >= 0 && gy < L.h && gx >= 0 && gx < L.pitch;
    cp_async4(tile + row * kImgTilePitch + 4 * j, ok ? plane + (size_t)gy * L.pitch + gx : plane,
              ok);
  }
}

// Scharr x / y from the halo tile (Tracker.cpp:1133-1134): the same integers as the frame
// kernel's.  BORDER_REFLECT_101 is already written into the halo (patch_img_tile_borders), so a
// pixel is three 4-byte windows and five dp4a.
__device__ __forceinline__ int dp4a_us(uint32_t pix, uint32_t wgt, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pix), "r"(wgt), "r"(acc));
  return d;
}
// Writes the reflected border pixels (x = -1 -> 1, x = w -> w-2, then y = -1 -> 1, y = h -> h-2)
// into the halo of an image tile.  Block-wide; only tiles on the image border do any work.
__device__ __forceinline__ void patch_img_tile_borders(uint8_t* tile, const LevelGeom& L, int x0,
                                                       int y0, int t) {
  const bool edge_l = (x0 == 0), edge_r = (x0 + kStripW >= L.w);
  const bool edge_t = (y0 == 0), edge_b = (y0 + kSegRows >= L.h);
  if (edge_l || edge_r) {
    if (t < kImgTileRows) {
      uint8_t* row = tile + t * kImgTilePitch;
      if (edge_l) row[3] = row[5];
      if (edge_r) {
        const int cx = L.w - x0 + 4;
        row[cx] = row[cx - 2];
      }
    }
    __syncthreads();
  }
  if (edge_t || edge_b) {
    constexpr int kRowWords = kImgTilePitch / 4;
    if (t < kRowWords) {
      uint32_t* rows = reinterpret_cast<uint32_t*>(tile);
      if (edge_t) rows[t] = rows[2 * kRowWords + t];
      if (edge_b) {
        const int ry = L.h - (y0 - 1);  // tile row of image row h
        rows[ry * kRowWords + t] = rows[(ry - 2) * kRowWords + t];
      }
    }
    __syncthreads();
  }
}

// count: no shared-memory staging.  A thread reads 32-bit words (4 pixels of one row, coalesced
// across the warp), compares the four bytes against the threshold at once (SWAR) and adds the
// 0/1 results into four packed byte counters: a (column, 64-row segment) count is at most 64.
template <bool kDepth>
__global__ void __launch_bounds__(256)
cand_count_kernel(const __grid_constant__ Geom geom, const Pools pools,
                  const int* __restrict__ slots, int n_slots, int item_begin, int item_count) {
  __shared__ uint32_t part[8][32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int total = item_count * n_slots;
  constexpr int kRowsPerWarp = kSegRows / 8;
  for (int work = blockIdx.x; work < total; work += gridDim.x) {
    const TileItem it = locate_item(geom, work % item_count + item_begin);
    const int slot = slots[work / item_count];
    const LevelGeom& L = geom.lv[it.lvl];
    const int x0 = it.strip * kStripW, y0 = it.seg * kSegRows;
    const int ithr = pools.ithr[(size_t)slot * kMaxLevels + it.lvl];
    const uint8_t* plane = pools.g + (size_t)slot * geom.plane_elems + L.plane_off;
    const int gx = x0 + lane * 4;
    uint32_t acc = 0;
    // ObtainAllPoints (UWT_DEPTH_ALL_POINTS): every pixel with depth, no gradient test
    const bool all = kDepth && geom.depth_mode == UWT_DEPTH_ALL_POINTS;
    if ((ithr < 255 || all) && gx < L.pitch) {  // g <= 255: a threshold >= 255 selects nothing
      const uint32_t thr4 = (uint32_t)ithr * 0x01010101u;
      uint32_t w[kRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const int gy = y0 + wid * kRowsPerWarp + r;
        // row-pitch padding beyond the image width is zero (never written), so it never counts
        w[r] = (gy < L.h) ? __ldg(reinterpret_cast<const uint32_t*>(plane + (size_t)gy * L.pitch + gx))
                          : 0u;
      }
      if constexpr (!kDepth) {
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) acc += __vsetgtu4(w[r], thr4);
      } else {
        // depth branch (Tracker.cpp:1339): a pixel counts only if its depth is non-zero too
        const uint16_t* dplane = pools.dep + (size_t)slot * geom.plane_elems + L.plane_off;
        const uint32_t vm = valid4(gx, L.w);
        uint32_t nz[kRowsPerWarp];
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {  // one vector load per row, all in flight
          const int gy = y0 + wid * kRowsPerWarp + r;
          nz[r] = (gy < L.h) ? depth_nz4(dplane + (size_t)gy * L.pitch, gx, geom.depth_mode) : 0u;
        }
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r)
          acc += (all ? 0x01010101u : __vsetgtu4(w[r], thr4)) & nz[r] & vm;
      }
    }
    part[wid][lane] = acc;
    __syncthreads();
    if (wid == 0) {
      uint32_t sum = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += part[k][lane];  // bytes stay <= 64: no carry
      uint32_t* cnt = pools.cnt + (size_t)slot * geom.cnt_elems + L.cnt_off;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int x = gx + i;
        if (x < L.w) cnt[(size_t)x * L.nseg + it.seg] = (sum >> (8 * i)) & 0xFFu;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024)
cand_scan_kernel(const __grid_constant__ Geom geom, const Pools pools,
                 const int* __restrict__ slots, int lvl_begin) {
  __shared__ uint32_t warp_tot[32];
  const int lvl = blockIdx.x + lvl_begin;
  const int slot = slots[blockIdx.y];
  const LevelGeom& L = geom.lv[lvl];
  if (L.rec_off < 0) {
    // a level EstimatePose never optimises: cand_mask_kernel counts it with atomics, start at 0
    if (threadIdx.x == 0) pools.ncand[(size_t)slot * kMaxLevels + lvl] = 0u;
    return;
  }
  const int M = L.w * L.nseg;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  uint32_t* cnt = pools.cnt + (size_t)slot * geom.cnt_elems + L.cnt_off;
  const int per = (M + 1023) / 1024;
  const int lo = min(t * per, M), hi = min(lo + per, M);
  uint32_t s = 0;
  for (int i = lo; i < hi; ++i) s += cnt[i];
  // inclusive scan within the warp, then across warps
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    warp_tot[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  uint32_t run = inc - s + (wid > 0 ? warp_tot[wid - 1] : 0u);
  for (int i = lo; i < hi; ++i) {
    const uint32_t c = cnt[i];
    cnt[i] = run;
    run += c;
  }
  if (t == 1023) pools.ncand[(size_t)slot * kMaxLevels + lvl] = warp_tot[31];
}

// ---- scatter ----------------------------------------------------------------------------
// Tiles arrive by 2-D tensor copies (cp.async.bulk.tensor on per-level maps of the gradient and
// image pools, SASS UTMALDG; parts outside the image read as zero) into double-buffered shared
// memory: the copy of the next tile runs while this one is processed, and no thread spends
// instructions on addresses.
constexpr int kScImgW = kStripW + 32;           // image box: columns [x0 - 16, x0 + 144): the first
                                                // coordinate of a tensor copy is 16-byte aligned
constexpr int kScImgH = kSegRows + 2;           //            rows    [y0 - 1, y0 + 65)
constexpr int kScDerivPitch = kStripW + 4;      // words per row of the derivative tile (16-byte rows)
constexpr uint32_t kScTileBytes = kStripW * kSegRows + kScImgW * kScImgH;
struct ScatterShared {
  alignas(128) uint8_t sg[2][kStripW * kSegRows];      // gradient-image tile, pitch 128
  alignas(128) uint8_t si[2][kScImgW * kScImgH + 64];  // image tile + halo, pitch 160 (size % 128 = 0)
  alignas(16) uint32_t sd[kSegRows * kScDerivPitch];   // gx:13 | gy:13 << 13 of every pixel
  alignas(16) uint32_t cm[(kSegRows / 8) * (kStripW / 4)];  // [band][column group]: byte i = 8-row
                                                            // selection mask of column 4*group + i
  alignas(16) int4 info[2];  // {level, x0, y0, slot} of the tile in each buffer
  int seg[2];
  alignas(8) uint64_t bar[2];
};
struct ScatterMaps {
  CUtensorMap g[kMaxLevels];
  CUtensorMap img[kMaxLevels];
};

__device__ __forceinline__ uint32_t sc_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void sc_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(sc_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void sc_tma_load_3d(void* dst, const CUtensorMap* map, int x, int y,
                                               int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(sc_smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(sc_smem_u32(bar))
      : "memory");
}

// BORDER_REFLECT_101 into the halo of the image tile (x = -1 -> 1, x = w -> w - 2, then
// y = -1 -> 1, y = h -> h - 2 including the patched columns).  Only border tiles do any work.
__device__ __forceinline__ void sc_patch_borders(uint8_t* tile, const LevelGeom& L, int x0, int y0,
                                                 int t) {
  const bool edge_l = (x0 == 0), edge_r = (x0 + kStripW >= L.w);
  const bool edge_t = (y0 == 0), edge_b = (y0 + kSegRows >= L.h);
  if (edge_l || edge_r) {
    if (t < kScImgH) {
      uint8_t* row = tile + t * kScImgW;
      if (edge_l) row[15] = row[17];
      if (edge_r) {
        const int cx = L.w - x0 + 16;
        row[cx] = row[cx - 2];
      }
    }
    __syncthreads();
  }
  if (edge_t || edge_b) {
    constexpr int kRowWords = kScImgW / 4;
    if (t < kRowWords) {
      uint32_t* rows = reinterpret_cast<uint32_t*>(tile);
      if (edge_t) rows[t] = rows[2 * kRowWords + t];
      if (edge_b) {
        const int ry = L.h - (y0 - 1);  // tile row of image row h
        rows[ry * kRowWords + t] = rows[(ry - 2) * kRowWords + t];
      }
    }
    __syncthreads();
  }
}

// Two phases per 128 x 64 tile, both free of shared-memory bank conflicts:
//   dense  : warp = 8-row band, lane = 4-pixel column group.  A thread slides a 3-row window
//            down its band (three shared loads per 4 pixels), evaluates Scharr x / y
//            (Tracker.cpp:1133-1134; the frame kernel's integers) for every pixel, stores
//            gx | gy << 13 -- the upper word of the packed record -- with one 16-byte store per
//            row, and thresholds the gradient image four pixels at a time: `acc |= sel4 << row`
//            leaves the 8-row selection mask of each of its 4 columns in one byte of a register;
//   column : a warp walks 4 columns x 8 rows per step (lane = (row, column): the derivative
//            tile's 132-word pitch spreads these over all 32 banks).  Every lane keeps the
//            running output offset of its column (x outer, y inner: Tracker.cpp:1334-1335), so
//            a selected pixel's slot is base + popc(mask below its row): no ballots, and a
//            record is two shared loads, three logic operations and one 8-byte store.
// The first form of this kernel evaluated the stencil inside the divergent branch of a
// column walk and copied its tiles with 4-byte cp.async: 19.7 k warp instructions per tile.
template <bool kDepth>
__global__ void __launch_bounds__(256)
cand_scatter_kernel(const __grid_constant__ Geom geom, const Pools pools,
                    const int* __restrict__ slots, int n_slots, int item_begin, int item_count,
                    const StencilWeights sw, const __grid_constant__ ScatterMaps maps) {
  extern __shared__ __align__(128) uint8_t scatter_smem[];
  ScatterShared& sh = *reinterpret_cast<ScatterShared*>(scatter_smem);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int total = item_count * n_slots;
  const int cq = lane & 3, rq = lane >> 2;  // column phase: lane = (row rq, column cq) of a 4 x 8 step
  // Thread 0 decodes a work item once (level, tile origin, slot), starts its tensor copies and
  // leaves the result in shared memory; everybody else reads it after the next barrier.
  auto issue = [&](int work, int buf) {  // thread 0 only
    const TileItem it = locate_item(geom, work % item_count + item_begin);
    const int slot = slots[work / item_count];
    const int x0 = it.strip * kStripW, y0 = it.seg * kSegRows;
    sh.info[buf] = make_int4(it.lvl, x0, y0, slot);
    sh.seg[buf] = it.seg;
    // the buffers were read (and the image tile patched) through the generic proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     sc_smem_u32(&sh.bar[buf])),
                 "r"(kScTileBytes)
                 : "memory");
    sc_tma_load_3d(sh.sg[buf], &maps.g[it.lvl], x0, y0, slot, &sh.bar[buf]);
    sc_tma_load_3d(sh.si[buf], &maps.img[it.lvl], x0 - 16, y0 - 1, slot, &sh.bar[buf]);
  };
  // Per-tile values that come from global memory -- the integer threshold and the start offsets
  // of this lane's four (column, segment) runs -- are requested one tile ahead.
  struct TileRegs {
    int lvl, x0, y0, slot;
    uint32_t ithr, base[4];
  };
  auto fetch = [&](int buf) {
    TileRegs r;
    const int4 in = sh.info[buf];
    r.lvl = in.x, r.x0 = in.y, r.y0 = in.z, r.slot = in.w;
    const LevelGeom& L = geom.lv[r.lvl];
    r.ithr = (uint32_t)__ldg(&pools.ithr[(size_t)r.slot * kMaxLevels + r.lvl]);
    const uint32_t* cnt = pools.cnt + (size_t)r.slot * geom.cnt_elems + L.cnt_off;
    const int seg = sh.seg[buf];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int x = r.x0 + (wid * 4 + q) * 4 + cq;
      r.base[q] = (x < L.w) ? __ldg(&cnt[(size_t)x * L.nseg + seg]) : 0u;
    }
    return r;
  };
  int work = blockIdx.x, buf = 0;
  uint32_t phases = 0u;  // bit b: parity the next wait on bar[b] expects
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sc_smem_u32(&sh.bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sc_smem_u32(&sh.bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (work < total) issue(work, 0);
  }
  __syncthreads();
  if (work >= total) return;
  TileRegs cur = fetch(0), nxt = cur;
  for (; work < total; work += gridDim.x, buf ^= 1) {
    const int next = work + gridDim.x;
    if (next < total && t == 0) issue(next, buf ^ 1);
    sc_mbar_wait(&sh.bar[buf], (phases >> buf) & 1u);
    phases ^= 1u << buf;
    const LevelGeom& L = geom.lv[cur.lvl];
    const int x0 = cur.x0, y0 = cur.y0, slot = cur.slot;
    const uint32_t ithr = cur.ithr;
    const uint8_t* tg = sh.sg[buf];
    uint8_t* ti = sh.si[buf];
    uint64_t* rec = pools.rec + (size_t)slot * geom.rec_elems + L.rec_off;
    asm volatile("" : "+l"(rec));  // keep base + 32-bit index addressing (one IMAD.WIDE per store)
    const bool all_points = kDepth && geom.depth_mode == UWT_DEPTH_ALL_POINTS;
    const uint16_t* dplane = kDepth ? pools.dep + (size_t)slot * geom.plane_elems + L.plane_off
                                    : nullptr;
    uint16_t* recz = kDepth ? pools.recz + (size_t)slot * geom.rec_elems + L.rec_off : nullptr;
    sc_patch_borders(ti, L, x0, y0, t);
    // ---- dense phase ----
    {
      const int band = wid, xg = x0 + 4 * lane;
      uint32_t acc = 0;
      if (xg < L.w && y0 + band * 8 < L.h) {
        // image-tile row of image row y0 + band * 8 - 1; words: [3 + lane] left of the group,
        // [4 + lane] the group, [5 + lane] right of it
        const uint32_t* ir = reinterpret_cast<const uint32_t*>(ti + (band * 8) * kScImgW) + lane + 3;
        const uint32_t* gr = reinterpret_cast<const uint32_t*>(tg + (band * 8) * kStripW) + lane;
        uint4* out = reinterpret_cast<uint4*>(sh.sd + (band * 8) * kScDerivPitch + 4 * lane);
        constexpr int kIW = kScImgW / 4;
        uint32_t tw[4], mw[4];
        {
          const uint32_t l = ir[0], c = ir[1], r = ir[2];
          tw[0] = __funnelshift_r(l, c, 24), tw[1] = c, tw[2] = __funnelshift_r(c, r, 8),
          tw[3] = __funnelshift_r(c, r, 16);
          const uint32_t l1 = ir[kIW], c1 = ir[kIW + 1], r1 = ir[kIW + 2];
          mw[0] = __funnelshift_r(l1, c1, 24), mw[1] = c1, mw[2] = __funnelshift_r(c1, r1, 8),
          mw[3] = __funnelshift_r(c1, r1, 16);
        }
        const uint32_t thr4 = min(ithr, 255u) * 0x01010101u;
        // Tracker.cpp:1339: depth != 0 as well.  The depth test of the band's 8 rows is requested
        // up front (8 independent vector loads, one latency) and folded into the selection mask,
        // so the column phase is the same as without depth.
        uint32_t nzr[8];
        if constexpr (kDepth) {
          // rows below the image bottom are clamped (their selection bits are cleared anyway);
          // the mode is tested once, outside the loads
          const uint32_t r0 = (uint32_t)(y0 + band * 8), rmax = (uint32_t)(L.h - 1);
          if (geom.depth_mode == UWT_DEPTH_REFERENCE) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              nzr[j] = depth_nz4(dplane + min(r0 + j, rmax) * (uint32_t)L.pitch, xg,
                                 UWT_DEPTH_REFERENCE);
          } else if (geom.depth_mode == UWT_DEPTH_ALL_POINTS) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              nzr[j] = depth_nz4(dplane + min(r0 + j, rmax) * (uint32_t)L.pitch, xg,
                                 UWT_DEPTH_ALL_POINTS);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              nzr[j] = depth_nz4(dplane + min(r0 + j, rmax) * (uint32_t)L.pitch, xg,
                                 UWT_DEPTH_U16);
          }
        }
        // bytes of the group beyond the image width never count (all-points mode has no
        // gradient test; elsewhere their gradient is zero anyway)
        const int nvalid = min(4, L.w - xg);
        const uint32_t vmask = nvalid == 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - nvalid)));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t l = ir[(j + 2) * kIW], c = ir[(j + 2) * kIW + 1], r = ir[(j + 2) * kIW + 2];
          const uint32_t bw[4] = {__funnelshift_r(l, c, 24), c, __funnelshift_r(c, r, 8),
                                  __funnelshift_r(c, r, 16)};
          uint32_t d[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int gx = dp4a_us(tw[i], sw.d, dp4a_us(mw[i], sw.dm, dp4a_us(bw[i], sw.d, 0)));
            const int gy = dp4a_us(bw[i], sw.sp, dp4a_us(tw[i], sw.sm, 0));
            d[i] = ((uint32_t)gx & 0x1FFFu) | (((uint32_t)gy & 0x1FFFu) << 13);
            tw[i] = mw[i];
            mw[i] = bw[i];
          }
          out[j * (kScDerivPitch / 4)] = make_uint4(d[0], d[1], d[2], d[3]);
          uint32_t sel4 = all_points ? 0x01010101u : __vsetgtu4(gr[j * (kStripW / 4)], thr4);
          if (y0 + band * 8 + j >= L.h) sel4 = 0u;
          if constexpr (kDepth) sel4 &= nzr[j];
          acc |= (sel4 & vmask) << j;
        }
      }
      sh.cm[band * (kStripW / 4) + lane] = acc;
    }
    __syncthreads();
    if (next < total) nxt = fetch(buf ^ 1);  // info[buf ^ 1] was written before the barrier above
    // ---- column phase: 4 columns x 8 rows per step ----
    {
      const uint32_t below = (1u << rq) - 1u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cg = wid * 4 + q;           // column group of this step
        const int c = cg * 4 + cq, x = x0 + c;
        if (x0 + cg * 4 >= L.w) break;        // warp-uniform
        uint32_t base = cur.base[q];
        const uint8_t* cmb = reinterpret_cast<const uint8_t*>(sh.cm) + cg * 4 + cq;
        const uint32_t* dcol = sh.sd + rq * kScDerivPitch + c;
        const uint8_t* icol = ti + (rq + 1) * kScImgW + 16 + c;
        uint32_t xy = (uint32_t)x | ((uint32_t)(y0 + rq) << 12);
        // depth of a record (depth modes): this lane's column of the depth plane.  REFERENCE mode
        // reads byte x of the 16-bit row (Tracker.cpp:1344 at<uchar>); a selected pixel has passed
        // the depth test, so ALL_POINTS needs no sign check here.
        const uint16_t* dlane = nullptr;
        uint32_t dshift = 0, dmask = 0xFFFFu;
        if constexpr (kDepth) {
          const int xx = min(x, L.w - 1);
          const bool ref = geom.depth_mode == UWT_DEPTH_REFERENCE;
          dlane = dplane + (ref ? (xx >> 1) : xx);
          dshift = (ref && (xx & 1)) ? 8u : 0u;
          dmask = ref ? 0xFFu : 0xFFFFu;
        }
#pragma unroll
        for (int b = 0; b < kSegRows / 8; ++b) {
          const uint32_t m8 = cmb[b * kStripW];
          uint32_t sel = (m8 >> rq) & 1u;
          uint32_t o = base + __popc(m8 & below);
          base += __popc(m8);
          if constexpr (kDepth) {
            // the integer depth of a selected pixel travels next to its record: loaded by every
            // lane (clamped into the image, so the 8 loads of a column group pipeline), stored by
            // the selected ones
            const uint32_t yy = min((uint32_t)(y0 + b * 8 + rq), (uint32_t)(L.h - 1));
            const uint32_t dz = ((uint32_t)__ldg(dlane + yy * (uint32_t)L.pitch) >> dshift) & dmask;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.u32 p, %0, 0;\n"
                "@p st.global.u16 [%1], %2;\n"
                "}\n" ::"r"(sel),
                "l"(recz + o), "h"((unsigned short)dz)
                : "memory");
          }
          // branch-free: every lane assembles a record, the selected ones store it
          const uint32_t i1 = icol[b * 8 * kScImgW];
          const uint32_t hi = dcol[b * 8 * kScDerivPitch];
          const uint32_t lo = xy | (i1 << 24);
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.u32 p, %0, 0;\n"
              "@p st.global.v2.u32 [%1], {%2, %3};\n"
              "}\n" ::"r"(sel),
              "l"(rec + o), "r"(lo), "r"(hi)
              : "memory");
          xy += 8u << 12;
        }
      }
    }
    cur = nxt;
    __syncthreads();  // both buffers of this stage are refilled in the next iteration
  }
}

// Levels EstimatePose never optimises (level 0 by default: 75 % of all pixels; Tracker.cpp:389 only
// walks first_level..last_level): ObtainCandidatePoints' selection is kept as a BITMASK plus the
// count N_l.  Same threshold, same selected set; the x-major N x 4 list of candidatePoints_[l] is
// expanded from the mask when an accessor asks for it (uwt_get_candidates), exactly like the float
// matrix of the optimised levels is expanded from the packed records.  One thread = one mask
// word = 32 pixels: two 16-byte loads, eight SWAR compares, a multiply that gathers the four 0/1
// bytes of a compare into a nibble.
template <bool kDepth>
__global__ void __launch_bounds__(256)
cand_mask_kernel(const __grid_constant__ Geom geom, const Pools pools,
                 const int* __restrict__ slots, int lvl_lo, int lvl_hi) {
  __shared__ uint32_t wsum[8];
  const int slot = slots[blockIdx.y];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  // a CTA stays inside one level: per-level CTA ranges are computed from the level sizes
  int lvl = -1, first_word = 0;
  {
    int cta = blockIdx.x;
    for (int l = lvl_lo; l <= lvl_hi; ++l) {
      if (geom.lv[l].rec_off >= 0) continue;
      const int ctas = (geom.lv[l].mask_wpr * geom.lv[l].h + 255) / 256;
      if (cta < ctas) {
        lvl = l;
        first_word = cta * 256;
        break;
      }
      cta -= ctas;
    }
  }
  if (lvl < 0) return;
  const LevelGeom& L = geom.lv[lvl];
  const int idx = first_word + t;
  uint32_t word = 0;
  if (idx < L.mask_wpr * L.h) {
    const int row = idx / L.mask_wpr, x0 = (idx % L.mask_wpr) * 32;
    const int ithr = pools.ithr[(size_t)slot * kMaxLevels + lvl];
    const bool all = kDepth && geom.depth_mode == UWT_DEPTH_ALL_POINTS;
    if (ithr < 255 || all) {
      const uint32_t thr4 = (uint32_t)ithr * 0x01010101u;
      const uint8_t* grow = pools.g + (size_t)slot * geom.plane_elems + L.plane_off +
                            (size_t)row * L.pitch + x0;
      uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
      // the pitch is a multiple of 16 and its padding is zero: 16-byte groups are all-or-nothing
      if (x0 < L.pitch) a = __ldg(reinterpret_cast<const uint4*>(grow));
      if (x0 + 16 < L.pitch) b = __ldg(reinterpret_cast<const uint4*>(grow + 16));
      const uint32_t w8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t m = all ? 0x01010101u : __vsetgtu4(w8[j], thr4);  // 0 / 1 per byte
        if constexpr (kDepth) {  // Tracker.cpp:1339: depth != 0 as well, 4 pixels per load
          const uint16_t* drow = pools.dep + (size_t)slot * geom.plane_elems + L.plane_off +
                                 (size_t)row * L.pitch;
          const int gx = x0 + 4 * j;
          m &= valid4(gx, L.w);
          if (m) m &= depth_nz4(drow, gx, geom.depth_mode);
        }
        word |= ((m * 0x01020408u) >> 24) << (4 * j);          // -> 4 bits
      }
    }
    pools.sel_mask[(size_t)slot * geom.mask_elems + L.mask_off + idx] = word;
  }
  const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(word));
  if (lane == 0) wsum[wid] = c;
  __syncthreads();
  if (t == 0) {
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += wsum[i];
    if (s) atomicAdd(&pools.ncand[(size_t)slot * kMaxLevels + lvl], s);  // integer: order-free
  }
}

int launch_candidates(const Geom& g, const Pools& p, int n, const int* d_slots, cudaStream_t st,
                      const LevelRange& lr_all, bool with_mask_levels) {
  // Ordered stream compaction (count -> scan -> scatter) runs on the levels EstimatePose
  // optimises; the other levels of the requested range keep a selection bitmask + count.
  const LevelRange lr = level_range(g, g.last_level, g.first_level);
  // persistent grids: a multiple of the 148 SMs, as many CTAs per SM as the double-buffered
  // tiles allow; small jobs get one CTA per work item
  const long long total = (long long)lr.item_count * n;
  // occupancy is the same for every device of a box: computed once, published together (handles
  // may be driven from different host threads); the dynamic shared-memory limit of the scatter
  // kernel is a per-device function attribute and is raised once per device
  static int per_sm_count = 0, per_sm_scatter = 0, sms = 0;
  static std::mutex attr_mutex;
  static bool attr_set[64] = {};
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(attr_mutex);
    bool& set = attr_set[(dev < 0 ? 0 : dev) % 64];
    if (!set) {
      if (cudaFuncSetAttribute(cand_scatter_kernel<false>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(ScatterShared)) != cudaSuccess ||
          cudaFuncSetAttribute(cand_scatter_kernel<true>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(ScatterShared)) != cudaSuccess)
        return -1;
      set = true;
    }
    if (sms == 0) {
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_count, cand_count_kernel<false>, 256,
                                                    0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_scatter, cand_scatter_kernel<false>,
                                                    256, sizeof(ScatterShared));
      if (per_sm_count <= 0) per_sm_count = 4;
      if (per_sm_scatter <= 0) per_sm_scatter = 3;
      if (sms <= 0) sms = 148;
    }
  }
  const int grid_count = (int)std::min<long long>(total, (long long)sms * per_sm_count);
  const int grid_scatter = (int)std::min<long long>(total, (long long)sms * per_sm_scatter);
  const bool depth = g.depth_mode != UWT_DEPTH_NONE;
  int launches = 0;
  if (depth)
    cand_count_kernel<true><<<grid_count, 256, 0, st>>>(g, p, d_slots, n, lr.item_begin,
                                                        lr.item_count);
  else
    cand_count_kernel<false><<<grid_count, 256, 0, st>>>(g, p, d_slots, n, lr.item_begin,
                                                         lr.item_count);
  if (cudaGetLastError() != cudaSuccess) return -1;
  ++launches;
  // scan: one CTA per (slot, level); on a bitmask level it only resets the count
  const LevelRange& sr = with_mask_levels ? lr_all : lr;
  cand_scan_kernel<<<dim3(sr.lvl_count, n), 1024, 0, st>>>(g, p, d_slots, sr.lvl_begin);
  if (cudaGetLastError() != cudaSuccess) return -1;
  ++launches;
  ScatterMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int l = lr.lvl_begin; l < lr.lvl_begin + lr.lvl_count; ++l) {
    const LevelGeom& L = g.lv[l];
    if (!encode_u8_map3d(&maps.g[l], p.g + L.plane_off, L.pitch, L.h, g.max_slots, L.pitch,
                         g.plane_elems, kStripW, kSegRows) ||
        !encode_u8_map3d(&maps.img[l], p.img + L.plane_off, L.pitch, L.h, g.max_slots, L.pitch,
                         g.plane_elems, kScImgW, kScImgH))
      return -1;
  }
  if (depth)
    cand_scatter_kernel<true><<<grid_scatter, 256, sizeof(ScatterShared), st>>>(
        g, p, d_slots, n, lr.item_begin, lr.item_count, stencil_weights(g.gradient_op), maps);
  else
    cand_scatter_kernel<false><<<grid_scatter, 256, sizeof(ScatterShared), st>>>(
        g, p, d_slots, n, lr.item_begin, lr.item_count, stencil_weights(g.gradient_op), maps);
  if (cudaGetLastError() != cudaSuccess) return -1;
  ++launches;
  if (with_mask_levels) {
    const int lo = lr_all.lvl_begin, hi = lr_all.lvl_begin + lr_all.lvl_count - 1;
    int ctas = 0;
    for (int l = lo; l <= hi; ++l)
      if (g.lv[l].rec_off < 0) ctas += (g.lv[l].mask_wpr * g.lv[l].h + 255) / 256;
    if (ctas > 0) {
      if (depth)
        cand_mask_kernel<true><<<dim3(ctas, n), 256, 0, st>>>(g, p, d_slots, lo, hi);
      else
        cand_mask_kernel<false><<<dim3(ctas, n), 256, 0, st>>>(g, p, d_slots, lo, hi);
      if (cudaGetLastError() != cudaSuccess) return -1;
      ++launches;
    }
  }
  return launches;
}

}  // namespace uwt
