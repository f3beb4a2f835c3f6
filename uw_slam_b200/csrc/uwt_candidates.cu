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

// Loads the (kSegRows x kStripW) u8 tile at (x0, y0) row-wise into shared memory.
__device__ __forceinline__ void load_tile_u8(const uint8_t* __restrict__ plane, const LevelGeom& L,
                                             int x0, int y0, uint8_t* tile, int t) {
  const int lane = t & 31, wid = t >> 5;
  const int gx = x0 + lane * 4;
#pragma unroll
  for (int j = 0; j < kSegRows / 8; ++j) {
    const int row = wid + 8 * j, gy = y0 + row;
    uint32_t v = 0;
    if (gy < L.h && gx < L.pitch) v = *reinterpret_cast<const uint32_t*>(plane + (size_t)gy * L.pitch + gx);
    *reinterpret_cast<uint32_t*>(tile + row * kTilePitch8 + lane * 4) = v;
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
    uint32_t v = 0;
    if (gy >= 0 && gy < L.h && gx >= 0 && gx < L.pitch)
      v = *reinterpret_cast<const uint32_t*>(plane + (size_t)gy * L.pitch + gx);
    *reinterpret_cast<uint32_t*>(tile + row * kImgTilePitch + 4 * j) = v;
  }
}

// Scharr x / y at image pixel (x, y) from the halo tile (Tracker.cpp:1133-1134),
// BORDER_REFLECT_101 by index mapping: the same integers as K2's gradient_kernel.
__device__ __forceinline__ void scharr_from_tile(const uint8_t* tile, const LevelGeom& L, int x0,
                                                 int y0, int x, int y, int& gx, int& gy,
                                                 int& center) {
  const int xm = (x == 0) ? 1 : x - 1, xp = (x + 1 >= L.w) ? L.w - 2 : x + 1;
  const int ym = (y == 0) ? 1 : y - 1, yp = (y + 1 >= L.h) ? L.h - 2 : y + 1;
  const uint8_t* rm = tile + (ym - (y0 - 1)) * kImgTilePitch - (x0 - 4);
  const uint8_t* r0 = tile + (y - (y0 - 1)) * kImgTilePitch - (x0 - 4);
  const uint8_t* rp = tile + (yp - (y0 - 1)) * kImgTilePitch - (x0 - 4);
  const int a = rm[xm], b = rm[x], cc = rm[xp];
  const int d = r0[xm], f = r0[xp];
  const int g = rp[xm], h = rp[x], i = rp[xp];
  center = r0[x];
  gx = 3 * (cc - a) + 10 * (f - d) + 3 * (i - g);
  gy = 3 * (g - a) + 10 * (h - b) + 3 * (i - cc);
}

__global__ void __launch_bounds__(256)
cand_count_kernel(const __grid_constant__ Geom geom, const Pools pools,
                  const int* __restrict__ slots) {
  __shared__ __align__(16) uint8_t sg[kSegRows * kTilePitch8];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const TileItem it = locate_item(geom, blockIdx.x);
  const int slot = slots[blockIdx.y];
  const LevelGeom& L = geom.lv[it.lvl];
  const int x0 = it.strip * kStripW, y0 = it.seg * kSegRows;
  const uint32_t ithr = (uint32_t)pools.ithr[(size_t)slot * kMaxLevels + it.lvl];
  load_tile_u8(pools.g + (size_t)slot * geom.plane_elems + L.plane_off, L, x0, y0, sg, t);
  __syncthreads();
  // warp `wid` owns tile columns [16 wid, 16 wid + 16); lane = row within a 32-row chunk
  uint32_t mine = 0;
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const int c = wid * 16 + j;
    uint32_t cnt = 0;
#pragma unroll
    for (int ch = 0; ch < kSegRows / 32; ++ch) {
      const int row = ch * 32 + lane;
      const bool sel = (y0 + row < L.h) && ((uint32_t)sg[row * kTilePitch8 + c] > ithr);
      cnt += __popc(__ballot_sync(0xffffffffu, sel));
    }
    if (lane == j) mine = cnt;
  }
  const int x = x0 + wid * 16 + lane;
  if (lane < 16 && x < L.w)
    pools.cnt[(size_t)slot * geom.cnt_elems + L.cnt_off + (size_t)x * L.nseg + it.seg] = mine;
}

__global__ void __launch_bounds__(1024)
cand_scan_kernel(const __grid_constant__ Geom geom, const Pools pools,
                 const int* __restrict__ slots) {
  __shared__ uint32_t warp_tot[32];
  const int lvl = blockIdx.x;
  const int slot = slots[blockIdx.y];
  const LevelGeom& L = geom.lv[lvl];
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

__global__ void __launch_bounds__(256)
cand_scatter_kernel(const __grid_constant__ Geom geom, const Pools pools,
                    const int* __restrict__ slots) {
  __shared__ __align__(16) uint8_t sg[kSegRows * kTilePitch8];
  __shared__ __align__(16) uint8_t si[kImgTileRows * kImgTilePitch];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const TileItem it = locate_item(geom, blockIdx.x);
  const int slot = slots[blockIdx.y];
  const LevelGeom& L = geom.lv[it.lvl];
  const int x0 = it.strip * kStripW, y0 = it.seg * kSegRows;
  const uint32_t ithr = (uint32_t)pools.ithr[(size_t)slot * kMaxLevels + it.lvl];
  const size_t pbase = (size_t)slot * geom.plane_elems + L.plane_off;
  const bool has_rec = L.rec_off >= 0;
  load_tile_u8(pools.g + pbase, L, x0, y0, sg, t);
  if (has_rec) load_img_tile_halo(pools.img + pbase, L, x0, y0, si, t);
  __syncthreads();
  const uint32_t* cnt = pools.cnt + (size_t)slot * geom.cnt_elems + L.cnt_off;
  uint32_t* xy = pools.cand_xy + (size_t)slot * geom.cand_elems + L.cand_off;
  uint64_t* rec = pools.rec + (size_t)slot * geom.rec_elems + (has_rec ? L.rec_off : 0);
  const uint32_t lt_mask = (1u << lane) - 1u;
  // start offsets of this warp's 16 (column, segment) runs: one per lane
  uint32_t my_base = 0;
  {
    const int x = x0 + wid * 16 + lane;
    if (lane < 16 && x < L.w) my_base = cnt[(size_t)x * L.nseg + it.seg];
  }
#pragma unroll 2
  for (int j = 0; j < 16; ++j) {
    const int c = wid * 16 + j, x = x0 + c;
    uint32_t base = __shfl_sync(0xffffffffu, my_base, j);
    if (x >= L.w) break;  // warp-uniform
#pragma unroll
    for (int ch = 0; ch < kSegRows / 32; ++ch) {
      const int row = ch * 32 + lane, y = y0 + row;
      const bool sel = (y < L.h) && ((uint32_t)sg[row * kTilePitch8 + c] > ithr);
      const uint32_t b = __ballot_sync(0xffffffffu, sel);
      if (sel) {
        const uint32_t o = base + __popc(b & lt_mask);
        xy[o] = (uint32_t)x | ((uint32_t)y << 16);
        if (has_rec) {
          int gx, gy, i1;
          scharr_from_tile(si, L, x0, y0, x, y, gx, gy, i1);
          rec[o] = pack_record((uint32_t)x, (uint32_t)y, (uint32_t)i1, gx, gy);
        }
      }
      base += __popc(b);
    }
  }
}

int launch_candidates(const Geom& g, const Pools& p, int n, const int* d_slots, cudaStream_t st) {
  dim3 grid(g.warp_items_total, n);
  cand_count_kernel<<<grid, 256, 0, st>>>(g, p, d_slots);
  if (cudaGetLastError() != cudaSuccess) return -1;
  cand_scan_kernel<<<dim3(g.levels, n), 1024, 0, st>>>(g, p, d_slots);
  if (cudaGetLastError() != cudaSuccess) return -1;
  cand_scatter_kernel<<<grid, 256, 0, st>>>(g, p, d_slots);
  if (cudaGetLastError() != cudaSuccess) return -1;
  return 3;
}

}  // namespace uwt
