// uwt_candidates.cu -- K3: high-gradient candidate selection in the reference's point order.
//
// Restates Tracker::ObtainCandidatePoints (/root/reference/src/Tracker.cpp:1314-1357, mono
// branch): threshold the gradient image at floor(mean + GRADIENT_THRESHOLD) and enumerate
// the selected pixels with x as the OUTER and y as the INNER loop (Tracker.cpp:1334-1335).
//
// Stream compaction in that (column-major) order over row-major images:
//   count   : one warp per (128-column strip, 32-row segment); lane = 4 columns (one 32-bit
//             load per row, 128 B per warp request); per-(column, segment) counts
//   scan    : one CTA per (slot, level): exclusive prefix sum over (column major, segment
//             minor) -> start offset of every (column, segment) run; total = N_l
//   scatter : same decomposition as count; each lane walks its 4 columns down the segment
//             and appends (x, y) -- and, on the levels EstimatePose optimises, the packed
//             8-byte record (x, y, I1, gx, gy) the Gauss-Newton kernel streams.
#include "uwt_internal.cuh"

namespace uwt {

struct WarpItem {
  int lvl, strip, seg;
  bool valid;
};

__device__ __forceinline__ WarpItem locate_item(const Geom& geom, int item) {
  WarpItem w;
  w.valid = false;
  w.lvl = w.strip = w.seg = 0;
  for (int l = 0; l < geom.levels; ++l) {
    const int cnt = geom.lv[l].nstrip * geom.lv[l].nseg;
    if (item < cnt) {
      w.lvl = l;
      w.strip = item % geom.lv[l].nstrip;
      w.seg = item / geom.lv[l].nstrip;
      w.valid = true;
      return w;
    }
    item -= cnt;
  }
  return w;
}

__global__ void __launch_bounds__(128)
cand_count_kernel(const __grid_constant__ Geom geom, const Pools pools,
                  const int* __restrict__ slots) {
  const int lane = threadIdx.x & 31;
  const WarpItem it = locate_item(geom, blockIdx.x * 4 + (threadIdx.x >> 5));
  if (!it.valid) return;
  const int slot = slots[blockIdx.y];
  const LevelGeom& L = geom.lv[it.lvl];
  const int x = it.strip * kStripW + lane * 4;
  if (x >= L.w) return;
  const int nvalid = min(4, L.w - x);
  const int y_lo = it.seg * kSegRows, y_hi = min(y_lo + kSegRows, L.h);
  const uint32_t ithr = (uint32_t)pools.ithr[(size_t)slot * kMaxLevels + it.lvl];
  const uint8_t* g = pools.g + (size_t)slot * geom.plane_elems + L.plane_off + x;
  uint32_t c[4] = {0, 0, 0, 0};
  for (int y = y_lo; y < y_hi; ++y) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(g + (size_t)y * L.pitch);
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] += (((v >> (8 * i)) & 0xFFu) > ithr) ? 1u : 0u;
  }
  uint32_t* cnt = pools.cnt + (size_t)slot * geom.cnt_elems + L.cnt_off;
  for (int i = 0; i < nvalid; ++i) cnt[(size_t)(x + i) * L.nseg + it.seg] = c[i];
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

__global__ void __launch_bounds__(128)
cand_scatter_kernel(const __grid_constant__ Geom geom, const Pools pools,
                    const int* __restrict__ slots) {
  const int lane = threadIdx.x & 31;
  const WarpItem it = locate_item(geom, blockIdx.x * 4 + (threadIdx.x >> 5));
  if (!it.valid) return;
  const int slot = slots[blockIdx.y];
  const LevelGeom& L = geom.lv[it.lvl];
  const int x = it.strip * kStripW + lane * 4;
  if (x >= L.w) return;
  const int nvalid = min(4, L.w - x);
  const int y_lo = it.seg * kSegRows, y_hi = min(y_lo + kSegRows, L.h);
  const uint32_t ithr = (uint32_t)pools.ithr[(size_t)slot * kMaxLevels + it.lvl];
  const size_t pbase = (size_t)slot * geom.plane_elems + L.plane_off + x;
  const uint8_t* g = pools.g + pbase;
  const uint32_t* cnt = pools.cnt + (size_t)slot * geom.cnt_elems + L.cnt_off;
  uint32_t cur[4] = {0, 0, 0, 0};
  for (int i = 0; i < nvalid; ++i) cur[i] = cnt[(size_t)(x + i) * L.nseg + it.seg];
  uint32_t* xy = pools.cand_xy + (size_t)slot * geom.cand_elems + L.cand_off;
  const bool has_rec = L.rec_off >= 0;
  uint64_t* rec = pools.rec + (size_t)slot * geom.rec_elems + (has_rec ? L.rec_off : 0);
  const uint8_t* img = pools.img + pbase;
  const int16_t* gxp = pools.gx + pbase;
  const int16_t* gyp = pools.gy + pbase;
  const bool vec = (nvalid == 4);
  for (int y = y_lo; y < y_hi; ++y) {
    const size_t ro = (size_t)y * L.pitch;
    const uint32_t v = *reinterpret_cast<const uint32_t*>(g + ro);
    uint32_t sel = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nvalid && ((v >> (8 * i)) & 0xFFu) > ithr) sel |= 1u << i;
    if (!sel) continue;
    uint32_t iv = 0;
    int gxs[4] = {0, 0, 0, 0}, gys[4] = {0, 0, 0, 0};
    if (has_rec) {
      iv = *reinterpret_cast<const uint32_t*>(img + ro);
      if (vec) {
        const uint2 a = *reinterpret_cast<const uint2*>(gxp + ro);
        const uint2 b = *reinterpret_cast<const uint2*>(gyp + ro);
        gxs[0] = (short)(a.x & 0xFFFF); gxs[1] = (short)(a.x >> 16);
        gxs[2] = (short)(a.y & 0xFFFF); gxs[3] = (short)(a.y >> 16);
        gys[0] = (short)(b.x & 0xFFFF); gys[1] = (short)(b.x >> 16);
        gys[2] = (short)(b.y & 0xFFFF); gys[3] = (short)(b.y >> 16);
      } else {
        for (int i = 0; i < nvalid; ++i) {
          gxs[i] = gxp[ro + i];
          gys[i] = gyp[ro + i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (sel & (1u << i)) {
        const uint32_t o = cur[i]++;
        xy[o] = (uint32_t)(x + i) | ((uint32_t)y << 16);
        if (has_rec)
          rec[o] = pack_record((uint32_t)(x + i), (uint32_t)y, (iv >> (8 * i)) & 0xFFu, gxs[i],
                               gys[i]);
      }
    }
  }
}

int launch_candidates(const Geom& g, const Pools& p, int n, const int* d_slots, cudaStream_t st) {
  dim3 grid((g.warp_items_total + 3) / 4, n);
  cand_count_kernel<<<grid, 128, 0, st>>>(g, p, d_slots);
  if (cudaGetLastError() != cudaSuccess) return -1;
  cand_scan_kernel<<<dim3(g.levels, n), 1024, 0, st>>>(g, p, d_slots);
  if (cudaGetLastError() != cudaSuccess) return -1;
  cand_scatter_kernel<<<grid, 128, 0, st>>>(g, p, d_slots);
  if (cudaGetLastError() != cudaSuccess) return -1;
  return 3;
}

}  // namespace uwt
