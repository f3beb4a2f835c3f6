// uwt_internal.cuh -- shared declarations of libuwtrack (sm_100a only).
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through the runtime)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/uwtrack.h"

namespace uwt {

constexpr int kMaxLevels = UWT_MAX_LEVELS;

// Gradient tile (K2): 128 x kGradTileH pixels, staged with a 16-byte-aligned halo by
// cp.async.bulk; 8 warps x (kGradTileH / 8) rows.
constexpr int kGradTileW = 128;
#ifndef UWT_GRAD_TILE_H
#define UWT_GRAD_TILE_H 128
#endif
constexpr int kGradTileH = UWT_GRAD_TILE_H;
// Candidate compaction (K3): one CTA owns a 128-column strip x kSegRows rows.
constexpr int kStripW = 128;
#ifndef UWT_SEG_ROWS
#define UWT_SEG_ROWS 64
#endif
constexpr int kSegRows = UWT_SEG_ROWS;  // <= 255: the count kernel packs per-column counts in bytes
// Pyramid tile (K1): 64 x 64 level-0 pixels -> 32x32, 16x16, 8x8, 4x4, 2x2, 1x1.
constexpr int kPyrTile = 64;

// Packed candidate record consumed by the Gauss-Newton kernel (8 bytes / point):
//   bits  0..11 x, 12..23 y, 24..31 I1, 32..44 gx (13-bit two's complement), 45..57 gy.
// |gx|,|gy| <= 16*255 = 4080 < 4096 (Scharr weights sum to 16).
__host__ __device__ inline uint64_t pack_record(uint32_t x, uint32_t y, uint32_t i1, int gx,
                                                int gy) {
  return (uint64_t)(x & 0xFFFu) | ((uint64_t)(y & 0xFFFu) << 12) | ((uint64_t)(i1 & 0xFFu) << 24) |
         ((uint64_t)((uint32_t)gx & 0x1FFFu) << 32) | ((uint64_t)((uint32_t)gy & 0x1FFFu) << 45);
}

struct LevelGeom {
  int w, h, pitch;     // pitch in elements, multiple of 16
  int plane_off;       // element offset of this level inside a slot's image/gradient planes
  int mask_off;        // word offset of this level's selection bitmask, -1 if it has records
  int mask_wpr;        // 32-bit words per bitmask row = ceil(w / 32)
  int rec_off;         // element offset inside a slot's record list, -1 if not optimised
  int nstrip, nseg;    // compaction decomposition
  int cnt_off;         // offset inside a slot's per-(column,segment) count array
  int tiles_x, tiles_y;
  int tile_off;        // offset inside a slot's per-tile gradient partial sums
  float fx, fy, cx, cy, invfx, invfy;
  float wf, hf;        // (float)w, (float)h: bounds of the validity test (Tracker.cpp:450)
  int wm1, hm1;        // w - 1, h - 1: clamp of the nearest-pixel index (ARITHMETIC.md U1)
};

struct Geom {
  LevelGeom lv[kMaxLevels];
  int levels, first_level, last_level, max_iterations;
  int max_slots;      // slots per pool (uwt_config.max_frames)
  float epsilon, residual_scale;
  double gradient_threshold;
  int solve_mode;
  float lm_lambda;    // UWT_SOLVE_CHOLESKY_LM damping
  int weight_mode;    // UWT_WEIGHT_*
  float huber_delta;
  int depth_mode;     // UWT_DEPTH_*
  int gradient_op;    // UWT_GRADIENT_*
  int sampling;       // UWT_SAMPLE_*
  int residual_scale_int, residual_scale_is_int;  // residual_scale as an integer, if it is one
  int exact_div;      // 1: a principal point of an optimised level is (nearly) 0 -> the residual
                      // sweep uses the generic IEEE division for every point (uwt_estimate.cu)
  // per-slot strides (elements)
  size_t plane_elems, mask_elems, rec_elems, cnt_elems, tile_elems;
  int mask_words_total;  // bitmask words per slot over the levels without records
  int grad_tiles_total;  // gradient tiles per slot over all levels
  int warp_items_total;  // compaction tiles per slot over all levels
};

// Device memory pools, indexed [slot].
struct Pools {
  uint8_t* img;         // [slot][plane_elems]
  uint8_t* g;           // [slot][plane_elems]
  uint32_t* gpart;      // [slot][tile_elems]   per-tile sums of g
  uint32_t* ticket;     // [slot][levels]       last-block tickets (self-resetting)
  unsigned long long* gsum;  // [slot][kMaxLevels] sums of g per level (fused frame kernel; self-resetting)
  int* ithr;            // [slot][kMaxLevels]   integer threshold per level
  uint32_t* cnt;        // [slot][cnt_elems]    counts, then exclusive offsets
  uint32_t* ncand;      // [slot][kMaxLevels]
  uint32_t* sel_mask;   // [slot][mask_elems]   selection bitmask of the levels without records
  uint64_t* rec;        // [slot][rec_elems]    packed records, same order
  uint16_t* dep;        // [slot][plane_elems]  16-bit depth pyramid (depth modes only)
  uint16_t* recz;       // [slot][rec_elems]    integer depth of every record (depth modes only)
};

// 3x3 derivative stencil as four packed signed-byte rows over a (left, centre, right, unused)
// window: d = horizontal difference row (outer rows), dm = its middle row, sp / sm = +/- smoothing
// row of the vertical difference.  Scharr (3, 10, 3) is the reference (Tracker.cpp:1133-1134),
// Sobel (1, 2, 1) the north-star wording.
struct StencilWeights {
  uint32_t d, dm, sp, sm;
};
__host__ __device__ inline StencilWeights stencil_weights(int gradient_op) {
  if (gradient_op == UWT_GRADIENT_SOBEL)
    return {0x000100FFu, 0x000200FEu, 0x00010201u, 0x00FFFEFFu};
  return {0x000300FDu, 0x000A00F6u, 0x00030A03u, 0x00FDF6FDu};
}

// The integer depth ObtainCandidatePoints reads for pixel (x, y) of a level (Tracker.cpp:1339,
// 1344).  REFERENCE mode reproduces depths_[lvl].at<uchar>(y, x) on the CV_16U image: byte x of
// row y, i.e. the low (x even) or high (x odd) byte of depth pixel x / 2.
__device__ __forceinline__ int depth_at(const uint16_t* __restrict__ plane, int pitch, int x, int y,
                                        int depth_mode) {
  if (depth_mode == UWT_DEPTH_REFERENCE) {
    const unsigned v = plane[(size_t)y * pitch + (x >> 1)];
    return (x & 1) ? (int)(v >> 8) : (int)(v & 0xFFu);
  }
  if (depth_mode == UWT_DEPTH_ALL_POINTS) {  // at<short>(y, x) > 0, Tracker.cpp:1273
    const int v = (int)(int16_t)plane[(size_t)y * pitch + x];
    return v > 0 ? v : 0;
  }
  return (int)plane[(size_t)y * pitch + x];
}

// The depth test of ObtainCandidatePoints / ObtainAllPoints (depth_at(..) != 0) for the 4 pixels
// gx .. gx + 3 of one depth row (gx a multiple of 4, gx < pitch) from ONE vector load: 0 / 1 per
// byte.  Pixels beyond the image width are the caller's business (valid4).
__device__ __forceinline__ uint32_t depth_nz4(const uint16_t* __restrict__ row, int gx,
                                              int depth_mode) {
  if (depth_mode == UWT_DEPTH_REFERENCE) {
    // at<uchar>(y, x) on the CV_16U image: byte x of the row
    const uint32_t v = __ldg(
        reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(row) + gx));
    return ((v | ((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) >> 7) & 0x01010101u;
  }
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(row + gx));
  const bool positive = depth_mode == UWT_DEPTH_ALL_POINTS;  // at<short>(y, x) > 0
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t w = k ? v.y : v.x;
    // bit 15 / 31 of (w | ((w & 0x7FFF..) + 0x7FFF..)): the halfword is non-zero
    uint32_t t = (w | ((w & 0x7FFF7FFFu) + 0x7FFF7FFFu)) >> 15;
    if (positive) t &= ~(w >> 15);  // sign bit clear
    t &= 0x00010001u;
    out |= ((t | (t >> 8)) & 0x0101u) << (16 * k);
  }
  return out;
}
// 0x01 in the bytes of the pixels gx .. gx + 3 that lie inside an image of width w
__device__ __forceinline__ uint32_t valid4(int gx, int w) {
  const int n = w - gx;
  return n >= 4 ? 0x01010101u : (n <= 0 ? 0u : (0x01010101u >> (8 * (4 - n))));
}

struct EstimateIO {
  const int* prev_slots;
  const int* cur_slots;
  const float* init_poses;  // nullable, [n][7]
  float* out_poses;         // [n][7]
  uwt_track_stats* stats;   // [n]
  uwt_iter_trace* trace;    // nullable, [n][trace_cap]
  int* trace_count;         // [n]
  int trace_cap;
};

// Device-resident state of the sharded single-frame mode (one problem split over ranks).
struct ShardState {
  float pose[7];
  float last_error;
  int level, k, done;
  int rank, nranks, prev_slot, cur_slot;
  unsigned ticket;
  uwt_track_stats stats;
};
constexpr int kShardMaxGrid = 148 * 2;
constexpr int kShardMaxRanks = 16;

// Mailbox of the fused (in-kernel) all-reduce: peers store their 32 partial sums directly into
// this rank's copy over NVLink (IPC- or peer-mapped memory).  "LL" protocol: every 8-byte word
// carries 32 data bits and the 32-bit sequence number of the sweep, so data and flag arrive in
// ONE atomic 8-byte store and no memory fence is needed on either side; a double takes two words.
struct ShardMailbox {
  unsigned long long ll[2][kShardMaxRanks][64];  // [parity][source rank][2 * value + half]
};

// Control block of the persistent fused kernel (local device memory).
struct ShardFused {
  unsigned long long seq;        // sweeps completed so far on this handle (monotonic)
  unsigned int generation;       // grid barrier: sweeps whose update is published
  unsigned int error;            // != 0: a bounded wait expired
  unsigned long long dbg[8];     // leader-phase cycle counters (profiling aid)
  ShardMailbox* peer[kShardMaxRanks];  // peer[r] = rank r's mailbox as mapped in this process
  // The update's result for the next sweep, broadcast to the waiting CTAs in the same "LL" form as
  // the mailbox: word i = {32-bit payload | generation << 32}, payload = pose[0..6], level, done.
  // A CTA learns "the update is published" and the new state in ONE round trip.
  unsigned long long state_ll[16];
};

// Undistortion front-end fused into the pyramid kernel's level-0 load (System.cpp:232-235):
// fixed-point maps of cv::initUndistortRectifyMap(.., CV_16SC2, ..) in device memory.
struct RemapArgs {
  const short2* map1 = nullptr;    // [map_h][map_w] integer source (x, y); nullptr = no remap
  const uint16_t* map2 = nullptr;  // [map_h][map_w] (fy << 5) | fx
  int map_w = 0, map_h = 0;
  int in_w = 0, in_h = 0;          // size of the distorted source frames
  int roi_x = 0, roi_y = 0;        // top-left corner of the crop inside the maps
};

// A contiguous range of pyramid levels and the matching ranges of the flattened per-level tile
// (K2) and work-item (K3) indices: K2 / K3 normally run on all levels; with
// UWT_FLAG_LAZY_LEVELS only on the levels EstimatePose optimises.
struct LevelRange {
  int lvl_begin = 0, lvl_count = 0;
  int tile_begin = 0, tile_count = 0;
  int item_begin = 0, item_count = 0;
};
inline LevelRange level_range(const Geom& g, int lo, int hi) {
  LevelRange r;
  r.lvl_begin = lo;
  r.lvl_count = hi - lo + 1;
  for (int l = 0; l <= hi; ++l) {
    const int tiles = g.lv[l].tiles_x * g.lv[l].tiles_y, items = g.lv[l].nstrip * g.lv[l].nseg;
    if (l < lo) {
      r.tile_begin += tiles;
      r.item_begin += items;
    } else {
      r.tile_count += tiles;
      r.item_count += items;
    }
  }
  return r;
}

// kernel launchers (each returns the number of kernels launched, or <0 on launch error)
int launch_pyramid(const Geom& g, const Pools& p, int n, const int* d_slots, const uint8_t* src,
                   size_t row_stride, size_t frame_stride, bool src_is_slot, cudaStream_t st,
                   const RemapArgs& rm = RemapArgs());
// K1+K2 fused (pyramid + gradient images of all levels from one tensor-copy-staged read of the
// frame); -2 = this source / configuration needs the separate kernels
bool frame_fused_supported(const Geom& g);
int launch_frame_fused(const Geom& g, const Pools& p, int n, const int* d_slots,
                       const uint8_t* src, size_t row_stride, size_t frame_stride,
                       cudaStream_t st);
// tiled tensor map over u8 planes (x, y, frame) with a box_w x box_h x 1 box, zero fill outside;
// false if the driver has no encoder or the base / strides are not 16-byte aligned
bool encode_u8_map3d(CUtensorMap* out, const uint8_t* base, uint64_t w, uint64_t h, uint64_t n,
                     uint64_t row_stride, uint64_t frame_stride, uint32_t box_w, uint32_t box_h);
int launch_depth_import(const Geom& g, const Pools& p, int n, const int* d_slots,
                        const uint8_t* d_src, size_t row_stride, size_t frame_stride,
                        cudaStream_t st);
int launch_depth_pyramid(const Geom& g, const Pools& p, int n, const int* d_slots,
                         cudaStream_t st);
int launch_remap(const uint8_t* d_src, size_t row_stride, int in_w, int in_h, const short2* map1,
                 const uint16_t* map2, int out_w, int out_h, uint8_t* d_dst, cudaStream_t st);
int launch_gradient(const Geom& g, const Pools& p, int n, const int* d_slots, cudaStream_t st,
                    const LevelRange& lr, int16_t* gx_out = nullptr, int16_t* gy_out = nullptr);
int launch_candidates(const Geom& g, const Pools& p, int n, const int* d_slots, cudaStream_t st,
                      const LevelRange& lr, bool with_mask_levels);
constexpr int UWT_EST_MMA = 0;        // Gram accumulator in fp64 tensor-core fragments
constexpr int UWT_EST_REGISTERS = 1;  // 27 fp64 register accumulators per thread
int launch_estimate(const Geom& g, const Pools& p, int n, const EstimateIO& io, int cluster,
                    cudaStream_t st, int variant);
// persistent dataflow form for batches (chunk tasks from a global ring, no barriers)
size_t flow_workspace_bytes(const Geom& g, int nprob);
#ifdef UWT_FLOW_STATS
void flow_debug_dump();
#endif
// `grid_cache`: per-handle storage of the persistent grid size (0 = not computed yet)
int launch_estimate_flow(const Geom& g, const Pools& p, int n, const EstimateIO& io,
                         void* workspace, cudaStream_t st, int* grid_cache);
int launch_shard_accumulate(const Geom& g, const Pools& p, ShardState* st, double* partials,
                            double* out32, int grid, cudaStream_t stream);
int launch_shard_update(const Geom& g, const Pools& p, ShardState* st, const double* sums32,
                        int* done_out, cudaStream_t stream);
int launch_shard_fused(const Geom& g, const Pools& p, ShardState* st, ShardFused* ctl,
                       ShardMailbox* mine, double* partials, int grid, cudaStream_t stream);
int launch_warp_points(const Geom& g, const float* d_pts4, int n, const float* d_pose7, int level,
                       float* d_out4, cudaStream_t st);

}  // namespace uwt
