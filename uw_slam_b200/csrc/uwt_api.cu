// uwt_api.cu -- host side of the C ABI declared in include/uwtrack.h.
//
// Owns the device memory pools (one "slot" per uw::Frame), the per-level geometry and
// intrinsics (Tracker::InitializePyramid, /root/reference/src/Tracker.cpp:297-340), a ring
// of pinned argument buffers, and enqueues the kernels of the hot path on one stream.
// There is no CPU fallback anywhere in this file: without a usable CUDA device every entry
// point returns UWT_E_CUDA.
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "uwt_internal.cuh"

using namespace uwt;

namespace {

constexpr int kRing = 8;
// The dataflow estimate kernel serves the batch sizes where one-cluster-per-problem suffers from
// wave quantisation and per-sweep barriers (measured on B200, 128 problems at 1280x1024: 0.96 vs
// 1.03 ms).  Fewer problems: the 16-CTA cluster has the lower latency; from 2 CTAs per SM upwards
// single-CTA problems balance by themselves (8192 x 640x480: 330 k vs 310 k tracks/s).
constexpr int kFlowMinProblems = 24;
#ifndef UWT_FLOW_MAX_PROBLEMS
#define UWT_FLOW_MAX_PROBLEMS (2 * 148)
#endif
constexpr int kFlowMaxProblems = UWT_FLOW_MAX_PROBLEMS;
constexpr int kFlowMaxLargeProblems = 4096;  // bounds the workspace (82 KB per 1280x1024 problem)
constexpr int kTraceProblems = 64;

thread_local std::string g_create_error;

struct SlotState {
  bool pyramid = false, gradient = false, candidates = false, depth = false;
  // UWT_FLAG_LAZY_LEVELS: gradient / candidates exist on the optimised levels only until a
  // read-back asks for another level (then all levels are materialised)
  bool gradient_all = false, candidates_all = false;
};

struct ArgRegion {
  int* h_int = nullptr;      // pinned: [2 * max_frames]
  float* h_flt = nullptr;    // pinned: [7 * max_frames]
  int* d_int = nullptr;
  float* d_flt = nullptr;
  cudaEvent_t ev = nullptr;
  bool pending = false;
};

}  // namespace

struct uwt_tracker {
  uwt_config cfg;
  Geom geom;
  Pools pools;
  cudaStream_t stream = nullptr;
  std::vector<SlotState> slots;
  ArgRegion ring[kRing];
  int ring_next = 0;
  // upload path: H2D copies run on their own stream into one of two staging buffers so that
  // the copy of the next batch overlaps the kernels of the current one
  cudaStream_t copy_stream = nullptr;
  uint8_t* d_stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_ready[2] = {nullptr, nullptr};  // recorded on copy_stream after the H2D
  cudaEvent_t stage_free[2] = {nullptr, nullptr};   // recorded on stream after the pyramid
  bool stage_busy[2] = {false, false};
  int stage_next = 0;
  size_t stage_frames = 0;
  cudaEvent_t poses_ready = nullptr;  // recorded after the D2H of the last estimate
  // undistortion front-end (uwt_set_undistortion): device copies of the fixed-point maps
  RemapArgs remap;
  short2* d_map1 = nullptr;
  uint16_t* d_map2 = nullptr;
  // sharded single-frame mode
  ShardState* d_shard = nullptr;
  double* d_shard_partials = nullptr;
  int* d_shard_done = nullptr;
  int* h_shard_done = nullptr;  // pinned
  bool shard_active = false;
  // fused (in-kernel all-reduce) sharded mode
  ShardFused* d_fused = nullptr;
  ShardMailbox* d_mailbox = nullptr;
  ShardFused h_fused;                 // host mirror of the peer table
  void* ipc_opened[kShardMaxRanks] = {};
  int fused_rank = -1, fused_nranks = 0;
  // private single-rank instance of the fused kernel: whole-GPU path of ONE large problem
  ShardFused* d_fused_self = nullptr;
  ShardMailbox* d_mailbox_self = nullptr;
  void* d_flow_ws = nullptr;          // workspace of the batched dataflow estimate kernel
  size_t flow_ws_bytes = 0;
  int* h_flow_ctl = nullptr;          // pinned copy of its control block {head, tail, active, error}
  bool flow_last = false;
  uint8_t* d_depth_stage = nullptr;   // device staging of host depth frames (grow-only)
  size_t depth_stage_frames = 0;
  void* d_scratch = nullptr;          // grow-only scratch of the read-back accessors
  size_t scratch_bytes = 0;
  ShardState* h_shard_in[4] = {};     // pinned staging of uwt_shard_begin (ring of 4)
  int h_shard_next = 0;
  cudaEvent_t h_shard_ev[4] = {};
  ShardState* h_shard_out = nullptr;  // pinned: state read back after a fused estimate
  ShardFused* h_fused_out = nullptr;  // pinned: control block (error flag, phase counters)
  cudaEvent_t fused_done = nullptr;
  int shard_grid = 0;                 // CTAs of the fused single-problem kernel (0 = all that fit)
  int flow_grid = 0;                  // persistent grid of the dataflow kernel (0 = not computed)
  int sm_count = 148;
  bool last_traced = false;           // the last estimate filled d_trace / d_trace_count
  float* d_out_poses = nullptr;
  uwt_track_stats* d_stats = nullptr;
  float* h_out_poses = nullptr;       // pinned
  uwt_track_stats* h_stats = nullptr; // pinned
  uwt_iter_trace* d_trace = nullptr;
  int* d_trace_count = nullptr;
  int trace_cap = 0;
  int last_n = 0;
  int max_cluster = 1;
  long long launches = 0;
  long long aux_launches = 0;  // argument-staging kernels (push_words)
  std::string error;
  // optional per-kernel-class event timing
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span { int cls; cudaEvent_t a, b; int launches; };
  std::vector<Span> spans;
};

namespace {

// Frames per staging buffer.  The two buffers alternate between uploads, so the copy of the next
// batch overlaps the kernels of the current one only while ONE upload fits ONE buffer: a tracker
// that alternates two halves of its slots (previous / current frames) uploads max_frames / 2
// frames per call.  At least 256 MiB, at most 1 GiB per buffer.
size_t stage_capacity(size_t max_frames, size_t frame_bytes) {
  const size_t lo = ((size_t)256 << 20) / frame_bytes, hi = ((size_t)1 << 30) / frame_bytes;
  const size_t want = std::max(lo, std::min((max_frames + 1) / 2, hi));
  return std::max<size_t>(1, std::min(max_frames, want));
}

int fail(uwt_tracker* t, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (t)
    t->error = buf;
  else
    g_create_error = buf;
  return code;
}

#define UWT_CUDA(t, expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return fail((t), UWT_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));   \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Grow-only device scratch for the accessors that are not on the hot path (uwt_warp_points,
// uwt_get_gradients): no cudaMalloc / cudaFree per call.  Calls on a handle are serialised by
// the caller and every user synchronises the stream before it returns.
void* scratch(uwt_tracker* t, size_t bytes) {
  if (bytes > t->scratch_bytes) {
    cudaStreamSynchronize(t->stream);
    cudaFree(t->d_scratch);
  cudaFree(t->d_depth_stage);
    t->d_scratch = nullptr;
    t->scratch_bytes = 0;
    const size_t want = align_up(bytes, (size_t)1 << 20);
    if (cudaMalloc(&t->d_scratch, want) != cudaSuccess) return nullptr;
    t->scratch_bytes = want;
  }
  return t->d_scratch;
}

cudaEvent_t prof_event(uwt_tracker* t) {
  if (t->ev_used == t->ev_pool.size()) {
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    t->ev_pool.push_back(e);
  }
  return t->ev_pool[t->ev_used++];
}

// RAII span: records an event pair around one kernel-class call when profiling is on.
struct ProfSpan {
  uwt_tracker* t;
  int cls;
  cudaEvent_t a = nullptr;
  ProfSpan(uwt_tracker* t_, int cls_) : t(t_), cls(cls_) {
    if (t->profiling) {
      a = prof_event(t);
      cudaEventRecord(a, t->stream);
    }
  }
  void done(int launches) {
    if (a) {
      cudaEvent_t b = prof_event(t);
      cudaEventRecord(b, t->stream);
      t->spans.push_back({cls, a, b, launches});
      a = nullptr;
    }
  }
};

int build_geom(const uwt_config& c, Geom& g) {
  std::memset(&g, 0, sizeof(g));
  g.levels = c.levels;
  g.max_slots = c.max_frames;
  g.first_level = c.first_level;
  g.last_level = c.last_level;
  g.max_iterations = c.max_iterations;
  g.epsilon = c.epsilon;
  g.residual_scale = c.residual_scale;
  g.residual_scale_is_int =
      (c.residual_scale == std::trunc(c.residual_scale) && std::fabs(c.residual_scale) <= 32768.0f)
          ? 1 : 0;
  g.residual_scale_int = g.residual_scale_is_int ? (int)c.residual_scale : 0;
  g.gradient_threshold = c.gradient_threshold;
  g.solve_mode = c.solve_mode;
  g.lm_lambda = c.lm_lambda;
  g.weight_mode = c.weight_mode;
  g.huber_delta = c.huber_delta;
  g.depth_mode = c.depth_mode;
  g.gradient_op = c.gradient_op;
  g.sampling = c.sampling;
  size_t plane = 0, mask = 0, rec = 0, cnt = 0;
  int tiles = 0, items = 0;
  // Tracker::InitializePyramid, Tracker.cpp:297-340 (same expression types as the source:
  // float members, double literals)
  float fx[kMaxLevels], fy[kMaxLevels], cx[kMaxLevels], cy[kMaxLevels];
  fx[0] = c.fx; fy[0] = c.fy; cx[0] = c.cx; cy[0] = c.cy;
  for (int l = 0; l < c.levels; ++l) {
    LevelGeom& L = g.lv[l];
    if (l > 0) {
      fx[l] = fx[l - 1] * 0.5;
      fy[l] = fy[l - 1] * 0.5;
      cx[l] = (cx[0] + 0.5) / ((int)1 << l) - 0.5;
      cy[l] = (cy[0] + 0.5) / ((int)1 << l) - 0.5;
    }
    L.w = c.width >> l;
    L.h = c.height >> l;
    L.pitch = (int)align_up((size_t)L.w, 16);
    L.wf = (float)L.w; L.hf = (float)L.h;
    L.wm1 = L.w - 1; L.hm1 = L.h - 1;
    L.fx = fx[l]; L.fy = fy[l]; L.cx = cx[l]; L.cy = cy[l];
    L.invfx = 1 / fx[l];
    L.invfy = 1 / fy[l];
    L.plane_off = (int)plane;
    plane += align_up((size_t)L.pitch * L.h, 256);
    L.mask_wpr = (L.w + 31) / 32;
    if (l >= c.last_level && l <= c.first_level) {
      L.rec_off = (int)rec;
      rec += align_up((size_t)L.w * L.h, 64);
      L.mask_off = -1;
    } else {
      // a level EstimatePose never optimises keeps its selection as a bitmask
      L.rec_off = -1;
      L.mask_off = (int)mask;
      mask += align_up((size_t)L.mask_wpr * L.h, 64);
    }
    L.nstrip = (L.w + kStripW - 1) / kStripW;
    L.nseg = (L.h + kSegRows - 1) / kSegRows;
    L.cnt_off = (int)cnt;
    cnt += align_up((size_t)L.w * L.nseg, 64);
    L.tiles_x = (L.w + kGradTileW - 1) / kGradTileW;
    L.tiles_y = (L.h + kGradTileH - 1) / kGradTileH;
    L.tile_off = tiles;
    tiles += L.tiles_x * L.tiles_y;
    items += L.nstrip * L.nseg;
  }
  for (int l = c.last_level; l <= c.first_level; ++l)
    if (std::fabs(g.lv[l].cx) < 0.00390625f || std::fabs(g.lv[l].cy) < 0.00390625f)
      g.exact_div = 1;
  g.plane_elems = plane;
  g.mask_elems = mask ? mask : 64;
  g.mask_words_total = (int)mask;
  g.rec_elems = rec ? rec : 64;
  g.cnt_elems = cnt;
  g.tile_elems = align_up((size_t)tiles, 64);
  g.grad_tiles_total = tiles;
  g.warp_items_total = items;
  return 0;
}

int check_slots(uwt_tracker* t, int n, const int* slots) {
  if (!t) return UWT_E_INVALID;
  if (n <= 0 || n > t->cfg.max_frames || !slots)
    return fail(t, UWT_E_INVALID, "n=%d out of range (1..%d) or slots NULL", n, t->cfg.max_frames);
  for (int i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= t->cfg.max_frames)
      return fail(t, UWT_E_INVALID, "slot %d out of range (0..%d)", slots[i], t->cfg.max_frames - 1);
  return UWT_OK;
}

// Acquire the next argument region (waits only if the GPU is kRing calls behind).
int acquire(uwt_tracker* t, ArgRegion** out) {
  ArgRegion& r = t->ring[t->ring_next];
  t->ring_next = (t->ring_next + 1) % kRing;
  if (r.pending) {
    UWT_CUDA(t, cudaEventSynchronize(r.ev));
    r.pending = false;
  }
  *out = &r;
  return UWT_OK;
}

int release(uwt_tracker* t, ArgRegion* r) {
  UWT_CUDA(t, cudaEventRecord(r->ev, t->stream));
  r->pending = true;
  return UWT_OK;
}

// Argument arrays reach the device through a KERNEL that reads the pinned host buffer, not
// through a host-to-device copy: a small copy ordered behind a kernel on the compute stream sits
// in the copy engine's queue until that kernel ends and holds up the next batch's frame upload
// queued after it (measured at 256 sequences per step: 7.6 instead of 6.1 ms per step).
__global__ void arg_copy_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src,
                                int words) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x)
    dst[i] = src[i];
}
int push_words(uwt_tracker* t, void* d_dst, const void* h_pinned, size_t words) {
  void* mapped = nullptr;
  UWT_CUDA(t, cudaHostGetDevicePointer(&mapped, const_cast<void*>(h_pinned), 0));
  const int grid = (int)std::min<size_t>(32, (words + 255) / 256);
  arg_copy_kernel<<<std::max(grid, 1), 256, 0, t->stream>>>(
      static_cast<uint32_t*>(d_dst), static_cast<const uint32_t*>(mapped), (int)words);
  UWT_CUDA(t, cudaGetLastError());
  t->aux_launches += 1;
  return UWT_OK;
}

int push_slots(uwt_tracker* t, ArgRegion* r, int n, const int* a, const int* b) {
  std::memcpy(r->h_int, a, sizeof(int) * n);
  if (b) std::memcpy(r->h_int + n, b, sizeof(int) * n);
  return push_words(t, r->d_int, r->h_int, (size_t)n * (b ? 2 : 1));
}

void destroy_impl(uwt_tracker* t) {
  if (!t) return;
  cudaSetDevice(t->cfg.device);
  if (t->copy_stream) cudaStreamSynchronize(t->copy_stream);
  if (t->stream) cudaStreamSynchronize(t->stream);
  Pools& p = t->pools;
  cudaFree(p.img); cudaFree(p.g); cudaFree(p.gpart); cudaFree(p.gsum);
  cudaFree(p.ticket); cudaFree(p.ithr); cudaFree(p.cnt); cudaFree(p.ncand);
  cudaFree(p.sel_mask); cudaFree(p.rec); cudaFree(p.dep); cudaFree(p.recz);
  for (cudaEvent_t e : t->ev_pool) cudaEventDestroy(e);
  for (ArgRegion& r : t->ring) {
    if (r.h_int) cudaFreeHost(r.h_int);
    if (r.h_flt) cudaFreeHost(r.h_flt);
    cudaFree(r.d_int);
    cudaFree(r.d_flt);
    if (r.ev) cudaEventDestroy(r.ev);
  }
  for (int i = 0; i < 2; ++i) {
    cudaFree(t->d_stage[i]);
    if (t->stage_ready[i]) cudaEventDestroy(t->stage_ready[i]);
    if (t->stage_free[i]) cudaEventDestroy(t->stage_free[i]);
  }
  if (t->poses_ready) cudaEventDestroy(t->poses_ready);
  cudaFree(t->d_map1);
  cudaFree(t->d_map2);
  for (void* p : t->ipc_opened)
    if (p) cudaIpcCloseMemHandle(p);
  cudaFree(t->d_fused);
  cudaFree(t->d_mailbox);
  cudaFree(t->d_fused_self);
  cudaFree(t->d_mailbox_self);
  cudaFree(t->d_shard);
  cudaFree(t->d_shard_partials);
  cudaFree(t->d_shard_done);
  if (t->h_shard_done) cudaFreeHost(t->h_shard_done);
  for (int i = 0; i < 4; ++i) {
    if (t->h_shard_in[i]) cudaFreeHost(t->h_shard_in[i]);
    if (t->h_shard_ev[i]) cudaEventDestroy(t->h_shard_ev[i]);
  }
  if (t->h_shard_out) cudaFreeHost(t->h_shard_out);
  if (t->h_fused_out) cudaFreeHost(t->h_fused_out);
  if (t->fused_done) cudaEventDestroy(t->fused_done);
  if (t->copy_stream) cudaStreamDestroy(t->copy_stream);
  cudaFree(t->d_flow_ws);
  cudaFree(t->d_scratch);
  cudaFree(t->d_depth_stage);
  if (t->h_flow_ctl) cudaFreeHost(t->h_flow_ctl);
  cudaFree(t->d_out_poses);
  cudaFree(t->d_stats);
  cudaFree(t->d_trace);
  cudaFree(t->d_trace_count);
  if (t->h_out_poses) cudaFreeHost(t->h_out_poses);
  if (t->h_stats) cudaFreeHost(t->h_stats);
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
}

}  // namespace

extern "C" {

int uwt_default_config(uwt_config* cfg) {
  if (!cfg) return UWT_E_INVALID;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->width = 640;
  cfg->height = 480;
  cfg->fx = 525.0f;  // calibration/calibrationTUM.xml:20
  cfg->fy = 525.0f;
  cfg->cx = 319.5f;
  cfg->cy = 239.5f;
  cfg->levels = 5;            // Options.cpp:26
  cfg->first_level = 4;       // Tracker.cpp:368
  cfg->last_level = 1;        // Tracker.cpp:369
  cfg->max_iterations = 50;   // Tracker.cpp:366
  cfg->epsilon = 0.001f;      // Tracker.cpp:364
  cfg->residual_scale = 50.0f;      // Tracker.cpp:559
  cfg->gradient_threshold = 20.0;   // Options.cpp:27
  cfg->solve_mode = UWT_SOLVE_LU;
  cfg->device = 0;
  cfg->max_frames = 2;
  cfg->cluster_size = 0;
  cfg->flags = 0;
  cfg->weight_mode = UWT_WEIGHT_IDENTITY;  // Tracker.cpp:495
  cfg->huber_delta = 10.0f;
  cfg->depth_mode = UWT_DEPTH_NONE;  // Tracker(depth_available = false), System.cpp:121
  cfg->lm_lambda = 0.2f;             // "float LM_lambda = 0.2" in the comment at Tracker.cpp:546
  cfg->gradient_op = UWT_GRADIENT_SCHARR;  // Tracker.cpp:1133
  cfg->sampling = UWT_SAMPLE_NEAREST;      // Tracker.cpp:472
  return UWT_OK;
}

int uwt_create(const uwt_config* cfg, uwt_tracker** out) {
  if (out) *out = nullptr;
  if (!cfg || !out) return fail(nullptr, UWT_E_INVALID, "cfg/out is NULL");
  const uwt_config& c = *cfg;
  if (c.levels < 1 || c.levels > kMaxLevels)
    return fail(nullptr, UWT_E_INVALID, "levels=%d unsupported (1..%d)", c.levels, kMaxLevels);
  const int div = 1 << (c.levels - 1);
  // The reference sizes levels as w>>l while cv::resize rounds: they only agree when the
  // size is divisible by 2^(levels-1); 16-byte rows keep loads vectorised.
  if (c.width <= 0 || c.height <= 0 || c.width % div || c.height % div || c.width % 16 ||
      c.height % 2)
    return fail(nullptr, UWT_E_INVALID,
                "frame %dx%d must be divisible by 2^(levels-1)=%d and the width by 16", c.width,
                c.height, div);
  if ((c.width >> (c.levels - 1)) < 2 || (c.height >> (c.levels - 1)) < 2)
    return fail(nullptr, UWT_E_INVALID, "coarsest level smaller than 2x2");
  if (c.width > 4096 || c.height > 4096)
    return fail(nullptr, UWT_E_INVALID, "frame larger than 4096 not supported (12-bit records)");
  if (c.first_level >= c.levels || c.last_level < 0 || c.last_level > c.first_level)
    return fail(nullptr, UWT_E_INVALID, "bad level range first=%d last=%d", c.first_level,
                c.last_level);
  if (!(c.gradient_threshold >= 0.0))
    return fail(nullptr, UWT_E_INVALID, "gradient_threshold must be >= 0 (the reference uses 20)");
  if (c.max_iterations < 1 || c.max_frames < 1)
    return fail(nullptr, UWT_E_INVALID, "max_iterations and max_frames must be >= 1");
  if (c.solve_mode == UWT_SOLVE_CHOLESKY_LM && !(c.lm_lambda >= 0.0f))
    return fail(nullptr, UWT_E_INVALID, "lm_lambda must be >= 0");
  if (c.solve_mode != UWT_SOLVE_LU && c.solve_mode != UWT_SOLVE_INVERSE &&
      c.solve_mode != UWT_SOLVE_CHOLESKY_LM)
    return fail(nullptr, UWT_E_INVALID, "bad solve_mode %d", c.solve_mode);
  if (c.weight_mode < UWT_WEIGHT_IDENTITY || c.weight_mode > UWT_WEIGHT_HUBER)
    return fail(nullptr, UWT_E_INVALID, "bad weight_mode %d", c.weight_mode);
  if (c.weight_mode == UWT_WEIGHT_HUBER && !(c.huber_delta > 0.0f))
    return fail(nullptr, UWT_E_INVALID, "huber_delta must be > 0");
  if (c.sampling != UWT_SAMPLE_NEAREST && c.sampling != UWT_SAMPLE_BILINEAR)
    return fail(nullptr, UWT_E_INVALID, "bad sampling %d", c.sampling);
  if (c.sampling == UWT_SAMPLE_BILINEAR &&
      (c.weight_mode != UWT_WEIGHT_IDENTITY || c.depth_mode != UWT_DEPTH_NONE ||
       (c.flags & UWT_FLAG_DMMA_ACCUM)))
    return fail(nullptr, UWT_E_INVALID,
                "bilinear sampling supports identity weights, mono input, register accumulator");
  if (c.gradient_op != UWT_GRADIENT_SCHARR && c.gradient_op != UWT_GRADIENT_SOBEL)
    return fail(nullptr, UWT_E_INVALID, "bad gradient_op %d", c.gradient_op);
  if (c.depth_mode < UWT_DEPTH_NONE || c.depth_mode > UWT_DEPTH_ALL_POINTS)
    return fail(nullptr, UWT_E_INVALID, "bad depth_mode %d", c.depth_mode);
  if (c.depth_mode != UWT_DEPTH_NONE &&
      (c.weight_mode != UWT_WEIGHT_IDENTITY || (c.flags & UWT_FLAG_DMMA_ACCUM)))
    return fail(nullptr, UWT_E_INVALID,
                "depth input supports identity weights and the register accumulator only");
  if (c.weight_mode != UWT_WEIGHT_IDENTITY && (c.flags & UWT_FLAG_DMMA_ACCUM))
    return fail(nullptr, UWT_E_INVALID, "UWT_FLAG_DMMA_ACCUM supports identity weights only");
  if (c.cluster_size != 0 && c.cluster_size != 1 && c.cluster_size != 2 && c.cluster_size != 4 &&
      c.cluster_size != 8 && c.cluster_size != 16)
    return fail(nullptr, UWT_E_INVALID, "cluster_size must be 0, 1, 2, 4, 8 or 16");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, UWT_E_CUDA, "no CUDA device available (%s); there is no CPU fallback",
                cudaGetErrorString(e));
  if (c.device < 0 || c.device >= ndev)
    return fail(nullptr, UWT_E_INVALID, "device %d out of range (0..%d)", c.device, ndev - 1);

  uwt_tracker* t = new (std::nothrow) uwt_tracker();
  if (!t) return fail(nullptr, UWT_E_NOMEM, "out of host memory");
  t->cfg = c;
  build_geom(c, t->geom);
  t->slots.resize(c.max_frames);
#define CREATE_CUDA(expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      fail(nullptr, UWT_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));      \
      destroy_impl(t);                                                                 \
      return UWT_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)
  CREATE_CUDA(cudaSetDevice(c.device));
  CREATE_CUDA(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
  CREATE_CUDA(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
  const Geom& g = t->geom;
  const size_t F = (size_t)c.max_frames;
  Pools& p = t->pools;
  CREATE_CUDA(cudaMalloc(&p.img, F * g.plane_elems));
  CREATE_CUDA(cudaMalloc(&p.g, F * g.plane_elems));
  CREATE_CUDA(cudaMalloc(&p.gpart, F * g.tile_elems * sizeof(uint32_t)));
  CREATE_CUDA(cudaMalloc(&p.ticket, F * kMaxLevels * sizeof(uint32_t)));
  CREATE_CUDA(cudaMalloc(&p.gsum, F * kMaxLevels * sizeof(unsigned long long)));
  CREATE_CUDA(cudaMemsetAsync(p.gsum, 0, F * kMaxLevels * sizeof(unsigned long long), t->stream));
  CREATE_CUDA(cudaMalloc(&p.ithr, F * kMaxLevels * sizeof(int)));
  CREATE_CUDA(cudaMalloc(&p.cnt, F * g.cnt_elems * sizeof(uint32_t)));
  CREATE_CUDA(cudaMalloc(&p.ncand, F * kMaxLevels * sizeof(uint32_t)));
  CREATE_CUDA(cudaMalloc(&p.sel_mask, F * g.mask_elems * sizeof(uint32_t)));
  CREATE_CUDA(cudaMalloc(&p.rec, F * g.rec_elems * sizeof(uint64_t)));
  if (c.depth_mode != UWT_DEPTH_NONE) {
    CREATE_CUDA(cudaMalloc(&p.dep, F * g.plane_elems * sizeof(uint16_t)));
    CREATE_CUDA(cudaMalloc(&p.recz, F * g.rec_elems * sizeof(uint16_t)));
    CREATE_CUDA(cudaMemsetAsync(p.dep, 0, F * g.plane_elems * sizeof(uint16_t), t->stream));
  }
  CREATE_CUDA(cudaMemsetAsync(p.ticket, 0, F * kMaxLevels * sizeof(uint32_t), t->stream));
  CREATE_CUDA(cudaMemsetAsync(p.ncand, 0, F * kMaxLevels * sizeof(uint32_t), t->stream));
  // pitch padding bytes are read (never used) by vector loads: keep them defined
  CREATE_CUDA(cudaMemsetAsync(p.img, 0, F * g.plane_elems, t->stream));
  CREATE_CUDA(cudaMemsetAsync(p.g, 0, F * g.plane_elems, t->stream));
  for (ArgRegion& r : t->ring) {
    CREATE_CUDA(cudaHostAlloc(&r.h_int, sizeof(int) * 2 * F, cudaHostAllocDefault));
    CREATE_CUDA(cudaHostAlloc(&r.h_flt, sizeof(float) * 7 * F, cudaHostAllocDefault));
    CREATE_CUDA(cudaMalloc(&r.d_int, sizeof(int) * 2 * F));
    CREATE_CUDA(cudaMalloc(&r.d_flt, sizeof(float) * 7 * F));
    CREATE_CUDA(cudaEventCreateWithFlags(&r.ev, cudaEventDisableTiming));
  }
  const size_t frame_bytes = (size_t)c.width * c.height;
  t->stage_frames = stage_capacity(F, frame_bytes);
  for (int i = 0; i < 2; ++i) {
    CREATE_CUDA(cudaMalloc(&t->d_stage[i], t->stage_frames * frame_bytes));
    CREATE_CUDA(cudaEventCreateWithFlags(&t->stage_ready[i], cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&t->stage_free[i], cudaEventDisableTiming));
  }
  CREATE_CUDA(cudaEventCreateWithFlags(&t->poses_ready, cudaEventDisableTiming));
  CREATE_CUDA(cudaMalloc(&t->d_shard, sizeof(ShardState)));
  CREATE_CUDA(cudaMalloc(&t->d_fused, sizeof(ShardFused)));
  CREATE_CUDA(cudaMalloc(&t->d_mailbox, sizeof(ShardMailbox)));
  CREATE_CUDA(cudaMemsetAsync(t->d_fused, 0, sizeof(ShardFused), t->stream));
  CREATE_CUDA(cudaMemsetAsync(t->d_mailbox, 0, sizeof(ShardMailbox), t->stream));
  std::memset(&t->h_fused, 0, sizeof(t->h_fused));
  CREATE_CUDA(cudaMalloc(&t->d_fused_self, sizeof(ShardFused)));
  CREATE_CUDA(cudaMalloc(&t->d_mailbox_self, sizeof(ShardMailbox)));
  CREATE_CUDA(cudaMemsetAsync(t->d_mailbox_self, 0, sizeof(ShardMailbox), t->stream));
  {
    ShardFused self;
    std::memset(&self, 0, sizeof(self));
    self.peer[0] = t->d_mailbox_self;
    CREATE_CUDA(cudaMemcpy(t->d_fused_self, &self, sizeof(self), cudaMemcpyHostToDevice));
  }
  CREATE_CUDA(cudaMalloc(&t->d_shard_partials, sizeof(double) * 32 * kShardMaxGrid));
  for (int i = 0; i < 4; ++i) {
    CREATE_CUDA(cudaHostAlloc(&t->h_shard_in[i], sizeof(ShardState), cudaHostAllocDefault));
    CREATE_CUDA(cudaEventCreateWithFlags(&t->h_shard_ev[i], cudaEventDisableTiming));
  }
  CREATE_CUDA(cudaHostAlloc(&t->h_shard_out, sizeof(ShardState), cudaHostAllocDefault));
  CREATE_CUDA(cudaHostAlloc(&t->h_fused_out, sizeof(ShardFused), cudaHostAllocDefault));
  CREATE_CUDA(cudaEventCreateWithFlags(&t->fused_done, cudaEventDisableTiming));
  CREATE_CUDA(cudaMalloc(&t->d_shard_done, sizeof(int)));
  CREATE_CUDA(cudaHostAlloc(&t->h_shard_done, sizeof(int), cudaHostAllocDefault));
  CREATE_CUDA(cudaMalloc(&t->d_out_poses, sizeof(float) * 7 * F));
  CREATE_CUDA(cudaMalloc(&t->d_stats, sizeof(uwt_track_stats) * F));
  CREATE_CUDA(cudaHostAlloc(&t->h_out_poses, sizeof(float) * 7 * F, cudaHostAllocDefault));
  CREATE_CUDA(cudaHostAlloc(&t->h_stats, sizeof(uwt_track_stats) * F, cudaHostAllocDefault));
  CREATE_CUDA(cudaHostAlloc(&t->h_flow_ctl, sizeof(int) * 4, cudaHostAllocDefault));
  if (c.flags & UWT_FLAG_TRACE) {
    t->trace_cap = (c.first_level - c.last_level + 1) * c.max_iterations;
    CREATE_CUDA(cudaMalloc(&t->d_trace, sizeof(uwt_iter_trace) * t->trace_cap * kTraceProblems));
    CREATE_CUDA(cudaMalloc(&t->d_trace_count, sizeof(int) * kTraceProblems));
  }
  // largest cluster the device can co-schedule for the estimate kernel
  t->max_cluster = 8;
  {
    cudaDeviceProp prop;
    CREATE_CUDA(cudaGetDeviceProperties(&prop, c.device));
    if (prop.major < 9) {
      fail(nullptr, UWT_E_CUDA, "device sm_%d%d has no thread-block clusters; built for sm_100a",
           prop.major, prop.minor);
      destroy_impl(t);
      return UWT_E_CUDA;
    }
    if (prop.major >= 10) t->max_cluster = 16;
    if (prop.multiProcessorCount > 0) t->sm_count = prop.multiProcessorCount;
  }
  // one persistent CTA per SM: measured better than two (the leader's reduction over the per-CTA
  // partials grows with the grid: 0.186 vs 0.197 ms per 3840x2160 estimate)
  t->shard_grid = std::min(t->sm_count, kShardMaxGrid);
  if (const char* e = getenv("UWT_SHARD_GRID")) t->shard_grid = atoi(e);
  CREATE_CUDA(cudaStreamSynchronize(t->stream));
#undef CREATE_CUDA
  *out = t;
  return UWT_OK;
}

int uwt_destroy(uwt_tracker* t) {
  if (!t) return UWT_E_INVALID;
  destroy_impl(t);
  return UWT_OK;
}

const char* uwt_last_error(const uwt_tracker* t) {
  return t ? t->error.c_str() : g_create_error.c_str();
}

int uwt_get_level_info(const uwt_tracker* t, int level, uwt_level_info* info) {
  if (!t || !info || level < 0 || level >= t->geom.levels) return UWT_E_INVALID;
  const LevelGeom& L = t->geom.lv[level];
  info->width = L.w; info->height = L.h;
  info->fx = L.fx; info->fy = L.fy; info->cx = L.cx; info->cy = L.cy;
  info->invfx = L.invfx; info->invfy = L.invfy;
  return UWT_OK;
}

void* uwt_stream(const uwt_tracker* t) { return t ? (void*)t->stream : nullptr; }

int uwt_synchronize(uwt_tracker* t) {
  if (!t) return UWT_E_INVALID;
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  UWT_CUDA(t, cudaStreamSynchronize(t->copy_stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  return UWT_OK;
}

long long uwt_launch_count(const uwt_tracker* t) { return t ? t->launches : 0; }
long long uwt_aux_launch_count(const uwt_tracker* t) { return t ? t->aux_launches : 0; }

int uwt_profile_enable(uwt_tracker* t, int on) {
  if (!t) return UWT_E_INVALID;
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  t->profiling = on != 0;
  t->spans.clear();
  t->ev_used = 0;
  return UWT_OK;
}

int uwt_profile_read(uwt_tracker* t, double ms[UWT_K_COUNT], long long launches[UWT_K_COUNT]) {
  if (!t || !ms || !launches) return UWT_E_INVALID;
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  for (int i = 0; i < UWT_K_COUNT; ++i) {
    ms[i] = 0.0;
    launches[i] = 0;
  }
  for (const uwt_tracker::Span& s : t->spans) {
    float v = 0.f;
    UWT_CUDA(t, cudaEventElapsedTime(&v, s.a, s.b));
    ms[s.cls] += v;
    launches[s.cls] += s.launches;
  }
  return UWT_OK;
}

static int pyramid_common(uwt_tracker* t, int n, const int* slots, const uint8_t* dev_src,
                          size_t row_stride, size_t frame_stride) {
  ArgRegion* r = nullptr;
  int rc = acquire(t, &r);
  if (rc) return rc;
  rc = push_slots(t, r, n, slots, nullptr);
  if (rc) return rc;
  ProfSpan span(t, UWT_K_PYRAMID);
  // Default: pyramid AND gradient images of all levels in one kernel (one read of the frame,
  // staged by a 2-D tensor copy); uwt_apply_gradient then has nothing left to do for these
  // slots.  The separate kernels serve what the fused one does not: the undistortion front-end,
  // more than 5 levels, UWT_FLAG_LAZY_LEVELS / UWT_FLAG_SEPARATE_GRADIENT, unaligned sources.
  int k = -2;
  bool fused = false;
  if (!t->remap.map1 && !(t->cfg.flags & (UWT_FLAG_LAZY_LEVELS | UWT_FLAG_SEPARATE_GRADIENT))) {
    k = launch_frame_fused(t->geom, t->pools, n, r->d_int, dev_src, row_stride, frame_stride,
                           t->stream);
    fused = k > 0;
  }
  if (k == -2)
    k = launch_pyramid(t->geom, t->pools, n, r->d_int, dev_src, row_stride, frame_stride, false,
                       t->stream, t->remap);
  span.done(k);
  if (k < 0) return fail(t, UWT_E_CUDA, "pyramid kernel launch failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
  t->launches += k;
  rc = release(t, r);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    SlotState& s = t->slots[slots[i]];
    s.pyramid = true;
    s.gradient = s.gradient_all = fused;
    s.candidates = s.candidates_all = s.depth = false;
  }
  return UWT_OK;
}

int uwt_upload_frames(uwt_tracker* t, int n, const int* slots, const uint8_t* host,
                      size_t row_stride, size_t frame_stride) {
  int rc = check_slots(t, n, slots);
  if (rc) return rc;
  // with undistortion enabled the caller hands over DISTORTED in_w x in_h frames
  const size_t w = t->remap.map1 ? t->remap.in_w : t->cfg.width;
  const size_t h = t->remap.map1 ? t->remap.in_h : t->cfg.height;
  if (!host || row_stride < w)
    return fail(t, UWT_E_INVALID, "host NULL or row_stride < source width %zu", w);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  for (size_t done = 0; done < (size_t)n;) {
    const size_t cnt = std::min(t->stage_frames, (size_t)n - done);
    const uint8_t* src = host + done * frame_stride;
    const int sb = t->stage_next;
    t->stage_next ^= 1;
    // the copy engine may refill this staging buffer once the pyramid kernel that last read
    // it has finished
    if (t->stage_busy[sb]) UWT_CUDA(t, cudaStreamWaitEvent(t->copy_stream, t->stage_free[sb], 0));
    uint8_t* stage = t->d_stage[sb];
    if (row_stride == w && frame_stride == w * h) {
      // densely packed frames: one linear copy (largest DMA descriptors)
      UWT_CUDA(t, cudaMemcpyAsync(stage, src, w * h * cnt, cudaMemcpyHostToDevice,
                                  t->copy_stream));
    } else if (frame_stride == row_stride * h) {
      // frames are contiguous: one strided copy for the whole chunk
      UWT_CUDA(t, cudaMemcpy2DAsync(stage, w, src, row_stride, w, h * cnt,
                                    cudaMemcpyHostToDevice, t->copy_stream));
    } else {
      for (size_t i = 0; i < cnt; ++i)
        UWT_CUDA(t, cudaMemcpy2DAsync(stage + i * w * h, w, src + i * frame_stride, row_stride,
                                      w, h, cudaMemcpyHostToDevice, t->copy_stream));
    }
    UWT_CUDA(t, cudaEventRecord(t->stage_ready[sb], t->copy_stream));
    UWT_CUDA(t, cudaStreamWaitEvent(t->stream, t->stage_ready[sb], 0));
    rc = pyramid_common(t, (int)cnt, slots + done, stage, w, w * h);
    if (rc) return rc;
    UWT_CUDA(t, cudaEventRecord(t->stage_free[sb], t->stream));
    t->stage_busy[sb] = true;
    done += cnt;
  }
  return UWT_OK;
}

int uwt_set_frames_device(uwt_tracker* t, int n, const int* slots, const uint8_t* dev,
                          size_t row_stride, size_t frame_stride) {
  int rc = check_slots(t, n, slots);
  if (rc) return rc;
  if (!dev || row_stride < (size_t)(t->remap.map1 ? t->remap.in_w : t->cfg.width))
    return fail(t, UWT_E_INVALID, "dev NULL or row_stride < source width");
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  return pyramid_common(t, n, slots, dev, row_stride, frame_stride);
}

int uwt_upload_depth_frames(uwt_tracker* t, int n, const int* slots, const uint16_t* host,
                            size_t row_stride, size_t frame_stride) {
  int rc = check_slots(t, n, slots);
  if (rc) return rc;
  if (t->cfg.depth_mode == UWT_DEPTH_NONE)
    return fail(t, UWT_E_STATE, "the tracker was created without depth input (cfg.depth_mode)");
  const size_t w = t->cfg.width, h = t->cfg.height;
  if (!host || row_stride < w * sizeof(uint16_t))
    return fail(t, UWT_E_INVALID, "host NULL or row_stride < 2 * width bytes");
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  // Frames already in device memory are read in place; host frames go through a device staging
  // buffer (one strided copy per chunk when the frames are contiguous).  Either way level 0 of
  // all n slots is written by ONE kernel, then the pyramid kernels run on the slots.
  cudaPointerAttributes attr;
  const bool on_device = cudaPointerGetAttributes(&attr, host) == cudaSuccess &&
                         (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
  cudaGetLastError();
  const uint8_t* src = reinterpret_cast<const uint8_t*>(host);
  const size_t dense_row = w * sizeof(uint16_t), dense_frame = dense_row * h;
  const size_t cap = on_device ? (size_t)n
                               : std::max<size_t>(1, std::min<size_t>((size_t)n, ((size_t)128 << 20) / dense_frame));
  if (!on_device && t->depth_stage_frames < cap) {
    UWT_CUDA(t, cudaStreamSynchronize(t->stream));
    cudaFree(t->d_depth_stage);
    t->d_depth_stage = nullptr;
    t->depth_stage_frames = 0;
    UWT_CUDA(t, cudaMalloc(&t->d_depth_stage, cap * dense_frame));
    t->depth_stage_frames = cap;
  }
  for (size_t done = 0; done < (size_t)n; done += cap) {
    const size_t cnt = std::min(cap, (size_t)n - done);
    const uint8_t* csrc = src + done * frame_stride;
    size_t rs = row_stride, fs = frame_stride;
    if (!on_device) {
      if (frame_stride == row_stride * h) {
        UWT_CUDA(t, cudaMemcpy2DAsync(t->d_depth_stage, dense_row, csrc, row_stride, dense_row,
                                      h * cnt, cudaMemcpyHostToDevice, t->stream));
      } else {
        for (size_t i = 0; i < cnt; ++i)
          UWT_CUDA(t, cudaMemcpy2DAsync(t->d_depth_stage + i * dense_frame, dense_row,
                                        csrc + i * frame_stride, row_stride, dense_row, h,
                                        cudaMemcpyHostToDevice, t->stream));
      }
      csrc = t->d_depth_stage;
      rs = dense_row;
      fs = frame_stride == 0 ? 0 : dense_frame;
      if (frame_stride == 0 && cnt > 1) fs = 0;  // one host frame for every slot
    }
    ArgRegion* r = nullptr;
    if ((rc = acquire(t, &r))) return rc;
    if ((rc = push_slots(t, r, (int)cnt, slots + done, nullptr))) return rc;
    ProfSpan span(t, UWT_K_PYRAMID);
    int k = launch_depth_import(t->geom, t->pools, (int)cnt, r->d_int, csrc, rs, fs, t->stream);
    const int k2 = k < 0 ? -1 : launch_depth_pyramid(t->geom, t->pools, (int)cnt, r->d_int, t->stream);
    k = (k < 0 || k2 < 0) ? -1 : k + k2;
    span.done(k);
    if (k < 0) return fail(t, UWT_E_CUDA, "depth pyramid kernel launch failed: %s",
                           cudaGetErrorString(cudaGetLastError()));
    t->launches += k;
    if ((rc = release(t, r))) return rc;
  }
  for (int i = 0; i < n; ++i) {
    SlotState& s = t->slots[slots[i]];
    s.depth = true;
    s.candidates = s.candidates_all = false;  // the candidate rule depends on the depth
  }
  return UWT_OK;
}

int uwt_get_depth(uwt_tracker* t, int slot, int level, uint16_t* host) {
  if (!t) return UWT_E_INVALID;
  if (slot < 0 || slot >= t->cfg.max_frames || level < 0 || level >= t->geom.levels || !host)
    return fail(t, UWT_E_INVALID, "slot %d / level %d out of range or host NULL", slot, level);
  if (!t->slots[slot].depth) return fail(t, UWT_E_STATE, "slot %d has no depth frame", slot);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  const LevelGeom& L = t->geom.lv[level];
  UWT_CUDA(t, cudaMemcpy2DAsync(host, (size_t)L.w * 2,
                                t->pools.dep + (size_t)slot * t->geom.plane_elems + L.plane_off,
                                (size_t)L.pitch * 2, (size_t)L.w * 2, L.h, cudaMemcpyDeviceToHost,
                                t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  return UWT_OK;
}

int uwt_set_undistortion(uwt_tracker* t, const int16_t* map1, const uint16_t* map2, int map_w,
                         int map_h, int in_w, int in_h, int roi_x, int roi_y) {
  if (!t) return UWT_E_INVALID;
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  // frames in flight still use the current maps / staging buffers
  UWT_CUDA(t, cudaStreamSynchronize(t->copy_stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  const bool enable = map1 && map2;
  size_t frame_bytes = (size_t)t->cfg.width * t->cfg.height;
  if (enable) {
    if (map_w < 1 || map_h < 1 || in_w < 1 || in_h < 1 || roi_x < 0 || roi_y < 0 ||
        roi_x + t->cfg.width > map_w || roi_y + t->cfg.height > map_h)
      return fail(t, UWT_E_INVALID,
                  "ROI %dx%d at (%d,%d) does not fit the %dx%d undistortion maps", t->cfg.width,
                  t->cfg.height, roi_x, roi_y, map_w, map_h);
    frame_bytes = (size_t)in_w * in_h;
  } else if (map1 || map2) {
    return fail(t, UWT_E_INVALID, "map1 and map2 must both be given (or both NULL)");
  }
  cudaFree(t->d_map1);
  cudaFree(t->d_map2);
  t->d_map1 = nullptr;
  t->d_map2 = nullptr;
  t->remap = RemapArgs();
  if (enable) {
    const size_t n = (size_t)map_w * map_h;
    UWT_CUDA(t, cudaMalloc(&t->d_map1, n * sizeof(short2)));
    UWT_CUDA(t, cudaMalloc(&t->d_map2, n * sizeof(uint16_t)));
    UWT_CUDA(t, cudaMemcpy(t->d_map1, map1, n * sizeof(short2), cudaMemcpyHostToDevice));
    UWT_CUDA(t, cudaMemcpy(t->d_map2, map2, n * sizeof(uint16_t), cudaMemcpyHostToDevice));
    t->remap.map1 = t->d_map1;
    t->remap.map2 = t->d_map2;
    t->remap.map_w = map_w; t->remap.map_h = map_h;
    t->remap.in_w = in_w; t->remap.in_h = in_h;
    t->remap.roi_x = roi_x; t->remap.roi_y = roi_y;
  }
  // staging buffers hold source frames: resize them for the (larger) distorted input
  const size_t F = (size_t)t->cfg.max_frames;
  t->stage_frames = stage_capacity(F, frame_bytes);
  for (int i = 0; i < 2; ++i) {
    cudaFree(t->d_stage[i]);
    t->d_stage[i] = nullptr;
    t->stage_busy[i] = false;
    UWT_CUDA(t, cudaMalloc(&t->d_stage[i], t->stage_frames * frame_bytes));
  }
  return UWT_OK;
}

int uwt_apply_gradient(uwt_tracker* t, int n, const int* slots) {
  int rc = check_slots(t, n, slots);
  if (rc) return rc;
  for (int i = 0; i < n; ++i)
    if (!t->slots[slots[i]].pyramid)
      return fail(t, UWT_E_STATE, "slot %d has no frame", slots[i]);
  // the fused frame kernel has already produced the gradient images of freshly uploaded slots
  {
    bool have = true;
    for (int i = 0; i < n; ++i) have = have && t->slots[slots[i]].gradient_all;
    if (have) return UWT_OK;
  }
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  ArgRegion* r = nullptr;
  if ((rc = acquire(t, &r))) return rc;
  if ((rc = push_slots(t, r, n, slots, nullptr))) return rc;
  ProfSpan span(t, UWT_K_GRADIENT);
  const bool lazy = (t->cfg.flags & UWT_FLAG_LAZY_LEVELS) != 0;
  const LevelRange lr = lazy ? level_range(t->geom, t->cfg.last_level, t->cfg.first_level)
                             : level_range(t->geom, 0, t->geom.levels - 1);
  const int k = launch_gradient(t->geom, t->pools, n, r->d_int, t->stream, lr);
  span.done(k);
  if (k < 0) return fail(t, UWT_E_CUDA, "gradient kernel launch failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
  t->launches += k;
  if ((rc = release(t, r))) return rc;
  for (int i = 0; i < n; ++i) {
    t->slots[slots[i]].gradient = true;
    t->slots[slots[i]].gradient_all = !lazy;
    t->slots[slots[i]].candidates = t->slots[slots[i]].candidates_all = false;
  }
  return UWT_OK;
}

int uwt_select_candidates(uwt_tracker* t, int n, const int* slots) {
  int rc = check_slots(t, n, slots);
  if (rc) return rc;
  for (int i = 0; i < n; ++i)
    if (!t->slots[slots[i]].gradient)
      return fail(t, UWT_E_STATE, "slot %d has no gradients (call uwt_apply_gradient)", slots[i]);
  for (int i = 0; i < n; ++i)
    if (t->cfg.depth_mode != UWT_DEPTH_NONE && !t->slots[slots[i]].depth)
      return fail(t, UWT_E_STATE, "slot %d has no depth frame (call uwt_upload_depth_frames)",
                  slots[i]);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  ArgRegion* r = nullptr;
  if ((rc = acquire(t, &r))) return rc;
  if ((rc = push_slots(t, r, n, slots, nullptr))) return rc;
  ProfSpan span(t, UWT_K_CANDIDATES);
  const bool lazy = (t->cfg.flags & UWT_FLAG_LAZY_LEVELS) != 0;
  const LevelRange lr = lazy ? level_range(t->geom, t->cfg.last_level, t->cfg.first_level)
                             : level_range(t->geom, 0, t->geom.levels - 1);
  const int k = launch_candidates(t->geom, t->pools, n, r->d_int, t->stream, lr, !lazy);
  span.done(k);
  if (k < 0) return fail(t, UWT_E_CUDA, "candidate kernel launch failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
  t->launches += k;
  if ((rc = release(t, r))) return rc;
  for (int i = 0; i < n; ++i) {
    t->slots[slots[i]].candidates = true;
    t->slots[slots[i]].candidates_all = !lazy;
  }
  return UWT_OK;
}

static int pick_cluster(const uwt_tracker* t, int n) {
  int c = t->cfg.cluster_size;
  if (c == 0) {
    // Enough CTAs that the tail of unequal problems (different iteration counts) averages
    // out over >= 2 CTAs per SM, but no more: every extra CTA of a cluster repeats the
    // per-sweep table build / reduction / solve.  Measured on B200 (128 problems at
    // 1280x1024): C=1 1.57 ms, 2 1.43, 4 1.22, 8 1.31, 16 1.86 per launch.
    c = 1;
    while (c < 16 && n * c < 2 * 148) c *= 2;
  }
  return std::min(c, t->max_cluster);
}

int uwt_estimate_pose_async(uwt_tracker* t, int n, const int* prev_slots, const int* cur_slots,
                            const float* init_poses7) {
  int rc = check_slots(t, n, prev_slots);
  if (rc) return rc;
  if ((rc = check_slots(t, n, cur_slots))) return rc;
  for (int i = 0; i < n; ++i) {
    if (!t->slots[prev_slots[i]].candidates)
      return fail(t, UWT_E_STATE, "prev slot %d has no candidate points", prev_slots[i]);
    if (!t->slots[cur_slots[i]].pyramid)
      return fail(t, UWT_E_STATE, "cur slot %d has no frame", cur_slots[i]);
  }
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  // ONE large problem (>= 1 Mpixel on the finest optimised level): the persistent whole-GPU
  // kernel (148 CTAs, grid barrier) beats the 16-CTA cluster (measured at 3840x2160: 187 vs
  // 336 us); smaller frames stay on the cluster kernel (117 vs 165 us at 1280x1024).
  {
    const LevelGeom& Lf = t->geom.lv[t->cfg.last_level];
    if (n == 1 && t->cfg.cluster_size == 0 && !(t->cfg.flags & UWT_FLAG_TRACE) &&
        t->cfg.weight_mode == UWT_WEIGHT_IDENTITY && t->cfg.depth_mode == UWT_DEPTH_NONE &&
        t->cfg.sampling == UWT_SAMPLE_NEAREST && (long long)Lf.w * Lf.h >= (1 << 20)) {
      ShardState s;
      std::memset(&s, 0, sizeof(s));
      const float ident[7] = {0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};
      std::memcpy(s.pose, init_poses7 ? init_poses7 : ident, sizeof(ident));
      s.last_error = 50000.0f;
      s.level = t->cfg.first_level;
      s.rank = 0;
      s.nranks = 1;
      s.prev_slot = prev_slots[0];
      s.cur_slot = cur_slots[0];
      ArgRegion* r = nullptr;
      if ((rc = acquire(t, &r))) return rc;
      static_assert(sizeof(ShardState) <= sizeof(float) * 7 * 16, "staging too small");
      if ((size_t)t->cfg.max_frames * 7 * sizeof(float) >= sizeof(ShardState)) {
        std::memcpy(r->h_flt, &s, sizeof(s));  // pinned staging: no host sync
        UWT_CUDA(t, cudaMemcpyAsync(t->d_shard, r->h_flt, sizeof(s), cudaMemcpyHostToDevice,
                                    t->stream));
      } else {
        UWT_CUDA(t, cudaMemcpyAsync(t->d_shard, &s, sizeof(s), cudaMemcpyHostToDevice, t->stream));
        UWT_CUDA(t, cudaStreamSynchronize(t->stream));
      }
      ProfSpan span(t, UWT_K_ESTIMATE);
      // as many persistent CTAs as the device holds at once (cooperative launch)
      const int k = launch_shard_fused(t->geom, t->pools, t->d_shard, t->d_fused_self,
                                       t->d_mailbox_self, t->d_shard_partials, t->shard_grid,
                                       t->stream);
      if (k >= 0) {
        t->launches += k;
        span.done(k);
        UWT_CUDA(t, cudaMemcpyAsync(t->h_out_poses, reinterpret_cast<char*>(t->d_shard) +
                                    offsetof(ShardState, pose), sizeof(float) * 7,
                                    cudaMemcpyDeviceToHost, t->stream));
        UWT_CUDA(t, cudaMemcpyAsync(t->h_stats, reinterpret_cast<char*>(t->d_shard) +
                                    offsetof(ShardState, stats), sizeof(uwt_track_stats),
                                    cudaMemcpyDeviceToHost, t->stream));
        UWT_CUDA(t, cudaEventRecord(t->poses_ready, t->stream));
        if ((rc = release(t, r))) return rc;
        t->last_n = 1;
        t->shard_active = false;
        t->flow_last = false;
        t->last_traced = false;
        return UWT_OK;
      }
      // the cooperative launch was refused (fewer co-resident CTAs, shared memory): clear the
      // error and let the cluster kernel below serve the problem
      cudaGetLastError();
      if ((rc = release(t, r))) return rc;
    }
  }
  ArgRegion* r = nullptr;
  if ((rc = acquire(t, &r))) return rc;
  if ((rc = push_slots(t, r, n, prev_slots, cur_slots))) return rc;
  if (init_poses7) {
    std::memcpy(r->h_flt, init_poses7, sizeof(float) * 7 * n);
    if ((rc = push_words(t, r->d_flt, r->h_flt, (size_t)7 * n))) return rc;
  }
  EstimateIO io;
  io.prev_slots = r->d_int;
  io.cur_slots = r->d_int + n;
  io.init_poses = init_poses7 ? r->d_flt : nullptr;
  io.out_poses = t->d_out_poses;
  io.stats = t->d_stats;
  const bool tracing = t->d_trace && n <= kTraceProblems;
  io.trace = tracing ? t->d_trace : nullptr;
  io.trace_count = tracing ? t->d_trace_count : nullptr;
  io.trace_cap = t->trace_cap;
  int cluster = pick_cluster(t, n);
  // Batches: the persistent dataflow kernel (chunk tasks, no per-sweep barriers, no tail of
  // unequal problems), robust weights, depth input and bilinear sampling included.  Few
  // problems, the DMMA A/B variant and an explicit cluster size stay on the cluster kernel.
  // From 2 CTAs per SM upwards one-CTA problems balance by themselves -- as long as a problem is
  // small.  Large frames (finest optimised level >= 160 k pixels, e.g. 1280x1024) stay on the
  // dataflow kernel at any batch size: 384 / 512 problems measured 5.5 vs 8.4 us per problem.
  const LevelGeom& finest = t->geom.lv[t->geom.last_level];
  const bool large_frames = (long long)finest.w * finest.h >= 160 * 1024;
  const bool use_flow = n >= kFlowMinProblems &&
                        (n < kFlowMaxProblems || (large_frames && n <= kFlowMaxLargeProblems)) &&
                        t->cfg.cluster_size == 0 &&
                        !(t->cfg.flags & (UWT_FLAG_DMMA_ACCUM | UWT_FLAG_CLUSTER_KERNEL));
  if (use_flow) {
    const size_t need = flow_workspace_bytes(t->geom, n);
    if (need > t->flow_ws_bytes) {
      UWT_CUDA(t, cudaStreamSynchronize(t->stream));
      cudaFree(t->d_flow_ws);
      t->d_flow_ws = nullptr;
      t->flow_ws_bytes = 0;
      UWT_CUDA(t, cudaMalloc(&t->d_flow_ws, need));
      t->flow_ws_bytes = need;
    }
  }
  ProfSpan span(t, UWT_K_ESTIMATE);
  const int variant = (t->cfg.flags & UWT_FLAG_DMMA_ACCUM) ? UWT_EST_MMA : UWT_EST_REGISTERS;
  int k = -2;
  bool flow_used = false;
  if (use_flow) {
    k = launch_estimate_flow(t->geom, t->pools, n, io, t->d_flow_ws, t->stream, &t->flow_grid);
    flow_used = k > 0;
    if (k < 0) {  // not launchable here (task-word limits, shared memory): the cluster kernel can
      cudaGetLastError();
      k = -2;
    }
  }
  if (k == -2) k = launch_estimate(t->geom, t->pools, n, io, cluster, t->stream, variant);
  while (k < 0 && cluster > 1 && !flow_used) {  // a 16-CTA cluster may not be schedulable on every part
    cudaGetLastError();
    cluster /= 2;
    t->max_cluster = cluster;
    k = launch_estimate(t->geom, t->pools, n, io, cluster, t->stream, variant);
  }
  if (k < 0) return fail(t, UWT_E_CUDA, "estimate kernel launch failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
  t->launches += k;
  span.done(k);
  UWT_CUDA(t, cudaMemcpyAsync(t->h_out_poses, t->d_out_poses, sizeof(float) * 7 * n,
                              cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaMemcpyAsync(t->h_stats, t->d_stats, sizeof(uwt_track_stats) * n,
                              cudaMemcpyDeviceToHost, t->stream));
  t->flow_last = flow_used;
  t->last_traced = tracing;
  if (t->flow_last)
    UWT_CUDA(t, cudaMemcpyAsync(t->h_flow_ctl, t->d_flow_ws, sizeof(int) * 4,
                                cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaEventRecord(t->poses_ready, t->stream));
  if ((rc = release(t, r))) return rc;
  t->last_n = n;
  return UWT_OK;
}

int uwt_fetch_poses(uwt_tracker* t, int n, float* out_poses7, uwt_track_stats* stats) {
  if (!t) return UWT_E_INVALID;
  if (n <= 0 || n > t->last_n) return fail(t, UWT_E_INVALID, "n=%d exceeds the last batch (%d)", n, t->last_n);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  // waits for the estimate + its D2H only, not for work enqueued after it
  UWT_CUDA(t, cudaEventSynchronize(t->poses_ready));
#ifdef UWT_FLOW_STATS
  if (t->flow_last) flow_debug_dump();
#endif
  if (t->flow_last && (t->h_flow_ctl[3] != 0 || t->h_flow_ctl[2] != 0))
    return fail(t, UWT_E_CUDA, "dataflow estimate kernel did not complete (error %d, %d problems "
                "unfinished)", t->h_flow_ctl[3], t->h_flow_ctl[2]);
  if (out_poses7) std::memcpy(out_poses7, t->h_out_poses, sizeof(float) * 7 * n);
  if (stats) std::memcpy(stats, t->h_stats, sizeof(uwt_track_stats) * n);
  return UWT_OK;
}

int uwt_estimate_pose(uwt_tracker* t, int n, const int* prev_slots, const int* cur_slots,
                      const float* init_poses7, float* out_poses7, uwt_track_stats* stats) {
  int rc = uwt_estimate_pose_async(t, n, prev_slots, cur_slots, init_poses7);
  if (rc) return rc;
  return uwt_fetch_poses(t, n, out_poses7, stats);
}

int uwt_shard_begin(uwt_tracker* t, int prev_slot, int cur_slot, int rank, int nranks,
                    const float* init_pose7) {
  int rc = check_slots(t, 1, &prev_slot);
  if (rc) return rc;
  if ((rc = check_slots(t, 1, &cur_slot))) return rc;
  if (nranks < 1 || rank < 0 || rank >= nranks)
    return fail(t, UWT_E_INVALID, "bad shard rank %d of %d", rank, nranks);
  if (t->cfg.weight_mode == UWT_WEIGHT_TUKEY || t->cfg.depth_mode != UWT_DEPTH_NONE ||
      t->cfg.sampling != UWT_SAMPLE_NEAREST)
    return fail(t, UWT_E_INVALID,
                "the sharded mode supports identity or Huber weights, mono input, nearest "
                "sampling only (Tukey/MAD weights need the sweep's residual histogram)");
  if (!t->slots[prev_slot].candidates)
    return fail(t, UWT_E_STATE, "prev slot %d has no candidate points", prev_slot);
  if (!t->slots[cur_slot].pyramid) return fail(t, UWT_E_STATE, "cur slot %d has no frame", cur_slot);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  // pinned staging (ring of 4, an event per entry): the state travels without a host sync
  const int si = t->h_shard_next;
  t->h_shard_next = (si + 1) & 3;
  UWT_CUDA(t, cudaEventSynchronize(t->h_shard_ev[si]));  // its previous copy has been consumed
  ShardState& s = *t->h_shard_in[si];
  std::memset(&s, 0, sizeof(s));
  // SE3::exp(0) is exactly the identity (Tracker.cpp:385)
  const float ident[7] = {0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};
  std::memcpy(s.pose, init_pose7 ? init_pose7 : ident, sizeof(ident));
  s.last_error = 50000.0f;
  s.level = t->cfg.first_level;
  s.rank = rank;
  s.nranks = nranks;
  s.prev_slot = prev_slot;
  s.cur_slot = cur_slot;
  UWT_CUDA(t, cudaMemcpyAsync(t->d_shard, &s, sizeof(s), cudaMemcpyHostToDevice, t->stream));
  UWT_CUDA(t, cudaEventRecord(t->h_shard_ev[si], t->stream));
  t->shard_active = true;
  return UWT_OK;
}

int uwt_shard_accumulate(uwt_tracker* t, double* d_sums32) {
  if (!t) return UWT_E_INVALID;
  if (!t->shard_active) return fail(t, UWT_E_STATE, "call uwt_shard_begin first");
  if (!d_sums32) return fail(t, UWT_E_INVALID, "d_sums32 is NULL");
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  ProfSpan span(t, UWT_K_ESTIMATE);
  const int k = launch_shard_accumulate(t->geom, t->pools, t->d_shard, t->d_shard_partials,
                                        d_sums32, kShardMaxGrid, t->stream);
  if (k < 0) return fail(t, UWT_E_CUDA, "shard accumulate launch failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
  t->launches += k;
  span.done(k);
  return UWT_OK;
}

int uwt_shard_update(uwt_tracker* t, const double* d_sums32, int* done) {
  if (!t) return UWT_E_INVALID;
  if (!t->shard_active) return fail(t, UWT_E_STATE, "call uwt_shard_begin first");
  if (!d_sums32 || !done) return fail(t, UWT_E_INVALID, "NULL argument");
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  const int k = launch_shard_update(t->geom, t->pools, t->d_shard, d_sums32, t->d_shard_done,
                                    t->stream);
  if (k < 0) return fail(t, UWT_E_CUDA, "shard update launch failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
  t->launches += k;
  UWT_CUDA(t, cudaMemcpyAsync(t->h_shard_done, t->d_shard_done, sizeof(int),
                              cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  *done = *t->h_shard_done;
  return UWT_OK;
}

int uwt_shard_result(uwt_tracker* t, float* out_pose7, uwt_track_stats* stats) {
  if (!t) return UWT_E_INVALID;
  if (!t->shard_active) return fail(t, UWT_E_STATE, "call uwt_shard_begin first");
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  ShardState s;
  UWT_CUDA(t, cudaMemcpyAsync(&s, t->d_shard, sizeof(s), cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  if (!s.done) return fail(t, UWT_E_STATE, "sharded estimate has not finished");
  if (out_pose7) std::memcpy(out_pose7, s.pose, sizeof(s.pose));
  if (stats) *stats = s.stats;
  return UWT_OK;
}

int uwt_shard_ipc_handle_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int uwt_shard_ipc_export(uwt_tracker* t, void* handle_out) {
  if (!t || !handle_out) return UWT_E_INVALID;
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  cudaIpcMemHandle_t h;
  UWT_CUDA(t, cudaIpcGetMemHandle(&h, t->d_mailbox));
  std::memcpy(handle_out, &h, sizeof(h));
  return UWT_OK;
}

static int fused_publish(uwt_tracker* t, int rank, int nranks) {
  t->fused_rank = rank;
  t->fused_nranks = nranks;
  // reset the sequence state: every rank does this at connect time, before any sweep
  t->h_fused.seq = 0;
  t->h_fused.generation = 0;
  t->h_fused.error = 0;
  UWT_CUDA(t, cudaMemsetAsync(t->d_mailbox, 0, sizeof(ShardMailbox), t->stream));
  UWT_CUDA(t, cudaMemcpyAsync(t->d_fused, &t->h_fused, sizeof(ShardFused), cudaMemcpyHostToDevice,
                              t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  return UWT_OK;
}

int uwt_shard_ipc_connect(uwt_tracker* t, int rank, int nranks, const void* all_handles) {
  if (!t) return UWT_E_INVALID;
  if (nranks < 1 || nranks > kShardMaxRanks || rank < 0 || rank >= nranks || !all_handles)
    return fail(t, UWT_E_INVALID, "bad rank %d / nranks %d (max %d)", rank, nranks, kShardMaxRanks);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  for (void*& p : t->ipc_opened)
    if (p) {
      cudaIpcCloseMemHandle(p);
      p = nullptr;
    }
  const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      t->h_fused.peer[r] = t->d_mailbox;
    } else {
      void* p = nullptr;
      UWT_CUDA(t, cudaIpcOpenMemHandle(&p, hs[r], cudaIpcMemLazyEnablePeerAccess));
      t->ipc_opened[r] = p;
      t->h_fused.peer[r] = static_cast<ShardMailbox*>(p);
    }
  }
  return fused_publish(t, rank, nranks);
}

int uwt_shard_connect_local(uwt_tracker* t, int rank, int nranks, uwt_tracker* const* peers) {
  if (!t) return UWT_E_INVALID;
  if (nranks < 1 || nranks > kShardMaxRanks || rank < 0 || rank >= nranks || !peers)
    return fail(t, UWT_E_INVALID, "bad rank %d / nranks %d (max %d)", rank, nranks, kShardMaxRanks);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  for (int r = 0; r < nranks; ++r) {
    if (!peers[r]) return fail(t, UWT_E_INVALID, "peer %d is NULL", r);
    if (peers[r]->cfg.device != t->cfg.device) {
      int can = 0;
      UWT_CUDA(t, cudaDeviceCanAccessPeer(&can, t->cfg.device, peers[r]->cfg.device));
      if (!can) return fail(t, UWT_E_CUDA, "device %d cannot access peer %d", t->cfg.device,
                            peers[r]->cfg.device);
      cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->cfg.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(t, UWT_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
      cudaGetLastError();
    }
    t->h_fused.peer[r] = peers[r]->d_mailbox;
  }
  return fused_publish(t, rank, nranks);
}

int uwt_shard_estimate_fused_async(uwt_tracker* t, int prev_slot, int cur_slot,
                                   const float* init_pose7, int grid) {
  if (!t) return UWT_E_INVALID;
  if (t->fused_rank < 0)
    return fail(t, UWT_E_STATE, "call uwt_shard_ipc_connect / uwt_shard_connect_local first");
  int rc = uwt_shard_begin(t, prev_slot, cur_slot, t->fused_rank, t->fused_nranks, init_pose7);
  if (rc) return rc;
  if (grid <= 0) grid = t->shard_grid;
  if (grid > kShardMaxGrid) grid = kShardMaxGrid;
  ProfSpan span(t, UWT_K_ESTIMATE);
  const int k = launch_shard_fused(t->geom, t->pools, t->d_shard, t->d_fused, t->d_mailbox,
                                   t->d_shard_partials, grid, t->stream);
  if (k < 0) return fail(t, UWT_E_CUDA, "fused shard kernel launch failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
  t->launches += k;
  span.done(k);
  // result and control block follow the kernel into pinned memory: the wait is ONE event
  UWT_CUDA(t, cudaMemcpyAsync(t->h_shard_out, t->d_shard, sizeof(ShardState),
                              cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaMemcpyAsync(t->h_fused_out, t->d_fused, sizeof(ShardFused),
                              cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaEventRecord(t->fused_done, t->stream));
  return UWT_OK;
}

int uwt_shard_estimate_fused_wait(uwt_tracker* t, float* out_pose7, uwt_track_stats* stats) {
  if (!t) return UWT_E_INVALID;
  if (!t->shard_active) return fail(t, UWT_E_STATE, "call uwt_shard_estimate_fused_async first");
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  UWT_CUDA(t, cudaEventSynchronize(t->fused_done));
  const ShardFused& f = *t->h_fused_out;
  if (getenv("UWT_DEBUG") && f.dbg[5])
    fprintf(stderr, "[uwt fused] sweeps %llu | cycles/sweep: own-accumulate %llu, until-all-CTAs %llu, "
            "reduce %llu, mailbox %llu, K5 %llu\n", f.dbg[5], f.dbg[0] / f.dbg[5], f.dbg[1] / f.dbg[5],
            f.dbg[2] / f.dbg[5], f.dbg[3] / f.dbg[5], f.dbg[4] / f.dbg[5]);
  if (f.error)
    return fail(t, UWT_E_CUDA, "fused sharded estimate: a peer did not arrive (bounded wait expired)");
  const ShardState& r = *t->h_shard_out;
  if (!r.done) return fail(t, UWT_E_STATE, "sharded estimate has not finished");
  if (out_pose7) std::memcpy(out_pose7, r.pose, sizeof(r.pose));
  if (stats) *stats = r.stats;
  return UWT_OK;
}

int uwt_warp_points(uwt_tracker* t, const float* pts4, int n, const float* pose7, int level,
                    float* out4) {
  if (!t) return UWT_E_INVALID;
  if (!pts4 || !pose7 || !out4 || n < 0 || level < 0 || level >= t->geom.levels)
    return fail(t, UWT_E_INVALID, "bad argument to uwt_warp_points");
  if (n == 0) return UWT_OK;
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  const size_t pts_bytes = align_up(sizeof(float) * 4 * (size_t)n, 256);
  char* base = static_cast<char*>(scratch(t, 2 * pts_bytes + 256));
  if (!base) return fail(t, UWT_E_NOMEM, "out of device memory for %d points", n);
  float* d_pts = reinterpret_cast<float*>(base);
  float* d_out = reinterpret_cast<float*>(base + pts_bytes);
  float* d_pose = reinterpret_cast<float*>(base + 2 * pts_bytes);
  UWT_CUDA(t, cudaMemcpyAsync(d_pts, pts4, sizeof(float) * 4 * n, cudaMemcpyHostToDevice,
                              t->stream));
  UWT_CUDA(t, cudaMemcpyAsync(d_pose, pose7, sizeof(float) * 7, cudaMemcpyHostToDevice,
                              t->stream));
  const int k = launch_warp_points(t->geom, d_pts, n, d_pose, level, d_out, t->stream);
  if (k < 0) return fail(t, UWT_E_CUDA, "warp kernel launch failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
  t->launches += k;
  UWT_CUDA(t, cudaMemcpyAsync(out4, d_out, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost,
                              t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  return UWT_OK;
}

// UWT_FLAG_LAZY_LEVELS: a read-back of a level outside [last_level, first_level] materialises
// gradient (and, if they were selected, candidates) of ALL levels of that slot, with the same
// kernels -- so every accessor returns exactly what the eager mode would.
static int complete_levels(uwt_tracker* t, int slot, int level, bool need_candidates) {
  SlotState& s = t->slots[slot];
  const bool inside = level >= t->cfg.last_level && level <= t->cfg.first_level;
  if (inside) return UWT_OK;
  const bool do_grad = s.gradient && !s.gradient_all;
  const bool do_cand = need_candidates && s.candidates && !s.candidates_all;
  if (!do_grad && !do_cand) return UWT_OK;
  ArgRegion* r = nullptr;
  int rc = acquire(t, &r);
  if (rc) return rc;
  if ((rc = push_slots(t, r, 1, &slot, nullptr))) return rc;
  const LevelRange all = level_range(t->geom, 0, t->geom.levels - 1);
  int k = 0;
  if (do_grad) {
    k = launch_gradient(t->geom, t->pools, 1, r->d_int, t->stream, all);
    if (k < 0) return fail(t, UWT_E_CUDA, "gradient kernel launch failed");
    t->launches += k;
    s.gradient_all = true;
  }
  if (do_cand) {
    k = launch_candidates(t->geom, t->pools, 1, r->d_int, t->stream, all, true);
    if (k < 0) return fail(t, UWT_E_CUDA, "candidate kernel launch failed");
    t->launches += k;
    s.candidates_all = true;
  }
  return release(t, r);
}

static int check_read(uwt_tracker* t, int slot, int level) {
  if (!t) return UWT_E_INVALID;
  if (slot < 0 || slot >= t->cfg.max_frames || level < 0 || level >= t->geom.levels)
    return fail(t, UWT_E_INVALID, "slot %d / level %d out of range", slot, level);
  return UWT_OK;
}

int uwt_get_image(uwt_tracker* t, int slot, int level, uint8_t* host) {
  int rc = check_read(t, slot, level);
  if (rc) return rc;
  if (!t->slots[slot].pyramid) return fail(t, UWT_E_STATE, "slot %d has no frame", slot);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  const LevelGeom& L = t->geom.lv[level];
  const size_t off = (size_t)slot * t->geom.plane_elems + L.plane_off;
  UWT_CUDA(t, cudaMemcpy2DAsync(host, L.w, t->pools.img + off, L.pitch, L.w, L.h,
                                cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  return UWT_OK;
}

int uwt_get_gradients(uwt_tracker* t, int slot, int level, int16_t* gx, int16_t* gy, uint8_t* g) {
  int rc = check_read(t, slot, level);
  if (rc) return rc;
  if (!t->slots[slot].gradient) return fail(t, UWT_E_STATE, "slot %d has no gradients", slot);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  if ((rc = complete_levels(t, slot, level, false))) return rc;
  const LevelGeom& L = t->geom.lv[level];
  const size_t off = (size_t)slot * t->geom.plane_elems + L.plane_off;
  if (gx || gy) {
    // gradientX_/gradientY_ planes are not kept on the device (the tracker consumes packed
    // records): materialise them for this slot with the same kernel and copy the level out
    const size_t pe = t->geom.plane_elems;
    int16_t* tmp = static_cast<int16_t*>(scratch(t, 2 * pe * sizeof(int16_t)));
    if (!tmp) return fail(t, UWT_E_NOMEM, "out of device memory for the gradient planes");
    ArgRegion* r = nullptr;
    if ((rc = acquire(t, &r))) return rc;
    if ((rc = push_slots(t, r, 1, &slot, nullptr))) return rc;
    const int k = launch_gradient(t->geom, t->pools, 1, r->d_int, t->stream,
                                  level_range(t->geom, 0, t->geom.levels - 1), tmp, tmp + pe);
    if (k < 0) return fail(t, UWT_E_CUDA, "gradient kernel launch failed: %s",
                           cudaGetErrorString(cudaGetLastError()));
    t->launches += k;
    if ((rc = release(t, r))) return rc;
    if (gx)
      UWT_CUDA(t, cudaMemcpy2DAsync(gx, L.w * 2, tmp + L.plane_off, L.pitch * 2, L.w * 2, L.h,
                                    cudaMemcpyDeviceToHost, t->stream));
    if (gy)
      UWT_CUDA(t, cudaMemcpy2DAsync(gy, L.w * 2, tmp + pe + L.plane_off, L.pitch * 2, L.w * 2, L.h,
                                    cudaMemcpyDeviceToHost, t->stream));
  }
  if (g)
    UWT_CUDA(t, cudaMemcpy2DAsync(g, L.w, t->pools.g + off, L.pitch, L.w, L.h,
                                  cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  return UWT_OK;
}

int uwt_get_candidate_count(uwt_tracker* t, int slot, int level, int* n) {
  int rc = check_read(t, slot, level);
  if (rc) return rc;
  if (!n) return fail(t, UWT_E_INVALID, "n is NULL");
  if (!t->slots[slot].candidates) return fail(t, UWT_E_STATE, "slot %d has no candidates", slot);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  if ((rc = complete_levels(t, slot, level, true))) return rc;
  uint32_t v = 0;
  UWT_CUDA(t, cudaMemcpyAsync(&v, t->pools.ncand + (size_t)slot * kMaxLevels + level, sizeof(v),
                              cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  *n = (int)v;
  return UWT_OK;
}

int uwt_get_candidates(uwt_tracker* t, int slot, int level, float* pts4, int capacity_rows,
                       int* n) {
  int cnt = 0;
  int rc = uwt_get_candidate_count(t, slot, level, &cnt);
  if (rc) return rc;
  if (n) *n = cnt;
  if (!pts4) return UWT_OK;
  if (cnt > capacity_rows)
    return fail(t, UWT_E_INVALID, "capacity %d < %d candidates", capacity_rows, cnt);
  if (cnt == 0) return UWT_OK;
  const LevelGeom& L = t->geom.lv[level];
  // depth modes: Z = depth * factor (Tracker.cpp:1344) from the level's depth plane
  std::vector<uint16_t> dplane;
  if (t->cfg.depth_mode != UWT_DEPTH_NONE) {
    dplane.resize((size_t)L.w * L.h);
    rc = uwt_get_depth(t, slot, level, dplane.data());
    if (rc) return rc;
  }
  // Tracker.cpp:1316; ObtainAllPoints: factor / 2^level (Tracker.cpp:1266)
  const float factor =
      t->cfg.depth_mode == UWT_DEPTH_ALL_POINTS ? std::ldexp(0.0002f, -level) : 0.0002f;
  auto z_of = [&](int x, int y) -> float {
    if (dplane.empty()) return 1.0f;  // depth_initialization, Tracker.cpp:1317
    int d;
    if (t->cfg.depth_mode == UWT_DEPTH_REFERENCE) {
      const unsigned v = dplane[(size_t)y * L.w + (x >> 1)];
      d = (x & 1) ? (int)(v >> 8) : (int)(v & 0xFFu);
    } else if (t->cfg.depth_mode == UWT_DEPTH_ALL_POINTS) {
      d = (int)(int16_t)dplane[(size_t)y * L.w + x];
    } else {
      d = dplane[(size_t)y * L.w + x];
    }
    return d * factor;
  };
  if (L.rec_off >= 0) {
    // optimised levels keep only the packed records; (x, y) are their low 24 bits
    std::vector<uint64_t> rec(cnt);
    UWT_CUDA(t, cudaMemcpyAsync(rec.data(),
                                t->pools.rec + (size_t)slot * t->geom.rec_elems + L.rec_off,
                                sizeof(uint64_t) * cnt, cudaMemcpyDeviceToHost, t->stream));
    UWT_CUDA(t, cudaStreamSynchronize(t->stream));
    for (int i = 0; i < cnt; ++i) {  // candidatePoints_ rows, Tracker.cpp:1351-1355
      pts4[i * 4 + 0] = (float)(rec[i] & 0xFFFu);
      pts4[i * 4 + 1] = (float)((rec[i] >> 12) & 0xFFFu);
      pts4[i * 4 + 2] = z_of((int)(rec[i] & 0xFFFu), (int)((rec[i] >> 12) & 0xFFFu));
      pts4[i * 4 + 3] = 1.0f;
    }
    return UWT_OK;
  }
  // a level without records: expand the selection bitmask in the reference's order, x outer and
  // y inner (Tracker.cpp:1334-1335)
  std::vector<uint32_t> mask((size_t)L.mask_wpr * L.h);
  UWT_CUDA(t, cudaMemcpyAsync(mask.data(),
                              t->pools.sel_mask + (size_t)slot * t->geom.mask_elems + L.mask_off,
                              sizeof(uint32_t) * mask.size(), cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  int i = 0;
  for (int x = 0; x < L.w && i < cnt; ++x)
    for (int y = 0; y < L.h && i < cnt; ++y)
      if ((mask[(size_t)y * L.mask_wpr + (x >> 5)] >> (x & 31)) & 1u) {
        pts4[i * 4 + 0] = (float)x;  // candidatePoints_ rows, Tracker.cpp:1351-1355
        pts4[i * 4 + 1] = (float)y;
        pts4[i * 4 + 2] = z_of(x, y);
        pts4[i * 4 + 3] = 1.0f;
        ++i;
      }
  if (i != cnt) return fail(t, UWT_E_CUDA, "selection bitmask holds %d points, count says %d", i, cnt);
  return UWT_OK;
}

int uwt_get_records(uwt_tracker* t, int slot, int level, uint64_t* packed, int capacity, int* n) {
  int cnt = 0;
  int rc = uwt_get_candidate_count(t, slot, level, &cnt);
  if (rc) return rc;
  const LevelGeom& L = t->geom.lv[level];
  if (L.rec_off < 0)
    return fail(t, UWT_E_INVALID, "level %d is not optimised: it has no packed records", level);
  if (n) *n = cnt;
  if (!packed) return UWT_OK;
  if (cnt > capacity) return fail(t, UWT_E_INVALID, "capacity %d < %d records", capacity, cnt);
  if (cnt == 0) return UWT_OK;
  UWT_CUDA(t, cudaMemcpyAsync(packed, t->pools.rec + (size_t)slot * t->geom.rec_elems + L.rec_off,
                              sizeof(uint64_t) * cnt, cudaMemcpyDeviceToHost, t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  return UWT_OK;
}

int uwt_get_trace(uwt_tracker* t, int index, uwt_iter_trace* out, int capacity, int* n) {
  if (!t) return UWT_E_INVALID;
  if (!t->d_trace) return fail(t, UWT_E_STATE, "tracker was not created with UWT_FLAG_TRACE");
  if (index < 0 || index >= t->last_n || index >= kTraceProblems || !out || !n)
    return fail(t, UWT_E_INVALID, "bad trace index %d", index);
  if (!t->last_traced)
    return fail(t, UWT_E_STATE, "the last estimate was not traced (UWT_FLAG_TRACE records batches "
                "of at most %d problems, and not the single-large-frame path)", kTraceProblems);
  UWT_CUDA(t, cudaSetDevice(t->cfg.device));
  int cnt = 0;
  UWT_CUDA(t, cudaMemcpyAsync(&cnt, t->d_trace_count + index, sizeof(int), cudaMemcpyDeviceToHost,
                              t->stream));
  UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  cnt = std::min(cnt, capacity);
  if (cnt > 0) {
    UWT_CUDA(t, cudaMemcpyAsync(out, t->d_trace + (size_t)index * t->trace_cap,
                                sizeof(uwt_iter_trace) * cnt, cudaMemcpyDeviceToHost, t->stream));
    UWT_CUDA(t, cudaStreamSynchronize(t->stream));
  }
  *n = cnt;
  return UWT_OK;
}

}  // extern "C"
