// uwt_estimate_flow.cu -- batched Gauss-Newton (Tracker::EstimatePose, Tracker.cpp:362-597, for
// many independent problems) as a persistent dataflow kernel.
#include "uwt_estimate_common.cuh"

namespace uwt {

// ----------------------------------------------------------------------------------------
// Batched Gauss-Newton as a persistent DATAFLOW kernel (many independent problems).
//
// The cluster kernel (uwt_estimate.cu) gives every problem a fixed set of CTAs for its whole life: CTAs idle
// at every per-sweep barrier, during the serial solve, and when problems of a wave finish at
// different times (measured: 20 % of the launch is tail, 14 % barrier stalls).  Here the unit of
// scheduling is one CHUNK of one residual sweep (kFlowChunk consecutive candidate records of one
// problem at its current level and pose).  Persistent CTAs pop chunk tasks from a ring in global
// memory; the CTA that completes the last chunk of a sweep reduces the per-chunk partial sums in
// chunk order (deterministic), runs the warp-collective update (break test, 6x6 LU, SE3 exp)
// for that problem and enqueues the chunks of its next sweep.  No grid- or cluster-wide barrier
// exists: a problem's update overlaps every other problem's streaming, chunks are equal-sized,
// and the GPU drains only when the last problems run out of sweeps.
//   * the per-task transform tables tab_x[3][kTab], tab_y[3][kTab] (compile-time row stride, all
//     columns and rows of the level) are rebuilt per task from the problem's pose: no dependent
//     read of the chunk's records sits on the hand-over path
//   * the point loop is the software-pipelined fast sweep of uwt_estimate_common.cuh (mono input,
//     nearest sampling; depth has its own pipelined sweep, everything else the generic loop)
//   * waits are bounded by construction: a consumer spins only on a ring slot whose producer is
//     a CTA that holds a real task, and leaves when the count of unfinished problems is zero
//   * arithmetic, and therefore every result, is identical to the cluster kernel's
// ----------------------------------------------------------------------------------------
#ifndef UWT_FLOW_THREADS
#define UWT_FLOW_THREADS 256
#endif
#ifndef UWT_FLOW_CHUNK
#define UWT_FLOW_CHUNK 8192
#endif
constexpr int kFlowThreads = UWT_FLOW_THREADS;
constexpr int kFlowChunk = UWT_FLOW_CHUNK;  // candidate records per task
constexpr unsigned kFlowExit = 0xFFFFFFFEu;

struct FlowProblem {  // device-resident state of one problem between tasks
  DPose pose;
  float last_error;
  int lvl, k, n, nchunks, ntrace;
  unsigned done;      // chunks of the current sweep completed so far
  int phase;          // Tukey weights: 1 = histogram pass of the sweep, 0 = accumulation pass
  int chunk;          // low 16 bits: candidate records per task of the current sweep / 16;
                      // high 16 bits: residual sweeps this problem has completed so far
};
__device__ __forceinline__ int flow_pack_chunk(int records, int sweeps) {
  return (records >> 4) | (min(sweeps, 0x7FFF) << 16);
}
__device__ __forceinline__ int flow_chunk_of(int packed) { return (packed & 0xFFFF) << 4; }
__device__ __forceinline__ int flow_sweeps_of(int packed) { return packed >> 16; }
static_assert(sizeof(FlowProblem) == 64, "FlowProblem is one 64-byte record");

struct FlowCtl {
  unsigned head;    // next ticket a consumer takes
  unsigned tail;    // next ring index a producer reserves
  int active;       // problems not finished yet
  int error;        // != 0: a bounded wait expired
};

// Producer side: publish `nchunks` tasks of problem `prob` (warp-collective, after the state of
// the problem has been written and fenced).
// Release / acquire building blocks of the task protocol.  A release store or atomic is
// MEMBAR.ALL.GPU + the access; only an acquire adds CCTL.IVALL, which drops the whole SM's L1
// (the co-resident CTA's gather lines included), so acquires are kept to the places that read
// data another CTA wrote: one per pop and one per completed sweep.
__device__ __forceinline__ unsigned atom_add_release_gpu(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v)
               : "memory");
  return old;
}
__device__ __forceinline__ void fence_acq_rel_gpu() {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

__device__ __forceinline__ void flow_enqueue(FlowCtl* ctl, unsigned long long* ring, unsigned cap, int prob,
                                             int nchunks, int lane) {
  unsigned base = 0;
  if (lane == 0) base = atomicAdd(&ctl->tail, (unsigned)nchunks);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int c = lane; c < nchunks; c += 32) {
    // slot word = {generation of the ring revolution | task}: a consumer recognises ITS task by
    // the generation of its ticket, so slots are never reset and nothing can be erased
    const unsigned idx = base + (unsigned)c;
    const unsigned long long word =
        ((unsigned long long)(idx / cap + 1u) << 32) | (((unsigned)prob << 12) | (unsigned)c);
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&ring[idx % cap]), "l"(word)
                 : "memory");
  }
}

// Consumer side: take the next ticket and wait until its slot is published, or until every
// problem has finished.  A CTA takes a ticket only when it holds no task, so the holder of a
// published slot is always actively waiting for it: published-but-unconsumed slots are at most
// (outstanding tasks) <= nprob * max_chunks, waiting tickets at most one per CTA, hence a ring of
// nprob * max_chunks + gridDim.x slots can never wrap onto a live slot.
#ifdef UWT_FLOW_STATS
// debug build only: where do the CTAs wait for work?  64-us buckets since the ring was armed
__device__ unsigned long long g_flow_stats[4][64];
__device__ unsigned long long g_flow_t0;
__device__ __forceinline__ unsigned long long flow_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
void flow_debug_dump() {
  unsigned long long h[4][64];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_flow_stats, sizeof(h));
  fprintf(stderr, "bucket(64us)  pops  empty_pops  wait_us_total  exit_wait_us\n");
  for (int b = 0; b < 64; ++b)
    if (h[0][b] || h[3][b])
      fprintf(stderr, "%3d %8llu %8llu %10.1f %10.1f\n", b, h[0][b], h[1][b], h[2][b] * 1e-3,
              h[3][b] * 1e-3);
  unsigned long long z[4][64] = {};
  cudaMemcpyToSymbol(g_flow_stats, z, sizeof(z));
}
#endif

__device__ __forceinline__ unsigned flow_ticket(FlowCtl* ctl) { return atomicAdd(&ctl->head, 1u); }
__device__ __forceinline__ unsigned flow_wait_raw(FlowCtl* ctl, unsigned long long* ring,
                                                  unsigned cap, unsigned ticket, unsigned& spins) {
  const volatile unsigned long long* const slot = &ring[ticket % cap];
  const unsigned want = ticket / cap + 1u;  // generation of this ticket's ring revolution
  unsigned long long w;
  spins = 0;
  // plain polling loads; the slot is consumed by reading it (no reset: the next revolution's
  // task carries the next generation)
  while ((unsigned)((w = *slot) >> 32) != want) {
    if (*reinterpret_cast<volatile int*>(&ctl->active) <= 0) return kFlowExit;
    __nanosleep(64);
    // bounded: a protocol error ends the kernel with an error flag instead of hanging the GPU
    if (++spins > (1u << 25)) {
      atomicExch(&ctl->error, 1);
      return kFlowExit;
    }
    if ((spins & 1023u) == 0 && *reinterpret_cast<volatile int*>(&ctl->error)) return kFlowExit;
  }
  const unsigned v = (unsigned)w;
  fence_acq_rel_gpu();  // acquire: the problem state written before the publish is visible
  return v;
}
__device__ __forceinline__ unsigned flow_wait(FlowCtl* ctl, unsigned long long* ring, unsigned cap,
                                              unsigned ticket) {
  unsigned spins;
#ifdef UWT_FLOW_STATS
  const unsigned long long t0 = flow_globaltimer();
  const unsigned v = flow_wait_raw(ctl, ring, cap, ticket, spins);
  const unsigned long long t1 = flow_globaltimer();
  const unsigned b = min(63u, (unsigned)((t1 - g_flow_t0) >> 16));
  if (v == kFlowExit) {
    atomicAdd(&g_flow_stats[3][b], t1 - t0);
  } else {
    atomicAdd(&g_flow_stats[0][b], 1ull);
    if (spins) atomicAdd(&g_flow_stats[1][b], 1ull);
    atomicAdd(&g_flow_stats[2][b], t1 - t0);
  }
  return v;
#else
  return flow_wait_raw(ctl, ring, cap, ticket, spins);
#endif
}
__device__ __forceinline__ unsigned flow_pop(FlowCtl* ctl, unsigned long long* ring, unsigned cap) {
  return flow_wait(ctl, ring, cap, flow_ticket(ctl));
}

struct FlowShared {
  double warp_part[kFlowThreads / 32][kNQ];
  double tot[kNQ];
  unsigned task;
};

// Enters level fp.lvl: candidate count, chunk count, fresh iteration state (Tracker.cpp:389-393).
// A level without points is one empty evaluation that breaks (ARITHMETIC.md U2) -- run through
// gn_update on zero sums so that stats and trace equal the cluster kernel's -- followed by the
// level transition; the walk continues downwards.  Returns true when no level is left.
// Warp-collective; `zero_tot` is a warp-private scratch of kNQ doubles.
// Records per task of a sweep over n points.  A function of the problem and the launch shape only
// (never of the queue state), so the partition of a sweep -- and with it the order of the fp64
// partial sums -- is the same in every run.  A level whose sweeps cannot occupy the grid (the
// coarse levels, small batches) is cut finer than kFlowChunk; measured with the wait statistics
// of the UWT_FLOW_STATS build, this halves the idle time of the first ~130 us of a 128-problem
// launch.  (Finer chunks for the late sweeps of a level, meant to shorten the tail of the
// launch, cost more in per-task overhead than they gained: 0.98 vs 0.91 ms.)
#ifndef UWT_FLOW_MIN_CHUNK
#define UWT_FLOW_MIN_CHUNK 1024
#endif
constexpr int kFlowMinChunk = UWT_FLOW_MIN_CHUNK;
// Late sweeps: a problem that has already run kFlowLateSweeps sweeps (the batch's mean is ~11.5
// on the headline workload) is one of the stragglers the launch ends with -- by then most CTAs
// have nothing to do, and the latency of a sweep is the time of ONE task.  Its sweeps are cut
// kFlowLateShift times finer.  The rule reads the problem's own history only, so it is as
// deterministic as the rest of the partition.  (256 problems, ms per call for a threshold of
// 12 / 13 / 14 / 15 sweeps: 1.368 / 1.360 / 1.374 / 1.400; 8x instead of 4x finer: no better.)
#ifndef UWT_FLOW_LATE_SWEEPS
#define UWT_FLOW_LATE_SWEEPS 13
#endif
#ifndef UWT_FLOW_LATE_SHIFT
#define UWT_FLOW_LATE_SHIFT 2
#endif
constexpr int kFlowLateSweeps = UWT_FLOW_LATE_SWEEPS, kFlowLateShift = UWT_FLOW_LATE_SHIFT;
// Base size: kFlowChunk records; twice that from kFlowWideBatch problems on, where the queue is
// deep enough that the per-task costs (hand-over, reduction, the barrier skew of a CTA's warps)
// weigh more than the granularity at the edges of the launch (measured, ms per estimate call,
// 8192 / 16384 records: 128 problems 0.786 / 0.801, 192: 1.126 / 1.073, 256: 1.441 / 1.378,
// 512: 2.796 / 2.607).
#ifndef UWT_FLOW_WIDE_BATCH
#define UWT_FLOW_WIDE_BATCH 192
#endif
constexpr int kFlowWideBatch = UWT_FLOW_WIDE_BATCH;
__host__ __device__ inline int flow_chunk_records(int nprob, int grid, int n, int sweeps_done) {
  int c = nprob >= kFlowWideBatch ? 2 * kFlowChunk : kFlowChunk;
  while (c > kFlowMinChunk && (long long)nprob * ((n + c - 1) / c) < (long long)grid) c >>= 1;
  if (sweeps_done >= kFlowLateSweeps) {
    c >>= kFlowLateShift;
    if (c < kFlowMinChunk) c = kFlowMinChunk;
  }
  return c;
}

__device__ bool flow_enter_level(const Geom& geom, const Pools& pools, const EstimateIO& io,
                                 int prob, FlowProblem& fp, double* zero_tot, int lane,
                                 int nprob) {
  const int prev_slot = io.prev_slots[prob];
  for (;;) {
    if (fp.lvl < geom.last_level) return true;
    fp.k = 0;
    fp.last_error = 50000.0f;  // Tracker.cpp:393
    fp.n = (int)pools.ncand[(size_t)prev_slot * kMaxLevels + fp.lvl];
    {
      const int sweeps = flow_sweeps_of(fp.chunk);
      const int csz = flow_chunk_records(nprob, (int)gridDim.x, fp.n, sweeps);
      fp.chunk = flow_pack_chunk(csz, sweeps);
      fp.nchunks = (fp.n + csz - 1) / csz;
    }
    if (lane == 0 && io.stats) io.stats[prob].n_points[fp.lvl] = fp.n;
    if (fp.n > 0) return false;
    zero_tot[lane] = 0.0;
    __syncwarp();
    uwt_iter_trace* tr = (io.trace && fp.ntrace < io.trace_cap)
                             ? &io.trace[(size_t)prob * io.trace_cap + fp.ntrace]
                             : nullptr;
    gn_update(geom, zero_tot, fp.lvl, 0, fp.pose, fp.last_error,
              io.stats ? &io.stats[prob] : nullptr, tr, lane);
    if (tr) fp.ntrace += 1;
    if (fp.lvl != 0) fp.pose = se3_scale_level(fp.pose);  // Tracker.cpp:580-590
    fp.lvl -= 1;
    __syncwarp();  // every lane has read zero_tot before the next empty level rewrites it
  }
}

// After one sweep's update: next iteration of the level, or the level transition.
__device__ bool flow_advance(const Geom& geom, const Pools& pools, const EstimateIO& io, int prob,
                             FlowProblem& fp, bool brk, double* zero_tot, int lane,
                             int nprob) {
  if (!brk) {
    fp.k += 1;
    // same level, next sweep: the partition may change once the problem counts as a straggler
    const int sweeps = flow_sweeps_of(fp.chunk);
    const int csz = flow_chunk_records(nprob, (int)gridDim.x, fp.n, sweeps);
    fp.chunk = flow_pack_chunk(csz, sweeps);
    fp.nchunks = (fp.n + csz - 1) / csz;
    return false;
  }
  if (fp.lvl != 0) fp.pose = se3_scale_level(fp.pose);  // Tracker.cpp:580-590
  fp.lvl -= 1;
  return flow_enter_level(geom, pools, io, prob, fp, zero_tot, lane, nprob);
}

// Publishes the new state of a problem: either its final pose, or its next sweep's tasks.
__device__ __forceinline__ void flow_commit(const EstimateIO& io, int prob, const FlowProblem& fp,
                                            bool finished, FlowCtl* ctl, unsigned long long* ring,
                                            unsigned cap, FlowProblem* probs, int lane) {
  if (finished) {
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) io.out_poses[prob * 7 + i] = fp.pose.q[i];
      for (int i = 0; i < 3; ++i) io.out_poses[prob * 7 + 4 + i] = fp.pose.t[i];
      if (io.trace_count) io.trace_count[prob] = fp.ntrace;
      __threadfence();
      atomicSub(&ctl->active, 1);
    }
  } else {
    // release: the state written by lane 0 is ordered (warp barrier) before the release stores
    // of the task words, one per publishing lane
    if (lane == 0) probs[prob] = fp;
    __syncwarp();
    flow_enqueue(ctl, ring, cap, prob, fp.nchunks, lane);
  }
}

__global__ void flow_init_kernel(FlowCtl* ctl, unsigned long long* ring, unsigned cap, int nprob) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    ctl->head = 0u;
    ctl->tail = 0u;
    ctl->active = nprob;
    ctl->error = 0;
#ifdef UWT_FLOW_STATS
    g_flow_t0 = flow_globaltimer();
#endif
  }
  for (unsigned j = i; j < cap; j += gridDim.x * blockDim.x) ring[j] = 0ull;  // generation 0
}

#ifndef UWT_FLOW_MIN_BLOCKS
#define UWT_FLOW_MIN_BLOCKS (512 / UWT_FLOW_THREADS)
#endif

// Robust weights (kWeighted) on the dataflow kernel.  Huber weights are a fixed function of the
// integer residual: one table per CTA, built at kernel start.  Tukey weights depend on the
// median / MAD of the sweep's residuals (see the note above RobustShared), so a Tukey sweep is
// two rounds of chunk tasks: phase 1 adds the chunk's residual histogram into the problem's
// 511-bin histogram in global memory; the CTA that completes it derives median, MAD and the
// three weight tables, stores them for the problem and publishes the phase-0 (accumulation)
// tasks, which load the tables into shared memory.
struct FlowRobust {
  unsigned hist[512];
  unsigned dev[256];
  float lut_s[512], lut_rs[512], lut_e[512];  // contiguous: loaded as one [3][512] block
};

// kMode: 0 = mono input, nearest sampling (the reference; the fast point loop), 1 = per-point
// depth (cfg.depth_mode; the rigid transform is 12 doubles instead of the separable tables and a
// record's integer depth comes from `recz`), 2 = bilinear sampling (north-star option; the
// residual is a float, sum r^2 travels in fp64).  Modes 1 and 2 take identity weights only.
constexpr int kFlowMono = 0, kFlowDepth = 1, kFlowBilinear = 2;
template <bool kWeighted, int kTab, int kMode = kFlowMono>
__global__ void __launch_bounds__(kFlowThreads, UWT_FLOW_MIN_BLOCKS)
estimate_flow_kernel(const __grid_constant__ Geom geom, const Pools pools, const EstimateIO io,
                     int nprob, FlowCtl* ctl, unsigned long long* ring, unsigned cap, FlowProblem* probs,
                     double* partials, int max_chunks, unsigned* robust_hist, float* robust_lut) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FlowShared& sh = *reinterpret_cast<FlowShared*>(smem_raw);
  constexpr int table_w = kTab, table_h = kTab;  // row stride of the transform tables (entries)
  double* const tab_x = reinterpret_cast<double*>(smem_raw + sizeof(FlowShared));  // [3][kTab]
  double* const tab_y = tab_x + 3 * table_w;                                       // [3][kTab]
  FlowRobust& fr = *reinterpret_cast<FlowRobust*>(tab_y + 3 * table_h);  // kWeighted only
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float rscale = geom.residual_scale;
  const bool rscale_is_int = (rscale == truncf(rscale)) && fabsf(rscale) <= 32768.0f;
  const int rscale_i = rscale_is_int ? (int)rscale : 0;
  const bool tukey = kWeighted && geom.weight_mode == UWT_WEIGHT_TUKEY;
  WeightLut lut = {};
  if constexpr (kWeighted) {
    lut.s = fr.lut_s;
    lut.rs = fr.lut_rs;
    lut.e = fr.lut_e;
    if (!tukey) {
      // Huber (ARITHMETIC.md R4), as in the cluster kernel
      for (int i = tid; i < 512; i += kFlowThreads) {
        const float r = (float)(i - 255);
        const float a = fabsf(r);
        const float w = (a <= geom.huber_delta) ? 1.0f : __fdiv_rn(geom.huber_delta, a);
        const float sq = __fsqrt_rn(w);
        fr.lut_s[i] = sq;
        fr.lut_rs[i] = __fmul_rn(__fmul_rn(r, rscale), sq);
        fr.lut_e[i] = __fmul_rn(r, w);
      }
    }
  }

  // ---- prologue: initialise the problems and publish their first sweeps ----
  if (wid == 0) {
    for (int prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
      FlowProblem fp = {};
      if (io.init_poses) {
        for (int i = 0; i < 4; ++i) fp.pose.q[i] = io.init_poses[prob * 7 + i];
        for (int i = 0; i < 3; ++i) fp.pose.t[i] = io.init_poses[prob * 7 + 4 + i];
      } else {
        const float zero6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        fp.pose = se3_exp(zero6);  // Tracker.cpp:385
      }
      if (lane == 0 && io.stats) {
        uwt_track_stats z = {};
        io.stats[prob] = z;
      }
      __syncwarp();
      fp.lvl = geom.first_level;
      fp.phase = tukey ? 1 : 0;
      const bool finished = flow_enter_level(geom, pools, io, prob, fp, sh.tot, lane, nprob);
      flow_commit(io, prob, fp, finished, ctl, ring, cap, probs, lane);
    }
  }
  if (tid == 0) sh.task = flow_pop(ctl, ring, cap);
  __syncthreads();

  // ---- task loop ----
  for (;;) {
    const unsigned task = sh.task;
    if (task == kFlowExit) break;
    const int prob = (int)(task >> 12), chunk = (int)(task & 0xFFFu);
    const FlowProblem* P = &probs[prob];
    DPose pose;
#pragma unroll
    for (int i = 0; i < 4; ++i) pose.q[i] = __ldcg(&P->pose.q[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) pose.t[i] = __ldcg(&P->pose.t[i]);
    const int lvl = __ldcg(&P->lvl), n = __ldcg(&P->n), nchunks = __ldcg(&P->nchunks);
    const int prev_slot = io.prev_slots[prob], cur_slot = io.cur_slots[prob];
    const LevelGeom& L = geom.lv[lvl];
    const uint64_t* __restrict__ recs = pools.rec + (size_t)prev_slot * geom.rec_elems + L.rec_off;
    const uint8_t* __restrict__ I2 = pools.img + (size_t)cur_slot * geom.plane_elems + L.plane_off;
    WarpConst wc;
    wc.fx = L.fx; wc.fy = L.fy; wc.cx = L.cx; wc.cy = L.cy;
    wc.cols = L.w; wc.rows = L.h; wc.pitch = L.pitch;
    wc.colsf = (float)L.w; wc.rowsf = (float)L.h;
    wc.invfx = L.invfx; wc.invfy = L.invfy;
    // Tracker.cpp:1316,1344: factor 0.0002; ObtainAllPoints divides it by 2^level (:1266)
    wc.zfactor = geom.depth_mode == UWT_DEPTH_ALL_POINTS ? ldexpf(0.0002f, -lvl) : 0.0002f;
    const uint16_t* __restrict__ recz =
        kMode == kFlowDepth ? pools.recz + (size_t)prev_slot * geom.rec_elems + L.rec_off
                            : nullptr;
    const int chunk_word = __ldcg(&P->chunk);
    const int csz = flow_chunk_of(chunk_word);
    const int lo = chunk * csz, hi = min(n, lo + csz);
    // The chunk's records stream from DRAM (the batch's working set exceeds the L2): ask for all
    // of them now, one 128-byte line per request, so the point loop finds them in the L2; the
    // first two records of this thread's stride walk are loaded before the table build.
    {
      const char* base = reinterpret_cast<const char*>(recs + lo);
      const int bytes = (hi - lo) * (int)sizeof(uint64_t);
      for (int off = tid * 128; off < bytes; off += kFlowThreads * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    }
    const int left0 = hi - lo - tid;
    const uint64_t rec0 = (left0 > 0) ? __ldg(&recs[lo + tid]) : 0ull;
    const uint64_t rec1 = (left0 > kFlowThreads) ? __ldg(&recs[lo + tid + kFlowThreads]) : rec0;
    // tables over ALL columns of the level: no dependent read of the chunk's first / last record
    // (x-major order would allow a narrower table) on the critical path of the hand-over
    if constexpr (kMode == kFlowDepth) {
      // per-sweep rigid transform as 12 doubles T[r][0..3] (Tracker.cpp:1423-1425)
      if (tid < 12) {
        float R[9];
        quat_to_R(pose.q, R);
        const int r = tid >> 2, c = tid & 3;
        tab_x[tid] = (double)(c < 3 ? R[r * 3 + c] : pose.t[r]);
      }
    } else {
      build_tables_range(pose, L, tab_x, table_w, 0, L.w - 1, tab_y, table_h, tid, kFlowThreads);
    }
    if constexpr (kWeighted) {
      if (tukey) {
        const int phase = __ldcg(&P->phase);
        if (phase == 1) {
          // ---- histogram pass of a Tukey sweep (Tracker.cpp:496) ----
          unsigned* const gh = robust_hist + (size_t)prob * 512;
          for (int i = tid; i < 512; i += kFlowThreads) fr.hist[i] = 0u;
          __syncthreads();
          for (int i = lo + tid; i < hi; i += kFlowThreads) {
            PointGeom pg;
            int i1;
            const uint8_t* target;
            if (point_geometry<false>(wc, __ldg(&recs[i]), tab_x, table_w, tab_y, table_h,
                                      I2, pg, i1, target))
              atomicAdd(&fr.hist[(int)__ldg(target) - i1 + 255], 1u);
          }
          __syncthreads();
          for (int i = tid; i < 511; i += kFlowThreads) {
            const unsigned v = fr.hist[i];
            if (v) atomicAdd(&gh[i], v);
          }
          __syncthreads();  // every thread's additions precede the release below
          if (tid == 32) sh.task = flow_pop(ctl, ring, cap);
          if (wid == 0) {
            int last = 0;
            if (lane == 0)
              last = (atom_add_release_gpu(&probs[prob].done, 1u) == (unsigned)nchunks - 1u) ? 1 : 0;
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
              fence_acq_rel_gpu();
              // totals of the sweep; the global histogram is cleared for the next sweep
              for (int i = lane; i < 512; i += 32) {
                fr.hist[i] = (i < 511) ? __ldcg(&gh[i]) : 0u;
                __stcg(&gh[i], 0u);
              }
              __syncwarp();
              // MedianMat(Residuals): negatives saturate to 0 (Tracker.cpp:1572-1573)
              unsigned neg = 0, all = 0;
              for (int i = lane; i < 511; i += 32) {
                const unsigned v = fr.hist[i];
                all += v;
                if (i <= 255) neg += v;
              }
              neg = __reduce_add_sync(0xffffffffu, neg);
              all = __reduce_add_sync(0xffffffffu, all);
              for (int i = lane; i < 256; i += 32) fr.dev[i] = (i == 0) ? neg : fr.hist[255 + i];
              __syncwarp();
              const int med = median_from_hist256(fr.dev, all, lane);
              __syncwarp();
              // histogram of |Residuals - median|, saturated at 255 (Tracker.cpp:1613-1616)
              for (int j = lane; j < 256; j += 32) {
                unsigned v = 0;
                if (j < 255) {
                  const int hi_i = med + j + 255, lo_i = med - j + 255;
                  if (hi_i <= 510) v += fr.hist[hi_i];
                  if (j > 0 && lo_i >= 0) v += fr.hist[lo_i];
                } else {
                  for (int r = -255; r <= 255; ++r)
                    if (abs(r - med) >= 255) v += fr.hist[r + 255];
                }
                fr.dev[j] = v;
              }
              __syncwarp();
              const int mad_bin = median_from_hist256(fr.dev, all, lane);
              // TukeyFunctionWeights (Tracker.cpp:1626-1651) as tables over r
              float MAD = __fmul_rn(1.4826f, (float)mad_bin);  // Tracker.cpp:1608,1618
              if (MAD == 0.0f) MAD = 1.0f;                     // Tracker.cpp:1634-1637
              const float inv_MAD = (float)(1.0 / (double)MAD);
              const float inv_b2 = (float)(1.0 / (double)__fmul_rn(4.6851f, 4.6851f));
              float* const gl = robust_lut + (size_t)prob * 1536;
              for (int i = lane; i < 512; i += 32) {
                const float r = (float)(i - 255);
                const float w = (i < 511) ? tukey_weight(r, inv_MAD, inv_b2) : 0.0f;
                __stcg(&gl[i], w);
                __stcg(&gl[512 + i], __fmul_rn(__fmul_rn(r, rscale), w));
                __stcg(&gl[1024 + i], __fmul_rn(r, w));
              }
              if (lane == 0) {
                *reinterpret_cast<volatile int*>(&probs[prob].phase) = 0;
                *reinterpret_cast<volatile unsigned*>(&probs[prob].done) = 0u;
              }
              __syncwarp();
              flow_enqueue(ctl, ring, cap, prob, nchunks, lane);  // release stores
            }
          }
          __syncthreads();
          continue;
        }
        // ---- accumulation pass: this sweep's weight tables ----
        const float* const gl = robust_lut + (size_t)prob * 1536;
        for (int i = tid; i < 1536; i += kFlowThreads) fr.lut_s[i] = __ldcg(&gl[i]);
      }
    }
    __syncthreads();
    double acc[kNQ];
#pragma unroll
    for (int i = 0; i < kNQ; ++i) acc[i] = 0.0;
    unsigned sum_r2 = 0, n_val = 0;
    {
      const uint32_t tabx = (uint32_t)__cvta_generic_to_shared(tab_x);
      const uint32_t taby = (uint32_t)__cvta_generic_to_shared(tab_y);
      // the fast loop assumes the reference's integer residual scale and principal points away
      // from 0 (Geom::exact_div); anything else, levels beyond 4, depth input and bilinear
      // sampling run the generic loop
      const bool fast = (kMode == kFlowMono && fast_sweep_applies(geom, lvl)) ||
                        (kMode == kFlowDepth && fast_depth_sweep_applies(geom, lvl));
      switch (fast ? 1 : 0) {  // CTA-uniform
        case 1:
          if constexpr (kMode == kFlowDepth)
            fast_depth_sweep(geom, lvl, recs, recz, lo + tid, hi, kFlowThreads, tabx, tab_x, I2,
                             rscale, acc, sum_r2, n_val);
          else
            fast_sweep<kWeighted, kTab>(geom, lvl, recs, lo + tid, hi, kFlowThreads, rec0, rec1,
                                        tabx, taby, tab_x, tab_y, I2, rscale, acc, sum_r2, n_val,
                                        lut);
          break;
        default: {
          int i = lo + tid;
          uint64_t rec = (i < hi) ? __ldg(&recs[i]) : 0ull;
          while (i < hi) {
            const int inext = i + kFlowThreads;
            const uint64_t rec_next = (inext < hi) ? __ldg(&recs[inext]) : 0ull;
            if constexpr (kMode == kFlowBilinear)
              accumulate_point_bilinear(wc, rec, tab_x, table_w, tab_y, table_h, I2, rscale, acc,
                                        n_val);
            else
              accumulate_point<kWeighted, kMode == kFlowDepth>(
                  wc, rec, tab_x, table_w, tab_y, table_h, I2, rscale, rscale_is_int,
                  rscale_i, acc, sum_r2, n_val, lut,
                  kMode == kFlowDepth ? (int)__ldg(&recz[i]) : 0);
            rec = rec_next;
            i = inext;
          }
        }
      }
    }
    acc[27] = (double)sum_r2;
    acc[28] = (double)n_val;
    // the ticket of the next task is requested now (its atomic's round trip overlaps the warp
    // reduction); the wait for the slot happens after the barrier
    unsigned next_ticket = 0u;
    if (tid == 32) next_ticket = flow_ticket(ctl);
    const double wtot = warp_reduce32(acc, lane);
    sh.warp_part[wid][lane] = wtot;
    __syncthreads();  // warp_part complete; every thread has read sh.task
    // The next task is fetched by warp 1 while warp 0 does this chunk's bookkeeping: the two
    // latency chains (ticket + slot + fence; partial store + fence + counter) run side by side.
    // Warp 0 never waits for warp 1 here, so a pop that has to wait for work -- possibly the
    // work warp 0 is about to publish -- cannot block it.
    if (tid == 32) sh.task = flow_wait(ctl, ring, cap, next_ticket);
    if (wid == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kFlowThreads / 32; ++w) s += sh.warp_part[w][lane];
      double* part = partials + ((size_t)prob * max_chunks + chunk) * kNQ;
      __stcg(&part[lane], s);
      __syncwarp();
      // release by the lane that counts the chunk, cumulative over the warp's partial stores
      int last = 0;
      if (lane == 0)
        last = (atom_add_release_gpu(&probs[prob].done, 1u) == (unsigned)nchunks - 1u) ? 1 : 0;
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        // ---- this CTA completed the sweep: reduce in chunk order, update, schedule next ----
        // acquire: the other chunks' partials (and the previous update's stats / trace rows)
        fence_acq_rel_gpu();
        const double* pp = partials + (size_t)prob * max_chunks * kNQ;
        double tsum = 0.0;
        for (int c = 0; c < nchunks; ++c) tsum += __ldcg(&pp[(size_t)c * kNQ + lane]);
        sh.tot[lane] = tsum;
        __syncwarp();
        FlowProblem fp;
        fp.pose = pose;
        fp.last_error = __ldcg(&P->last_error);
        fp.lvl = lvl;
        fp.k = __ldcg(&P->k);
        fp.n = n;
        fp.nchunks = nchunks;
        fp.ntrace = __ldcg(&P->ntrace);
        fp.done = 0;
        fp.phase = tukey ? 1 : 0;
        fp.chunk = flow_pack_chunk(csz, flow_sweeps_of(chunk_word) + 1);  // this sweep is done
        uwt_iter_trace* tr = (io.trace && fp.ntrace < io.trace_cap)
                                 ? &io.trace[(size_t)prob * io.trace_cap + fp.ntrace]
                                 : nullptr;
        const bool brk = gn_update(geom, sh.tot, lvl, fp.k, fp.pose, fp.last_error,
                                   io.stats ? &io.stats[prob] : nullptr, tr, lane);
        if (tr) fp.ntrace += 1;
        __syncwarp();
        const bool finished = flow_advance(geom, pools, io, prob, fp, brk, sh.tot, lane, nprob);
          flow_commit(io, prob, fp, finished, ctl, ring, cap, probs, lane);
      }
    }
    __syncthreads();  // next task published; tables and warp_part are free for reuse
  }
}

int flow_max_chunks(const Geom& g, int chunk_records) {
  long long m = 1;
  for (int l = g.last_level; l <= g.first_level; ++l)
    m = std::max(m, ((long long)g.lv[l].w * g.lv[l].h + chunk_records - 1) / chunk_records);
  return (int)m;
}


static size_t round256(size_t v) { return (v + 255) / 256 * 256; }

constexpr unsigned kFlowRingSlack = 148 * 8 + 64;  // >= CTAs of the persistent grid

size_t flow_workspace_bytes(const Geom& g, int nprob) {
  const size_t mc = (size_t)flow_max_chunks(g, kFlowMinChunk);
  return 256 + round256(((size_t)nprob * mc + kFlowRingSlack) * sizeof(unsigned long long)) +
         round256((size_t)nprob * sizeof(FlowProblem)) +
         round256((size_t)nprob * mc * kNQ * sizeof(double)) +
         (g.weight_mode == UWT_WEIGHT_TUKEY ? (size_t)nprob * (512 + 1536) * 4 : 0);
}

// workspace layout: [FlowCtl | ring | FlowProblem[] | partials | Tukey histograms | Tukey tables];
// the control block, the ring and the histograms are re-initialised on the stream before every
// launch.
template <bool kWeighted, int kTab, int kMode>
static int launch_estimate_flow_t(const Geom& g, const Pools& p, int n, const EstimateIO& io,
                                  void* workspace, cudaStream_t st, int* grid_cache) {
  const int mc = flow_max_chunks(g, kFlowMinChunk);
  if (mc > 4095 || n >= (1 << 20)) return -2;  // task word: 12-bit chunk, 20-bit problem
  const unsigned cap = (unsigned)((size_t)n * mc) + kFlowRingSlack;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  FlowCtl* ctl = reinterpret_cast<FlowCtl*>(w);
  unsigned long long* ring = reinterpret_cast<unsigned long long*>(w + 256);
  size_t off = 256 + round256((size_t)cap * sizeof(unsigned long long));
  FlowProblem* probs = reinterpret_cast<FlowProblem*>(w + off);
  off += round256((size_t)n * sizeof(FlowProblem));
  double* partials = reinterpret_cast<double*>(w + off);
  off += round256((size_t)n * mc * kNQ * sizeof(double));
  unsigned* robust_hist = reinterpret_cast<unsigned*>(w + off);  // [n][512]   (Tukey)
  off += (size_t)n * 512 * sizeof(unsigned);
  float* robust_lut = reinterpret_cast<float*>(w + off);         // [n][3][512] (Tukey)
  const bool tukey = kWeighted && g.weight_mode == UWT_WEIGHT_TUKEY;
  if (tukey &&
      cudaMemsetAsync(robust_hist, 0, (size_t)n * 512 * sizeof(unsigned), st) != cudaSuccess)
    return -1;
  const size_t smem = sizeof(FlowShared) + sizeof(double) * 6 * (size_t)kTab +
                      (kWeighted ? sizeof(FlowRobust) : 0);
  static size_t smem_cache[kMaxDevices];  // one per kernel instantiation and device
  if (!ensure_dynamic_smem(estimate_flow_kernel<kWeighted, kTab, kMode>, smem, smem_cache)) return -1;
  // Persistent grid = co-resident CTAs for THIS handle's shared-memory size, computed once per
  // handle (the caller owns `grid_cache`): the chunk partition of a sweep depends on the grid
  // (flow_chunk_records), so it must not depend on which other handles ran before.
  if (*grid_cache <= 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, estimate_flow_kernel<kWeighted, kTab, kMode>,
                                                  kFlowThreads, smem);
    if (sms <= 0) sms = 148;
    if (per_sm <= 0) per_sm = 1;
    *grid_cache = std::min(sms * per_sm, (int)kFlowRingSlack - 64);
  }
  flow_init_kernel<<<std::max(1u, std::min(cap / 256u + 1u, 296u)), 256, 0, st>>>(ctl, ring, cap, n);
  if (cudaGetLastError() != cudaSuccess) return -1;
  const int grid = *grid_cache;
  estimate_flow_kernel<kWeighted, kTab, kMode><<<grid, kFlowThreads, smem, st>>>(
      g, p, io, n, ctl, ring, cap, probs, partials, mc, robust_hist, robust_lut);
  return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

int launch_estimate_flow(const Geom& g, const Pools& p, int n, const EstimateIO& io,
                         void* workspace, cudaStream_t st, int* grid_cache) {
  // row stride of the transform tables: the smallest instantiated size that holds the finest
  // optimised level (larger levels: the caller falls back to the cluster kernel)
  const int dim = std::max(g.lv[g.last_level].w, g.lv[g.last_level].h);
  const bool ident = g.weight_mode == UWT_WEIGHT_IDENTITY;
  const int mode = g.depth_mode != UWT_DEPTH_NONE ? kFlowDepth
                   : g.sampling == UWT_SAMPLE_BILINEAR ? kFlowBilinear : kFlowMono;
  if (mode != kFlowMono && !ident) return -2;  // (uwt_create rejects these combinations)
#define UWT_FLOW_LAUNCH(W, T, M) \
  launch_estimate_flow_t<W, T, M>(g, p, n, io, workspace, st, grid_cache)
  if (dim <= 1024) {
    if (mode == kFlowDepth) return UWT_FLOW_LAUNCH(false, 1024, kFlowDepth);
    if (mode == kFlowBilinear) return UWT_FLOW_LAUNCH(false, 1024, kFlowBilinear);
    return ident ? UWT_FLOW_LAUNCH(false, 1024, kFlowMono) : UWT_FLOW_LAUNCH(true, 1024, kFlowMono);
  }
  if (dim <= 2048) {
    if (mode == kFlowDepth) return UWT_FLOW_LAUNCH(false, 2048, kFlowDepth);
    if (mode == kFlowBilinear) return UWT_FLOW_LAUNCH(false, 2048, kFlowBilinear);
    return ident ? UWT_FLOW_LAUNCH(false, 2048, kFlowMono) : UWT_FLOW_LAUNCH(true, 2048, kFlowMono);
  }
#undef UWT_FLOW_LAUNCH
  return -2;
}

}  // namespace uwt
