// uwt_estimate_shard.cu -- ONE tracking problem on the whole GPU or split over GPUs: the
// accumulate / update pair of the NCCL host loop and the persistent fused kernel with its
// in-kernel all-reduce over peer-mapped mailboxes (Tracker.cpp:414-577 split by candidate range).
#include "uwt_estimate_common.cuh"

namespace uwt {

// ----------------------------------------------------------------------------------------
// Sharded single-frame mode (BASELINE config 4): the candidate list of ONE
// tracking problem is split into `nranks` contiguous ranges, one per GPU.  Per Gauss-Newton
// sweep every rank runs shard_accumulate_kernel over its range, the caller all-reduces the 32
// fp64 partial sums across ranks (NCCL over NVLink), and every rank runs shard_update_kernel
// redundantly on the identical totals, so all ranks hold bit-identical poses without a
// broadcast.  State lives on the device between calls.
// ----------------------------------------------------------------------------------------
constexpr int kShardThreads = 256;

// Huber weights (UWT_WEIGHT_HUBER, ARITHMETIC.md R4) in the sharded mode: the weight of a point is
// a fixed function of its integer residual, so every rank builds the same three tables and the
// exchange stays the 32 sums of the identity case (sum 29 carries the weighted error term).
// Tukey / MAD weights would need the sweep's residual histogram all-reduced first: not offered.
template <bool kWeighted>
struct ShardLut {
  float s[kWeighted ? 512 : 1], rs[kWeighted ? 512 : 1], e[kWeighted ? 512 : 1];
};
template <bool kWeighted>
__device__ __forceinline__ WeightLut shard_build_lut(ShardLut<kWeighted>& t, const Geom& geom,
                                                     float rscale, int tid) {
  WeightLut lut = {};
  if constexpr (kWeighted) {
    for (int i = tid; i < 512; i += kShardThreads) {
      const float r = (float)(i - 255);
      const float a = fabsf(r);
      const float w = (a <= geom.huber_delta) ? 1.0f : __fdiv_rn(geom.huber_delta, a);
      const float sq = __fsqrt_rn(w);
      t.s[i] = sq;
      t.rs[i] = __fmul_rn(__fmul_rn(r, rscale), sq);
      t.e[i] = __fmul_rn(r, w);
    }
    lut.s = t.s;
    lut.rs = t.rs;
    lut.e = t.e;
  }
  return lut;
}

template <bool kWeighted>
__global__ void __launch_bounds__(kShardThreads, 2)
shard_accumulate_kernel(const __grid_constant__ Geom geom, const Pools pools, ShardState* st,
                        double* __restrict__ partials, double* __restrict__ out32, int table_w,
                        int table_h) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* const tab_x = reinterpret_cast<double*>(smem_raw);
  double* const tab_y = tab_x + 3 * table_w;
  __shared__ double warp_part[kShardThreads / 32][kNQ];
  __shared__ int is_last;
  __shared__ ShardLut<kWeighted> lut_mem;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int lvl = st->level;
  DPose pose;
  for (int i = 0; i < 4; ++i) pose.q[i] = st->pose[i];
  for (int i = 0; i < 3; ++i) pose.t[i] = st->pose[4 + i];
  const LevelGeom& L = geom.lv[lvl];
  const long long n = (long long)pools.ncand[(size_t)st->prev_slot * kMaxLevels + lvl];
  const int lo = (int)(n * st->rank / st->nranks), hi = (int)(n * (st->rank + 1) / st->nranks);
  const uint64_t* __restrict__ recs =
      pools.rec + (size_t)st->prev_slot * geom.rec_elems + L.rec_off;
  const uint8_t* __restrict__ I2 =
      pools.img + (size_t)st->cur_slot * geom.plane_elems + L.plane_off;
  WarpConst wc;
  wc.fx = L.fx; wc.fy = L.fy; wc.cx = L.cx; wc.cy = L.cy;
  wc.cols = L.w; wc.rows = L.h; wc.pitch = L.pitch;
  wc.colsf = (float)L.w; wc.rowsf = (float)L.h;
  const float rscale = geom.residual_scale;
  const bool rscale_is_int = (rscale == truncf(rscale)) && fabsf(rscale) <= 32768.0f;
  const int rscale_i = rscale_is_int ? (int)rscale : 0;
  const WeightLut lut = shard_build_lut<kWeighted>(lut_mem, geom, rscale, tid);
  build_tables(pose, L, tab_x, table_w, tab_y, table_h, tid, kShardThreads);
  __syncthreads();
  double acc[kNQ];
#pragma unroll
  for (int i = 0; i < kNQ; ++i) acc[i] = 0.0;
  unsigned sum_r2 = 0, n_val = 0;
  {
    const int stride = gridDim.x * kShardThreads;
    int i = lo + blockIdx.x * kShardThreads + tid;
    uint64_t rec = (i < hi) ? __ldg(&recs[i]) : 0ull;
    while (i < hi) {
      const int inext = i + stride;
      const uint64_t rec_next = (inext < hi) ? __ldg(&recs[inext]) : 0ull;
      accumulate_point<kWeighted>(wc, rec, tab_x, table_w, tab_y, table_h, I2, rscale,
                                  rscale_is_int, rscale_i, acc, sum_r2, n_val, lut);
      rec = rec_next;
      i = inext;
    }
  }
  acc[27] = (double)sum_r2;
  acc[28] = (double)n_val;
  const double wtot = warp_reduce32(acc, lane);
  warp_part[wid][lane] = wtot;
  __syncthreads();
  if (wid == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kShardThreads / 32; ++w) s += warp_part[w][lane];
    partials[(size_t)blockIdx.x * kNQ + lane] = s;
    __threadfence();
    if (lane == 0) is_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && wid == 0) {
    __threadfence();
    double s = 0.0;  // fixed block order: deterministic
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[(size_t)b * kNQ + lane]);
    out32[lane] = s;
    if (lane == 0) st->ticket = 0;
  }
}

__global__ void shard_update_kernel(const __grid_constant__ Geom geom, const Pools pools,
                                    ShardState* st, const double* __restrict__ sums32,
                                    int* __restrict__ done_out) {
  __shared__ double tot[kNQ];
  const int lane = threadIdx.x;  // launched with exactly one warp
  tot[lane] = sums32[lane];
  __syncwarp();
  DPose pose;
  for (int i = 0; i < 4; ++i) pose.q[i] = st->pose[i];
  for (int i = 0; i < 3; ++i) pose.t[i] = st->pose[4 + i];
  float last_error = st->last_error;
  int lvl = st->level, k = st->k;
  int done = st->done;
  __syncwarp();
  if (lane == 0)
    st->stats.n_points[lvl] = (int)pools.ncand[(size_t)st->prev_slot * kMaxLevels + lvl];
  const bool brk = gn_update(geom, tot, lvl, k, pose, last_error, &st->stats, nullptr, lane);
  if (brk) {
    if (lvl != 0) pose = se3_scale_level(pose);  // Tracker.cpp:580-590
    --lvl;
    k = 0;
    last_error = 50000.0f;  // Tracker.cpp:393
    if (lvl < geom.last_level) done = 1;
  } else {
    ++k;
  }
  if (lane == 0) {
    for (int i = 0; i < 4; ++i) st->pose[i] = pose.q[i];
    for (int i = 0; i < 3; ++i) st->pose[4 + i] = pose.t[i];
    st->last_error = last_error;
    st->level = lvl;
    st->k = k;
    st->done = done;
    *done_out = done;
  }
}

int launch_shard_accumulate(const Geom& g, const Pools& p, ShardState* st, double* partials,
                            double* out32, int grid, cudaStream_t stream) {
  const int tw = g.lv[g.last_level].w, th = g.lv[g.last_level].h;
  const size_t smem = sizeof(double) * 3 * (size_t)(tw + th);
  static size_t smem_cache[2][kMaxDevices];  // one per kernel instantiation and device
  if (g.weight_mode == UWT_WEIGHT_HUBER) {
    if (!ensure_dynamic_smem(shard_accumulate_kernel<true>, smem, smem_cache[1])) return -1;
    shard_accumulate_kernel<true><<<grid, kShardThreads, smem, stream>>>(g, p, st, partials, out32,
                                                                         tw, th);
  } else {
    if (!ensure_dynamic_smem(shard_accumulate_kernel<false>, smem, smem_cache[0])) return -1;
    shard_accumulate_kernel<false><<<grid, kShardThreads, smem, stream>>>(g, p, st, partials,
                                                                          out32, tw, th);
  }
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_shard_update(const Geom& g, const Pools& p, ShardState* st, const double* sums32,
                        int* done_out, cudaStream_t stream) {
  shard_update_kernel<<<1, 32, 0, stream>>>(g, p, st, sums32, done_out);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ----------------------------------------------------------------------------------------
// Fused compute + collective form of the sharded mode: ONE persistent kernel per rank runs the
// whole Gauss-Newton loop.  Per sweep every CTA accumulates its part of this rank's candidate
// range; the last CTA to finish reduces the per-CTA partials, STORES the rank's 32 sums and a
// sequence flag straight into every peer's mailbox (peer-mapped memory, i.e. NVLink writes),
// waits for the peers' flags, adds the mailbox rows in rank order (identical on every rank),
// runs the update and releases the other CTAs through a generation counter.  No host round
// trip and no NCCL call per sweep: the exchange is 32 x 8 B per peer, pure latency.
// Every wait is bounded (a lost peer sets ctl->error and ends the kernel instead of hanging).
// ----------------------------------------------------------------------------------------

constexpr long long kSpinLimit = 50LL * 1000 * 1000;  // x (20 ns sleep + one load): seconds

template <bool kWeighted>
__global__ void __launch_bounds__(kShardThreads, 2)
shard_fused_kernel(const __grid_constant__ Geom geom, const Pools pools, ShardState* st,
                   ShardFused* ctl, ShardMailbox* mine, double* __restrict__ partials,
                   int table_w, int table_h, unsigned poll_ns) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* const tab_x = reinterpret_cast<double*>(smem_raw);
  double* const tab_y = tab_x + 3 * table_w;
  __shared__ double warp_part[kShardThreads / 32][kNQ];
  __shared__ double tot[kNQ];
  __shared__ int is_last;
  __shared__ int s_level, s_done;
  __shared__ float s_pose[7];
  __shared__ int s_ncand[kMaxLevels];
  __shared__ ShardLut<kWeighted> lut_mem;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int rank = st->rank, nranks = st->nranks;
  const float rscale = geom.residual_scale;
  const bool rscale_is_int = (rscale == truncf(rscale)) && fabsf(rscale) <= 32768.0f;
  const int rscale_i = rscale_is_int ? (int)rscale : 0;
  const WeightLut lut = shard_build_lut<kWeighted>(lut_mem, geom, rscale, tid);  // before the first barrier
  const unsigned gen0 = ctl->generation;             // same value in every CTA at launch
  const unsigned long long seq0 = ctl->seq;
  unsigned local_sweep = 0;
  // state of the first sweep (published by uwt_shard_begin); later sweeps receive it from the
  // leader's update through ctl->state_ll
  if (tid == 0) {
    s_level = *(volatile int*)&st->level;
    s_done = *(volatile int*)&st->done;
    for (int i = 0; i < 7; ++i) s_pose[i] = ((volatile float*)st->pose)[i];
  }
  if (tid < kMaxLevels) s_ncand[tid] = (int)pools.ncand[(size_t)st->prev_slot * kMaxLevels + tid];
  __syncthreads();

  for (;;) {
    if (s_done) break;
    const long long c0 = clock64();
    const int lvl = s_level;
    DPose pose;
    for (int i = 0; i < 4; ++i) pose.q[i] = s_pose[i];
    for (int i = 0; i < 3; ++i) pose.t[i] = s_pose[4 + i];
    const LevelGeom& L = geom.lv[lvl];
    const long long n = (long long)s_ncand[lvl];
    const int lo = (int)(n * rank / nranks), hi = (int)(n * (rank + 1) / nranks);
    // this CTA's contiguous part of the rank's range: x-major order keeps it inside a few image
    // columns, so only those columns of the x table are built (plus all rows of the y table)
    const int clo = lo + (int)((long long)(hi - lo) * blockIdx.x / gridDim.x);
    const int chi = lo + (int)((long long)(hi - lo) * (blockIdx.x + 1) / gridDim.x);
    const uint64_t* __restrict__ recs =
        pools.rec + (size_t)st->prev_slot * geom.rec_elems + L.rec_off;
    const uint8_t* __restrict__ I2 =
        pools.img + (size_t)st->cur_slot * geom.plane_elems + L.plane_off;
    WarpConst wc;
    wc.fx = L.fx; wc.fy = L.fy; wc.cx = L.cx; wc.cy = L.cy;
    wc.cols = L.w; wc.rows = L.h; wc.pitch = L.pitch;
    wc.colsf = (float)L.w; wc.rowsf = (float)L.h;
    int xlo = 0, xhi = -1;
    if (chi > clo) {
      xlo = (int)(__ldg(&recs[clo]) & 0xFFFu);
      xhi = (int)(__ldg(&recs[chi - 1]) & 0xFFFu);
    }
    build_tables_range(pose, L, tab_x, table_w, xlo, xhi, tab_y, table_h, tid, kShardThreads);
    __syncthreads();
    double acc[kNQ];
#pragma unroll
    for (int i = 0; i < kNQ; ++i) acc[i] = 0.0;
    unsigned sum_r2 = 0, n_val = 0;
    {
      const int stride = kShardThreads;
      const int first = clo + tid;
      // the fast point loop (software-pipelined gather) where a thread has enough points for its
      // prologue to pay off; its tables have a compile-time row stride (launch_shard_fused)
      const bool fast = fast_sweep_applies(geom, lvl) && (chi - clo) >= 8 * stride &&
                        table_w == table_h && (table_w == 1024 || table_w == 2048);
      if (fast) {
        const uint64_t rec0 = (first < chi) ? __ldg(&recs[first]) : 0ull;
        const uint64_t rec1 = (first + stride < chi) ? __ldg(&recs[first + stride]) : rec0;
        const uint32_t ax = (uint32_t)__cvta_generic_to_shared(tab_x) - (uint32_t)xlo * 8u;
        const uint32_t ay = (uint32_t)__cvta_generic_to_shared(tab_y);
        if (table_w == 1024)
          fast_sweep<kWeighted, 1024>(geom, lvl, recs, first, chi, stride, rec0, rec1, ax, ay,
                                      tab_x - xlo, tab_y, I2, rscale, acc, sum_r2, n_val, lut);
        else
          fast_sweep<kWeighted, 2048>(geom, lvl, recs, first, chi, stride, rec0, rec1, ax, ay,
                                      tab_x - xlo, tab_y, I2, rscale, acc, sum_r2, n_val, lut);
      } else {
        int i = first;
        uint64_t rec = (i < chi) ? __ldg(&recs[i]) : 0ull;
        while (i < chi) {
          const int inext = i + stride;
          const uint64_t rec_next = (inext < chi) ? __ldg(&recs[inext]) : 0ull;
          accumulate_point<kWeighted>(wc, rec, tab_x - xlo, table_w, tab_y, table_h, I2, rscale,
                                      rscale_is_int, rscale_i, acc, sum_r2, n_val, lut);
          rec = rec_next;
          i = inext;
        }
      }
    }
    acc[27] = (double)sum_r2;
    acc[28] = (double)n_val;
    const double wtot = warp_reduce32(acc, lane);
    warp_part[wid][lane] = wtot;
    __syncthreads();
    if (wid == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kShardThreads / 32; ++w) s += warp_part[w][lane];
      partials[(size_t)blockIdx.x * kNQ + lane] = s;
      __threadfence();
      if (lane == 0) {
        if (blockIdx.x == 0) ctl->dbg[0] += (unsigned long long)(clock64() - c0);
        atomicAdd(&st->ticket, 1u);
        // CTA 0 is always the leader (deterministic; its update code stays in one SM's
        // instruction cache)
        is_last = (blockIdx.x == 0);
        if (is_last) {
          long long spins = 0;
          while (ld_acquire_gpu(&st->ticket) < gridDim.x) {
            if (++spins > 8 * kSpinLimit) break;
          }
        }
      }
    }
    __syncthreads();
    const long long c1 = clock64();
    const unsigned long long seq = seq0 + local_sweep + 1;   // sequence number of this sweep
    const int par = (int)(seq & 1ull);
    if (is_last) {
      // all 8 warps sum a fixed, strided slice of the per-CTA partials (deterministic), then
      // warp 0 combines them: 8x shorter dependent load chain than one warp walking all CTAs
      __threadfence();
      double ps = 0.0;
      for (unsigned b = wid; b < gridDim.x; b += kShardThreads / 32)
        ps += __ldcg(&partials[(size_t)b * kNQ + lane]);
      warp_part[wid][lane] = ps;
    }
    __syncthreads();
    if (is_last && wid == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kShardThreads / 32; ++w) s += warp_part[w][lane];
      if (lane == 0) st->ticket = 0;
      const long long c2 = clock64();
      // ---- all-reduce over peer memory: push my row to every rank, then pull the sum ----
      // LL protocol: {32 data bits | 32-bit sweep number} per 8-byte store; no fences.
      const unsigned long long tag = (seq & 0xffffffffull) << 32;
      {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
        const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
        for (int r = 0; r < nranks; ++r) {
          volatile unsigned long long* dst = ctl->peer[r]->ll[par][rank];
          dst[2 * lane] = w0;
          dst[2 * lane + 1] = w1;
        }
      }
      bool ok = true;
      double t = 0.0;
      for (int r = 0; r < nranks; ++r) {  // rank order: identical sum on every rank
        const volatile unsigned long long* src = mine->ll[par][r];
        unsigned long long a, b;
        long long spins = 0;
        for (;;) {
          a = src[2 * lane];
          b = src[2 * lane + 1];
          if ((a & 0xffffffff00000000ull) == tag && (b & 0xffffffff00000000ull) == tag) break;
          if (++spins > kSpinLimit) { ok = false; break; }
        }
        t += __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
      }
      ok = __all_sync(0xffffffffu, ok);
      tot[lane] = t;
      __syncwarp();
      const long long c3 = clock64();
      // ---- K5 on the totals, warp-collective (6x6 LU one row per lane, SE3 exp with two sincos
      //      lanes): every lane ends with the same pose; then the level bookkeeping (same as
      //      shard_update_kernel) by lane 0 ----
      DPose p2 = pose;
      float last_error = *(volatile float*)&st->last_error;
      int k = *(volatile int*)&st->k, lv = lvl;
      bool brk = false;
      if (ok) {
        if (lane == 0) st->stats.n_points[lv] = (int)n;
        __syncwarp();
        brk = gn_update(geom, tot, lv, k, p2, last_error, &st->stats, nullptr, lane);
      }
      if (lane == 0) {
        if (!ok) {
          ctl->error = 1;
          st->done = 1;
        } else {
          if (brk) {
            if (lv != 0) p2 = se3_scale_level(p2);  // Tracker.cpp:580-590
            --lv;
            k = 0;
            last_error = 50000.0f;  // Tracker.cpp:393
            if (lv < geom.last_level) st->done = 1;
          } else {
            ++k;
          }
          for (int i = 0; i < 4; ++i) st->pose[i] = p2.q[i];
          for (int i = 0; i < 3; ++i) st->pose[4 + i] = p2.t[i];
          st->last_error = last_error;
          st->level = lv;
          st->k = k;
        }
        ctl->seq = seq;
        ctl->dbg[1] += (unsigned long long)(c1 - c0);
        ctl->dbg[2] += (unsigned long long)(c2 - c1);
        ctl->dbg[3] += (unsigned long long)(c3 - c2);
        ctl->dbg[4] += (unsigned long long)(clock64() - c3);
        ctl->dbg[5] += 1;
        ctl->generation = gen0 + local_sweep + 1;  // read by the next launch only
      }
      // ---- publish the next sweep's state to the waiting CTAs: 9 LL words, one per lane ----
      __syncwarp();
      if (lane < 9) {
        unsigned payload;
        if (lane < 7) payload = __float_as_uint(((volatile float*)st->pose)[lane]);
        else if (lane == 7) payload = (unsigned)*(volatile int*)&st->level;
        else payload = (unsigned)*(volatile int*)&st->done;
        const unsigned long long tagw = (unsigned long long)(gen0 + local_sweep + 1) << 32;
        *(volatile unsigned long long*)&ctl->state_ll[lane] = tagw | payload;
      }
    }
    // ---- grid barrier + state hand-over in one round trip: poll the LL words of this sweep's
    //      update (tag = generation), payload = pose, level, done ----
    if (tid < 9) {
      const unsigned want = gen0 + local_sweep + 1;
      const volatile unsigned long long* w = &ctl->state_ll[tid];
      unsigned long long v;
      long long spins = 0;
      while ((unsigned)((v = *w) >> 32) != want) {
        __nanosleep(poll_ns);
        if (++spins > 4 * kSpinLimit) {  // the leader reports the error; just leave
          v = (tid == 8) ? 1ull : 0ull;  // done = 1
          break;
        }
      }
      const unsigned payload = (unsigned)(v & 0xffffffffull);
      if (tid < 7) s_pose[tid] = __uint_as_float(payload);
      else if (tid == 7) s_level = (int)payload;
      else s_done = (int)payload;
    }
    __syncthreads();
    ++local_sweep;
    if (local_sweep > 4096u) break;  // cannot happen: levels * max_iterations is far smaller
  }
}

int launch_shard_fused(const Geom& g, const Pools& p, ShardState* st, ShardFused* ctl,
                       ShardMailbox* mine, double* partials, int grid, cudaStream_t stream) {
  int tw = g.lv[g.last_level].w, th = g.lv[g.last_level].h;
  // transform tables with the compile-time row stride of the fast sweep where the level fits
  if (std::max(tw, th) <= 1024) tw = th = 1024;
  else if (std::max(tw, th) <= 2048) tw = th = 2048;
  const size_t smem = sizeof(double) * 3 * (size_t)(tw + th);
  const bool huber = g.weight_mode == UWT_WEIGHT_HUBER;
  const void* kernel = huber ? (const void*)shard_fused_kernel<true>
                             : (const void*)shard_fused_kernel<false>;
  static size_t smem_cache[2][kMaxDevices];  // one per kernel instantiation and device
  if (huber ? !ensure_dynamic_smem(shard_fused_kernel<true>, smem, smem_cache[1])
            : !ensure_dynamic_smem(shard_fused_kernel<false>, smem, smem_cache[0]))
    return -1;
  // cooperative launch: all CTAs must be co-resident (they wait on each other); grid <= 0 asks
  // for every CTA the device can hold at once (two per SM: twice the warps to hide the point
  // loop's latency)
  if (grid <= 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (huber)
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shard_fused_kernel<true>,
                                                    kShardThreads, smem);
    else
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shard_fused_kernel<false>,
                                                    kShardThreads, smem);
    grid = std::max(1, std::min(sms * std::max(per_sm, 1), kShardMaxGrid));
  }
  static unsigned poll_ns = 0;
  if (poll_ns == 0) {
    const char* e = getenv("UWT_POLL_NS");  // tuning knob of the grid / peer wait loops
    poll_ns = e ? (unsigned)atoi(e) : 64u;
    if (poll_ns == 0) poll_ns = 1;
  }
  void* args[] = {(void*)&g, (void*)&p, (void*)&st, (void*)&ctl, (void*)&mine, (void*)&partials,
                  (void*)&tw, (void*)&th, (void*)&poll_ns};
  cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kShardThreads), args, smem,
                                              stream);
  return e == cudaSuccess ? 1 : -1;
}

}  // namespace uwt
