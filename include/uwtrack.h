/*
 * uwtrack.h -- C ABI of the B200-native direct photometric tracker (libuwtrack.so).
 *
 * Drop-in boundary for the hot path of MecatronicaUSB/uw-slam: the calls `uw::System`
 * makes on `uw::Tracker` for the direct (photometric) tracker, plus the pyramid loop of
 * `System::AddFrame`.  Citations are file:line into the reference tree.
 *
 *   reference interface                                        replaced by
 *   ---------------------------------------------------------  --------------------------
 *   Tracker::Tracker(bool)              include/Tracker.h:97    uwt_create
 *   Tracker::InitializePyramid(w,h,K)   include/Tracker.h:112   uwt_create (cfg intrinsics),
 *                                       src/Tracker.cpp:297     uwt_get_level_info
 *   System::AddFrame pyramid loop       src/System.cpp:246-251  uwt_upload_frames,
 *                                                               uwt_set_frames_device
 *   Tracker::ApplyGradient(Frame*)      include/Tracker.h:137   uwt_apply_gradient
 *   Tracker::ObtainCandidatePoints(F*)  include/Tracker.h:145   uwt_select_candidates
 *   Tracker::ObtainAllPoints(Frame*)    include/Tracker.h:153   uwt_select_candidates with
 *                                       src/Tracker.cpp:1259    cfg.depth_mode = UWT_DEPTH_ALL_POINTS
 *   Tracker::EstimatePose(F*,F*)        include/Tracker.h:122   uwt_estimate_pose
 *   Tracker::WarpFunction(Mat,SE3,int)  include/Tracker.h:193   uwt_warp_points
 *   CameraModel::GetCameraModel (rectify) src/CameraModel.cpp:84 uwt_camera_optimal_matrix,
 *                                                               uwt_camera_undistort_maps
 *   CameraModel::Undistort              src/CameraModel.cpp:101 uwt_undistort_image
 *   System::CalculateROI                src/System.cpp:148      uwt_calculate_roi
 *   System::AddFrame remap + ROI crop   src/System.cpp:232-235  uwt_set_undistortion
 *   uw::Frame members (images_, gradientX_, gradientY_,        uwt_get_image, uwt_get_gradients,
 *     gradient_, candidatePoints_)      include/System.h:63-103 uwt_get_candidates
 *
 * Conventions: every call returns 0 on success or a negative UWT_E_* code and never
 * exits, aborts or prints; the caller owns host buffers, the library owns device memory;
 * a pose is 7 floats in Sophus storage order [qx qy qz qw tx ty tz]
 * (thirdparty/sophus/se3.hpp:469-472); one handle = one CUDA device + one stream; calls on
 * one handle must be serialised by the caller, different handles may be used concurrently.
 * A "slot" is the device-side storage of one uw::Frame.  All array calls take `n` slots so
 * that n independent tracking problems are processed by one set of kernel launches.
 * There is no CPU fallback: every entry point fails with UWT_E_CUDA if no device is usable.
 */
#ifndef UWTRACK_H_
#define UWTRACK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UWT_MAX_LEVELS 7

#define UWT_OK 0
#define UWT_E_INVALID (-1) /* bad argument / unsupported configuration */
#define UWT_E_CUDA (-2)    /* CUDA runtime error (see uwt_last_error)   */
#define UWT_E_STATE (-3)   /* call order violated (e.g. no gradients)   */
#define UWT_E_NOMEM (-4)

/* How `deltaMat = A.inv() * b` (src/Tracker.cpp:564) is evaluated. */
#define UWT_SOLVE_LU 0      /* cv::solve(A, b, DECOMP_LU) -- what cv::MatExpr folds it to */
#define UWT_SOLVE_INVERSE 1 /* cv::invert(A, DECOMP_LU) followed by the 6x6 * 6x1 gemm    */
/* North-star option, not in the reference (its "LM_lambda" / ldlt lines are comments,
 * src/Tracker.cpp:546-549): Levenberg-Marquardt damping A_ii += lm_lambda * A_ii and a float
 * Cholesky solve; a matrix that is not positive definite gives delta = 0. */
#define UWT_SOLVE_CHOLESKY_LM 2

/* Residual weights W of the Gauss-Newton step (src/Tracker.cpp:495-496). */
#define UWT_WEIGHT_IDENTITY 0 /* IdentityWeights: what the reference ships (Tracker.cpp:495)    */
#define UWT_WEIGHT_TUKEY 1    /* TukeyFunctionWeights with the MAD scale, the commented alternative
                                 (Tracker.cpp:496, 1571-1594, 1607-1654): a second pass per sweep
                                 builds the residual histogram the median / MAD come from        */
#define UWT_WEIGHT_HUBER 2    /* north-star option (not in the reference): w = 1 for |r| <= delta,
                                 delta/|r| beyond; sqrt(w) scales Jacobian rows and residuals    */

/* Depth input: Tracker::Tracker(depth_available), the depth pyramid of
 * System::AddFrame (src/System.cpp:241-250) and the depth branch of ObtainCandidatePoints
 * (src/Tracker.cpp:1338-1347): candidates need depth != 0 and carry Z = depth * 0.0002. */
#define UWT_DEPTH_NONE 0      /* mono: Z = 1 for every point (src/Tracker.cpp:1349-1355)          */
#define UWT_DEPTH_REFERENCE 1 /* as shipped: depths_[lvl].at<uchar>(y,x) on the 16-bit image, i.e.
                                 BYTE x of row y (low / high byte of depth pixel x/2)             */
#define UWT_DEPTH_U16 2       /* what the code evidently means: at<ushort>(y,x)                   */
#define UWT_DEPTH_ALL_POINTS 3 /* Tracker::ObtainAllPoints (src/Tracker.cpp:1259-1310) in place of
                                 ObtainCandidatePoints: EVERY pixel whose depth, read as
                                 at<short>(y,x), is > 0 is a point, with Z = depth * (0.0002 /
                                 2^level); no gradient threshold.  The reference appends a row
                                 (0,0,1,0) for every other pixel: such a row warps to (0,0), fails
                                 the strict bounds test of EstimatePose (src/Tracker.cpp:450) and
                                 contributes nothing, so those rows are not materialised here
                                 (n_points counts the points with depth); points are kept in the
                                 x-major order of the other modes                                */

/* Derivative stencil of Tracker::ApplyGradient (src/Tracker.cpp:1133-1134). */
#define UWT_GRADIENT_SCHARR 0 /* cv::Scharr, the reference                                        */
#define UWT_GRADIENT_SOBEL 1  /* cv::Sobel(ksize = 3), the north-star wording; not the reference  */

/* How the target image is sampled at the warped position (src/Tracker.cpp:472). */
#define UWT_SAMPLE_NEAREST 0  /* image2.at<uchar>(round(y2), round(x2)), the reference            */
#define UWT_SAMPLE_BILINEAR 1 /* north-star wording: float interpolation of the four neighbours;
                                 identity weights, mono input                                    */

/* cfg.flags */
#define UWT_FLAG_TRACE 1u /* record a per-iteration trace (uwt_get_trace); debugging/parity */
/* A/B switch: hold the 8x8 Gram accumulator [J | 50r]^T [J | 50r] in fp64 tensor-core (DMMA
 * m8n8k4) fragments per warp instead of 27 fp64 registers per thread.  Bit-identical results;
 * measured SLOWER on B200 (1.94 vs 1.19 ms per 128 problems), so it is off by default. */
#define UWT_FLAG_DMMA_ACCUM 2u
/* A/B switch: keep batches on the one-cluster-per-problem kernel instead of the persistent
 * dataflow kernel (chunk tasks from a global ring) that batches of >= 24 problems use. */
#define UWT_FLAG_CLUSTER_KERNEL 4u
/* Opt-in: uwt_apply_gradient / uwt_select_candidates work only on the levels uwt_estimate_pose
 * optimises ([last_level, first_level]); gradient_ / candidatePoints_ of the other levels (level 0
 * by default: 75 % of the pixels, never read by Tracker::EstimatePose, src/Tracker.cpp:389) are
 * materialised by the same kernels when a read-back accessor asks for them.  Every result is
 * identical to the default (eager) mode, which does what the reference does on all levels. */
#define UWT_FLAG_LAZY_LEVELS 8u
/* A/B switch: keep the pyramid (uwt_upload_frames) and the gradient images (uwt_apply_gradient)
 * in separate kernels.  By default a freshly uploaded frame gets its pyramid AND the gradient
 * images of all levels from one fused kernel (one read of the frame, staged by a 2-D tensor
 * copy), and uwt_apply_gradient has nothing left to launch for it; results are identical. */
#define UWT_FLAG_SEPARATE_GRADIENT 16u

typedef struct uwt_tracker uwt_tracker;

typedef struct {
  int width, height;         /* level-0 size, divisible by 2^(levels-1)                  */
  float fx, fy, cx, cy;      /* CameraModel::GetK(), src/CameraModel.cpp:113-115         */
  int levels;                /* PYRAMID_LEVELS, src/Options.cpp:26 (5)                   */
  int first_level;           /* src/Tracker.cpp:368 (levels-1)                           */
  int last_level;            /* src/Tracker.cpp:369 (1)                                  */
  int max_iterations;        /* src/Tracker.cpp:366 (50)                                 */
  float epsilon;             /* src/Tracker.cpp:364 (0.001)                              */
  float residual_scale;      /* src/Tracker.cpp:559 (50)                                 */
  double gradient_threshold; /* GRADIENT_THRESHOLD, src/Options.cpp:27 (20)              */
  int solve_mode;            /* UWT_SOLVE_*                                              */
  int device;                /* CUDA device ordinal                                      */
  int max_frames;            /* number of frame slots to allocate                        */
  int cluster_size;          /* CTAs cooperating on ONE problem in uwt_estimate_pose:
                                0 = choose from n (1 for large batches, 16 for n == 1)   */
  unsigned flags;            /* UWT_FLAG_*                                               */
  int weight_mode;           /* UWT_WEIGHT_* (0 = reference)                             */
  float huber_delta;         /* UWT_WEIGHT_HUBER threshold in gray levels                */
  int depth_mode;            /* UWT_DEPTH_* (0 = mono, the reference's default run mode)  */
  float lm_lambda;           /* UWT_SOLVE_CHOLESKY_LM damping (0.2 in the reference's comment) */
  int gradient_op;           /* UWT_GRADIENT_* (0 = Scharr, the reference)               */
  int sampling;              /* UWT_SAMPLE_* (0 = nearest, the reference)                */
} uwt_config;

typedef struct {
  int width, height;
  float fx, fy, cx, cy, invfx, invfy;
} uwt_level_info;

typedef struct {
  int iterations[UWT_MAX_LEVELS];  /* Gauss-Newton updates applied per level         */
  int evaluations[UWT_MAX_LEVELS]; /* residual sweeps per level                      */
  int n_points[UWT_MAX_LEVELS];    /* candidate points used per level                */
  float final_error[UWT_MAX_LEVELS];
} uwt_track_stats;

/* Per-sweep trace record (UWT_FLAG_TRACE): the quantities of one Gauss-Newton iteration. */
typedef struct {
  int level, k, n_valid, broke;
  long long sum_r2;
  float error;
  float A[36];
  float b[6];
  float delta[6];
  float pose[7];
} uwt_iter_trace;

/* Fills cfg with the reference's defaults (640x480 TUM calibration, 5/4/1 levels). */
int uwt_default_config(uwt_config* cfg);
/* Tracker::Tracker + Tracker::InitializePyramid.  On failure *out is NULL and the message
 * is available from uwt_last_error(NULL). */
int uwt_create(const uwt_config* cfg, uwt_tracker** out);
int uwt_destroy(uwt_tracker* t);
/* Message of the last failure on this handle (t == NULL: of the last uwt_create). */
const char* uwt_last_error(const uwt_tracker* t);
/* Per-level geometry and intrinsics computed as in src/Tracker.cpp:297-340. */
int uwt_get_level_info(const uwt_tracker* t, int level, uwt_level_info* info);
/* The CUDA stream (cudaStream_t) all work of this handle is enqueued on. */
void* uwt_stream(const uwt_tracker* t);
int uwt_synchronize(uwt_tracker* t);

/* System::AddFrame: copy n gray frames from HOST memory (frame i at host + i*frame_stride,
 * rows row_stride bytes apart) into slots[i] and build their pyramids.  Asynchronous with
 * respect to the host when `host` is pinned (the buffer must stay valid until the copy has
 * run); the copy uses its own stream and double-buffered staging, so it overlaps kernels
 * already enqueued on the handle.  Resets the slots' gradient/candidate state. */
int uwt_upload_frames(uwt_tracker* t, int n, const int* slots, const uint8_t* host,
                      size_t row_stride, size_t frame_stride);
/* Same, from frames already resident in DEVICE memory. */
int uwt_set_frames_device(uwt_tracker* t, int n, const int* slots, const uint8_t* dev,
                          size_t row_stride, size_t frame_stride);
/* ---- calibration / undistortion front-end ----
 * CameraModel::GetCameraModel, rectify branch (src/CameraModel.cpp:84-98): the two one-off
 * OpenCV calls, on the host (K9 / newK9: row-major 3x3 CV_32F like the reference's Mats;
 * dist4 = k1 k2 p1 p2 of calibration/<name>.xml <rectification>; results follow OpenCV 4.x):
 *   getOptimalNewCameraMatrix(K, dist, in_size, alpha, out_size, nullptr, false)
 *   initUndistortRectifyMap(K, dist, Mat(), newK, out_size, CV_16SC2, map1, map2)
 * map1: out_h*out_w*2 int16 (integer source x, y); map2: out_h*out_w uint16 ((fy<<5)|fx). */
int uwt_camera_optimal_matrix(const float* K9, const float* dist4, int in_w, int in_h,
                              double alpha, int out_w, int out_h, float* newK9);
int uwt_camera_undistort_maps(const float* K9, const float* dist4, const float* newK9, int out_w,
                              int out_h, int16_t* map1, uint16_t* map2);
/* CameraModel::Undistort (src/CameraModel.cpp:101-103): cv::remap(INTER_LINEAR, fixed-point
 * maps, constant border 0) of ONE host image on `device`; no handle needed (used once, for the
 * ROI search below, before the tracker size is known). */
int uwt_undistort_image(int device, const uint8_t* src, int in_w, int in_h, size_t src_stride,
                        const int16_t* map1, const uint16_t* map2, int out_w, int out_h,
                        uint8_t* dst);
/* System::CalculateROI (src/System.cpp:148-191) on the first undistorted image (host):
 * roi4 = x, y of the top-left corner and the new w_, h_. */
int uwt_calculate_roi(const uint8_t* undistorted, int w, int h, size_t stride, int* roi4);
/* System::AddFrame, `remap(...); images_[0] = distortion(ROI)` (src/System.cpp:232-235): after
 * this call uwt_upload_frames / uwt_set_frames_device take the DISTORTED in_w x in_h frames;
 * the remap and the crop at (roi_x, roi_y) are fused into the level-0 load of the pyramid
 * kernel.  The maps are map_w x map_h (the reference's out_width x out_height); the handle's
 * width x height is the cropped size.  map1 == map2 == NULL switches undistortion off. */
int uwt_set_undistortion(uwt_tracker* t, const int16_t* map1, const uint16_t* map2, int map_w,
                         int map_h, int in_w, int in_h, int roi_x, int roi_y);

/* System::AddFrame with depth_available_ (src/System.cpp:241-250): copy n 16-bit depth frames
 * (host or device memory, rows row_stride BYTES apart) into slots[i] and build their pyramids (cv::resize of
 * CV_16U by 0.5 = 2x2 mean, ties to even).  Needs cfg.depth_mode != 0; call it after
 * uwt_upload_frames for the same slots and before uwt_select_candidates. */
int uwt_upload_depth_frames(uwt_tracker* t, int n, const int* slots, const uint16_t* host,
                            size_t row_stride, size_t frame_stride);
/* depths_[level] of a slot, dense row-major host buffer. */
int uwt_get_depth(uwt_tracker* t, int slot, int level, uint16_t* host);

/* Tracker::ApplyGradient for n slots: gradient_ (u8) on every pyramid level.  The int16
 * gradientX_ / gradientY_ values reach the tracker through the packed candidate records; the
 * full planes are produced on demand by uwt_get_gradients (same stencil, same integers).
 * A slot filled by uwt_upload_frames / uwt_set_frames_device already carries its gradient images
 * (the fused frame kernel, see UWT_FLAG_SEPARATE_GRADIENT): the call then launches nothing. */
int uwt_apply_gradient(uwt_tracker* t, int n, const int* slots);
/* Tracker::ObtainCandidatePoints for n slots (needs gradients): candidatePoints_ on every
 * level, in the reference's x-major order. */
int uwt_select_candidates(uwt_tracker* t, int n, const int* slots);
/* Tracker::EstimatePose for n independent (prev, cur) pairs.  prev slots need candidates,
 * cur slots need a pyramid.  init_poses7 (n*7 floats) may be NULL = identity, which is the
 * reference (src/Tracker.cpp:385).  out_poses7 (n*7 floats, host) receives
 * prev->rigid_transformation_; stats (n entries, host) may be NULL.  Synchronous. */
int uwt_estimate_pose(uwt_tracker* t, int n, const int* prev_slots, const int* cur_slots,
                      const float* init_poses7, float* out_poses7, uwt_track_stats* stats);
/* Asynchronous form: results land in library-owned pinned memory; uwt_fetch_poses waits for
 * this estimate only (not for work enqueued after it) and copies them out.  Lets the host
 * enqueue the next upload before reading the poses of the current batch. */
int uwt_estimate_pose_async(uwt_tracker* t, int n, const int* prev_slots,
                            const int* cur_slots, const float* init_poses7);
int uwt_fetch_poses(uwt_tracker* t, int n, float* out_poses7, uwt_track_stats* stats);
/* Single-huge-frame mode: the Gauss-Newton loop of ONE (prev, cur) pair split over `nranks`
 * handles, one per GPU (each holds both frames; rank r sweeps the r-th contiguous range of
 * every level's candidate list).  Per sweep, on every rank:
 *     uwt_shard_accumulate(t, d_sums);          // this rank's 32 fp64 partial sums
 *     <sum d_sums over all ranks, e.g. ncclAllReduce(.., ncclDouble, ncclSum) on uwt_stream(t)>
 *     uwt_shard_update(t, d_sums, &done);       // break test, solve, update -- identical
 *                                               // on every rank, no broadcast needed
 * until done != 0, then uwt_shard_result.  d_sums32: 32 doubles in DEVICE memory owned by the
 * caller (layout: 21 upper-triangular J^T J terms, 6 J^T r terms, sum r^2, N_valid, the weighted
 * error term r^T W r of the Huber mode, 2 pad).  Weights: identity or Huber (a fixed function of
 * the integer residual: every rank builds the same tables); Tukey / MAD weights depend on the
 * sweep's residual histogram over ALL ranks and are rejected here (UWT_E_INVALID). */
int uwt_shard_begin(uwt_tracker* t, int prev_slot, int cur_slot, int rank, int nranks,
                    const float* init_pose7);
int uwt_shard_accumulate(uwt_tracker* t, double* d_sums32);
int uwt_shard_update(uwt_tracker* t, const double* d_sums32, int* done);
int uwt_shard_result(uwt_tracker* t, float* out_pose7, uwt_track_stats* stats);
/* Fused compute + collective form of the same mode: one persistent kernel per rank runs the
 * whole loop and all-reduces the 32 sums per sweep by storing them into the peers' mailboxes
 * over NVLink (no NCCL call, no host round trip per sweep).  Setup once per group:
 *   every rank: uwt_shard_ipc_export -> 64-byte handle; all-gather the handles (any transport);
 *   every rank: uwt_shard_ipc_connect(t, rank, nranks, handles)      [ranks in other processes]
 *   or uwt_shard_connect_local(t, rank, nranks, peers)               [handles of one process]
 * then per estimate, on every rank: uwt_shard_estimate_fused_async + _wait (all ranks must
 * call it; a rank whose peers never arrive fails after a bounded wait instead of hanging).
 * grid = CTAs of the persistent kernel (0 = 148, one per SM). */
int uwt_shard_ipc_handle_size(void);
int uwt_shard_ipc_export(uwt_tracker* t, void* handle_out);
int uwt_shard_ipc_connect(uwt_tracker* t, int rank, int nranks, const void* all_handles);
int uwt_shard_connect_local(uwt_tracker* t, int rank, int nranks, uwt_tracker* const* peers);
int uwt_shard_estimate_fused_async(uwt_tracker* t, int prev_slot, int cur_slot,
                                   const float* init_pose7, int grid);
int uwt_shard_estimate_fused_wait(uwt_tracker* t, float* out_pose7, uwt_track_stats* stats);
/* Tracker::WarpFunction on host points (n x 4 floats [x y Z W]) at a pyramid level. */
int uwt_warp_points(uwt_tracker* t, const float* pts4, int n, const float* pose7, int level,
                    float* out4);

/* Read-back accessors (uw::Frame members), dense row-major host buffers. */
int uwt_get_image(uwt_tracker* t, int slot, int level, uint8_t* host);
int uwt_get_gradients(uwt_tracker* t, int slot, int level, int16_t* gx, int16_t* gy,
                      uint8_t* g);
int uwt_get_candidate_count(uwt_tracker* t, int slot, int level, int* n);
/* candidatePoints_[level] as the reference stores it: rows [x, y, 1, 1] (CV_32FC1). */
int uwt_get_candidates(uwt_tracker* t, int slot, int level, float* pts4, int capacity_rows,
                       int* n);
/* The packed per-candidate records the Gauss-Newton kernel streams (levels first..last only),
 * same order as uwt_get_candidates.  Bit layout of one 64-bit record:
 *   0..11 x | 12..23 y | 24..31 I1 = images_[l](y,x) | 32..44 gradientX_[l](y,x) (13-bit two's
 *   complement) | 45..57 gradientY_[l](y,x) | 58..63 zero. */
int uwt_get_records(uwt_tracker* t, int slot, int level, uint64_t* packed, int capacity, int* n);
/* Trace of problem `index` of the last uwt_estimate_pose (needs UWT_FLAG_TRACE).  Traces are
 * recorded for batches of at most 64 problems; after a larger batch (not traced) the call returns
 * UWT_E_STATE instead of an earlier batch's rows. */
int uwt_get_trace(uwt_tracker* t, int index, uwt_iter_trace* out, int capacity, int* n);
/* Number of compute kernels this handle has launched so far. */
long long uwt_launch_count(const uwt_tracker* t);
/* Number of argument-staging kernels launched so far: the slot / pose arrays of a call reach the
 * device through a one-CTA copy kernel (not a host-to-device copy, which would queue behind frame
 * uploads).  bench.py's gpu_launches = uwt_launch_count + uwt_aux_launch_count. */
long long uwt_aux_launch_count(const uwt_tracker* t);

/* Per-kernel-class device timing with CUDA events on the handle's stream (bench.py's
 * roofline numbers).  Classes: */
#define UWT_K_PYRAMID 0    /* K1 pyramid                          */
#define UWT_K_GRADIENT 1   /* K2 Scharr + gradient image + sum    */
#define UWT_K_CANDIDATES 2 /* K3 count + scan + scatter           */
#define UWT_K_ESTIMATE 3   /* K4/K5 Gauss-Newton pose estimate    */
#define UWT_K_COUNT 4
/* on != 0: start (and reset) timing of every subsequent kernel class call; 0: stop. */
int uwt_profile_enable(uwt_tracker* t, int on);
/* Synchronises and returns the accumulated milliseconds and kernel launches per class. */
int uwt_profile_read(uwt_tracker* t, double ms[UWT_K_COUNT], long long launches[UWT_K_COUNT]);

#ifdef __cplusplus
}
#endif
#endif /* UWTRACK_H_ */
