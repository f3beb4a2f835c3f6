/*
 * uw_oracle.h -- CPU ORACLE for the UW-SLAM direct photometric tracker.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain C++ restatement of the reference's
 * algorithm (file:line citations are into /root/reference) used as the parity
 * checker in tests/, in __graft_entry__.smoke() and as bench.py's cpu_baseline /
 * --impl reference leg.  Nothing on the product path (uw_slam_b200/, include/)
 * links, imports or calls it.
 *
 * PARITY STATUS: "parity unpinned at the source" -- the reference ships no tests,
 * golden vectors or fixtures and cannot be compiled here
 * (needs OpenCV 3.2 C++, Eigen, Ceres, ROS).  Every OpenCV primitive the path
 * uses is instead pinned bit-for-bit against the real OpenCV (python cv2 4.13)
 * in tests/test_oracle_vs_cv2.py; the float op order of the few Eigen/Sophus
 * expressions is a documented canonical choice (docs/ARITHMETIC.md).
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UWO_MAX_LEVELS 8

/* solve_mode: how `A.inv() * b` (Tracker.cpp:564) is evaluated. */
#define UWO_SOLVE_LU 0      /* cv::solve(A,b,DECOMP_LU): what OpenCV's MatExpr does */
#define UWO_SOLVE_INVERSE 1 /* cv::invert(A) then gemm: the literal reading         */
#define UWO_SOLVE_CHOLESKY_LM 2 /* north-star option: LM damping + float Cholesky (ARITHMETIC S2) */

/* weight_mode: residual weights W of the Gauss-Newton step (Tracker.cpp:495-496). */
#define UWO_WEIGHT_IDENTITY 0 /* IdentityWeights: what the reference ships (Tracker.cpp:495)  */
#define UWO_WEIGHT_TUKEY 1    /* TukeyFunctionWeights + MAD scale: the commented alternative
                                 (Tracker.cpp:496, 1571-1594, 1607-1654)                       */
#define UWO_WEIGHT_HUBER 2    /* north-star option, not in the reference: w = 1 (|r| <= d),
                                 d/|r| otherwise, applied as sqrt(w) to J rows and residuals   */

/* accum_mode: accumulator of the normal equations (docs/ARITHMETIC.md U3). */
#define UWO_ACCUM_DOUBLE 0     /* sequential fp64 (timed baseline)     */
#define UWO_ACCUM_LONGDOUBLE 1 /* sequential 80-bit (checker, default) */

typedef struct {
  int width, height;          /* level-0 size; must be divisible by 2^(levels-1) */
  float fx, fy, cx, cy;       /* level-0 pinhole intrinsics (CameraModel::GetK)   */
  int levels;                 /* PYRAMID_LEVELS, Options.cpp:26 (5)               */
  int first_level;            /* Tracker.cpp:368 (levels-1 = 4)                   */
  int last_level;             /* Tracker.cpp:369 (1)                              */
  int max_iterations;         /* Tracker.cpp:366 (50)                             */
  float epsilon;              /* Tracker.cpp:364 (0.001)                          */
  float residual_scale;       /* Tracker.cpp:559 (50)                             */
  double gradient_threshold;  /* Options.cpp:27 (20)                              */
  int solve_mode;             /* UWO_SOLVE_*                                      */
  int accum_mode;             /* UWO_ACCUM_*                                      */
  int threads;                /* 1 = like the reference; >1 = std::thread over points  */
  int weight_mode;            /* UWO_WEIGHT_* (Tracker.cpp:495-496)                */
  float huber_delta;          /* UWO_WEIGHT_HUBER threshold in gray levels         */
  float lm_lambda;            /* UWO_SOLVE_CHOLESKY_LM damping                      */
  int sampling;               /* 0 nearest (Tracker.cpp:472), 1 bilinear (north-star, B1) */
} uwo_params;

/* One record per Gauss-Newton iteration (including the breaking one). */
typedef struct {
  int level, k;
  int n_valid;
  int broke;                  /* 1 if this iteration hit the break test           */
  long long sum_r2;           /* exact integer sum of squared residuals           */
  float error;                /* Tracker.cpp:499-502                              */
  float A[36];                /* row-major 6x6 (zero when broke)                  */
  float b[6];
  float delta[6];
  float pose[7];              /* pose AFTER this iteration: qx qy qz qw tx ty tz  */
} uwo_iter_trace;

typedef struct {
  int iterations[UWO_MAX_LEVELS];   /* GN updates applied per level               */
  int evaluations[UWO_MAX_LEVELS];  /* residual sweeps per level (= updates + 1)  */
  int n_points[UWO_MAX_LEVELS];
  float final_error[UWO_MAX_LEVELS];
} uwo_stats;

void uwo_default_params(uwo_params* p);

/* System::AddFrame pyramid loop, System.cpp:246-251: exact-half cv::resize. */
void uwo_pyr_down(const uint8_t* src, int w, int h, uint8_t* dst);
/* Tracker::ApplyGradient, Tracker.cpp:1133-1134: Scharr 16S, BORDER_REFLECT_101. */
void uwo_scharr(const uint8_t* img, int w, int h, int16_t* gx, int16_t* gy);
/* north-star wording: cv::Sobel(img, CV_16S, 1, 0 / 0, 1, 3), BORDER_REFLECT_101 (not the reference) */
void uwo_sobel(const uint8_t* img, int w, int h, int16_t* gx, int16_t* gy);
/* Tracker.cpp:1139-1142: convertScaleAbs x2 + addWeighted(.5,.5). */
void uwo_gradmag(const int16_t* gx, const int16_t* gy, long long n, uint8_t* g);
/* Tracker::ObtainCandidatePoints, Tracker.cpp:1314-1357 (mono branch).  pts4 must
 * hold w*h*4 floats.  Returns the number of candidates; rows are [x,y,1,1], x-major. */
int uwo_candidates(const uint8_t* g, int w, int h, double gradient_threshold,
                   float* pts4, double* mean_out, int* ithr_out);
/* ---- depth input ----
 * System::AddFrame depth pyramid (System.cpp:248-250): cv::resize(0.5, 0.5) of a CV_16U image =
 * 2x2 area mean rounded half-to-even (cvRound of sum * 0.25). */
void uwo_depth_pyr_down(const uint16_t* src, int w, int h, uint16_t* dst);
/* depth_mode of the candidate selection (Tracker.cpp:1338-1347): */
#define UWO_DEPTH_NONE 0      /* mono: Z = depth_initialization = 1 (Tracker.cpp:1349-1355)      */
#define UWO_DEPTH_REFERENCE 1 /* as shipped: depths_[lvl].at<uchar>(y,x), i.e. BYTE x of row y of
                                 the 16-bit image (low/high byte of pixel x/2), Z = byte * 0.0002  */
#define UWO_DEPTH_U16 2       /* what the code evidently means: at<ushort>(y,x), Z = d * 0.0002   */
/* Tracker::ObtainCandidatePoints with depth_available_ (Tracker.cpp:1338-1347): rows [x,y,Z,1]
 * for pixels with filtered != 0 AND depth != 0.  zsrc (optional) receives the integer depth used. */
int uwo_candidates_depth(const uint8_t* g, const uint16_t* depth, int w, int h,
                         double gradient_threshold, int depth_mode, float* pts4, uint16_t* zsrc);

/* Tracker::InitializePyramid, Tracker.cpp:297-340. Arrays of length `levels`. */
void uwo_init_pyramid(int w, int h, float fx, float fy, float cx, float cy, int levels,
                      int* wl, int* hl, float* fxl, float* fyl, float* cxl, float* cyl,
                      float* invfxl, float* invfyl);
/* Tracker::WarpFunction, Tracker.cpp:1417-1471. */
void uwo_warp(const float* pts4, int n, const float* pose7, float fx, float fy, float cx,
              float cy, float invfx, float invfy, float* out4);

/* Sophus pieces (se3.hpp:723-744, 317-321; so3.hpp:338-355, 434-440, 534-568). */
void uwo_se3_exp(const float* tangent6, float* pose7);
void uwo_se3_mul(const float* a7, const float* b7, float* out7);
void uwo_se3_matrix(const float* pose7, float* m16);
void uwo_se3_scale_level(const float* pose7, float* out7); /* Tracker.cpp:580-590 */

/* Tracker::MedianMat (Tracker.cpp:1571-1594): convertTo(CV_8UC1) + 256-bin calcHist + first
 * bin whose cumulative count exceeds n/2.  Returns -1 for n == 0. */
float uwo_median_mat(const float* v, int n);
/* Tracker::MedianAbsoluteDeviation (Tracker.cpp:1607-1619): 1.4826 * MedianMat(|v - MedianMat(v)|). */
float uwo_mad(const float* v, int n);
/* Tracker::TukeyFunctionWeights (Tracker.cpp:1626-1654): w[i] for residuals v[i]. */
void uwo_tukey_weights(const float* v, int n, float* w);
/* Huber weights (north-star option; docs/ARITHMETIC.md R4). */
void uwo_huber_weights(const float* v, int n, float delta, float* w);

/* Visualizer::UpdateMessages pose composition (Visualizer.cpp:303-325): current = SE3(q, scale*t)
 * (normalising constructor, se3.hpp:446-448), out = previous * current. */
void uwo_chain_pose(const float* previous7, const float* rigid7, float scale, float* out7);

/* LM-damped float Cholesky solve (north-star option).  Returns 0 if not positive definite. */
int uwo_cholesky_lm_solve6(const float* A36, const float* b6, float lambda, float* x6);
/* cv::solve / cv::invert (DECOMP_LU) on 6x6 f32.  Return 0 if singular. */
int uwo_lu_solve6(const float* A36, const float* b6, float* x6);
int uwo_lu_invert6(const float* A36, float* Ainv36);

/* ---- calibration / undistortion front-end ----------------------------
 * CameraModel::GetCameraModel rectify branch (CameraModel.cpp:84-98): the OpenCV calls it makes,
 * restated after OpenCV 4.x (the version that can be pinned here, cv2 4.13). */
/* cv::getOptimalNewCameraMatrix(K, dist(k1 k2 p1 p2), in_size, alpha, out_size, 0, false).
 * K, newK: row-major 3x3 float (CV_32F like the reference's Mats). */
void uwo_optimal_new_camera_matrix(const float* K9, const float* dist4, int in_w, int in_h,
                                   double alpha, int out_w, int out_h, float* newK9);
/* cv::initUndistortRectifyMap(K, dist, Mat(), newK, out_size, CV_16SC2, map1, map2):
 * map1 = out_h x out_w x 2 int16 (integer source x, y), map2 = out_h x out_w uint16
 * (5-bit fractions: (fy << 5) | fx). */
void uwo_init_undistort_rectify_map(const float* K9, const float* dist4, const float* newK9,
                                    int out_w, int out_h, int16_t* map1, uint16_t* map2);
/* cv::remap(src, dst, map1, map2, INTER_LINEAR) for 8-bit images with fixed-point maps
 * (BORDER_CONSTANT 0), CameraModel.cpp:101-103 / System.cpp:232-234. */
void uwo_remap_bilinear(const uint8_t* src, int sw, int sh, const int16_t* map1,
                        const uint16_t* map2, int dw, int dh, uint8_t* dst);
/* System::CalculateROI (System.cpp:148-191) on the first undistorted image: roi4 = x, y of the
 * top-left corner and the new w_, h_ (p2 - p1).  Returns 0, or -1 if a scan leaves the image. */
int uwo_calculate_roi(const uint8_t* undistorted, int w, int h, int* roi4);

/* Whole-frame preprocessing into caller-provided per-level arrays.
 * images[l]: u8 w_l*h_l (level 0 is input, 1.. are written). */
void uwo_build_pyramid(const uwo_params* p, uint8_t* const* images);

/* Tracker::EstimatePose, Tracker.cpp:362-597.
 * prev_images/cur_images/gx/gy: per-level pointers (levels entries);
 * cand[l]: N_l x 4 f32, ncand[l]: N_l.  init_pose7 may be NULL (identity).
 * trace may be NULL; at most trace_cap records are written, *n_trace gets the count. */
int uwo_estimate_pose(const uwo_params* p, const uint8_t* const* prev_images,
                      const uint8_t* const* cur_images, const int16_t* const* gx,
                      const int16_t* const* gy, const float* const* cand, const int* ncand,
                      const float* init_pose7, float* out_pose7, uwo_stats* stats,
                      uwo_iter_trace* trace, int trace_cap, int* n_trace);

/* The two halves of one Gauss-Newton iteration, exposed so that the sharded (multi-rank) host
 * protocol can be exercised on CPU: a residual sweep over candidate rows [lo,hi) -> 32 sums,
 * and the update on (reduced) sums.  uwo_estimate_pose is built from these two. */
int uwo_sweep_range(const uwo_params* p, int lvl, const uint8_t* I1, const uint8_t* I2,
                    const int16_t* gx, const int16_t* gy, const float* cand, int lo, int hi,
                    const float* pose7, double* sums32);
int uwo_gn_update(const uwo_params* p, const double* sums32, int k, float* pose7,
                  float* last_error, uwo_iter_trace* tr);

/* Convenience for benchmarks: full track on two level-0 frames (allocates internally):
 * pyramid(prev), pyramid(cur), ApplyGradient(prev), ObtainCandidatePoints(prev),
 * EstimatePose(prev,cur).  seconds[0..3] = pyramid, gradient, candidates, estimate. */
int uwo_track_pair(const uwo_params* p, const uint8_t* prev0, const uint8_t* cur0,
                   float* out_pose7, uwo_stats* stats, double* seconds4);

/* The reference's outer loop (main_uw_slam.cpp:139-151: AddFrame + Tracking per frame) as a
 * stateful stream: create() prepares frame 0 (pyramid, ApplyGradient, ObtainCandidatePoints);
 * track() is one track = pyramid(i), EstimatePose(i-1,i), ApplyGradient(i),
 * ObtainCandidatePoints(i), and keeps frame i as the next `prev`. */
typedef struct uwo_stream uwo_stream;
uwo_stream* uwo_stream_create(const uwo_params* p, const uint8_t* frame0);
void uwo_stream_destroy(uwo_stream* s);
int uwo_stream_track(uwo_stream* s, const uint8_t* frame, float* out_pose7, uwo_stats* stats);

/* Reference-shaped loop over a sequence of n_frames level-0 frames (contiguous, w*h bytes
 * each): one "track" per frame after the first = pyramid(i), EstimatePose(i-1,i),
 * ApplyGradient(i), ObtainCandidatePoints(i).  poses_out: (n_frames-1) x 7. */
int uwo_track_sequence(const uwo_params* p, const uint8_t* frames, int n_frames,
                       float* poses_out, double* seconds_out, uwo_stats* stats_out);

#ifdef __cplusplus
}
#endif
