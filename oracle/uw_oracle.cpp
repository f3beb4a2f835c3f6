/*
 * uw_oracle.cpp -- CPU ORACLE (test infrastructure only; see uw_oracle.h header).
 *
 * Restates, with the canonical arithmetic of docs/ARITHMETIC.md, the reference's
 *   System::AddFrame pyramid loop        /root/reference/src/System.cpp:246-251
 *   Tracker::InitializePyramid           src/Tracker.cpp:297-340
 *   Tracker::ApplyGradient               src/Tracker.cpp:1127-1143
 *   Tracker::ObtainCandidatePoints       src/Tracker.cpp:1314-1357
 *   Tracker::WarpFunction                src/Tracker.cpp:1417-1471
 *   Tracker::EstimatePose                src/Tracker.cpp:362-597
 *   Sophus SE3/SO3 pieces                thirdparty/sophus/se3.hpp, so3.hpp
 *
 * Build: g++ -O2 -ffp-contract=off (NO -ffast-math, NO -mfma contraction): every
 * float expression below rounds once per operation, exactly as written.
 * Structure is deliberately the reference's (AoS N x 4 points, a separate warp pass,
 * materialised N x 6 Jacobian and N x 1 residual arrays, then J^T J), so the timed
 * baseline is not an optimised rewrite.  It omits the reference's per-point cv::Mat
 * heap allocations (Tracker.cpp:433-436, 1336-1337) and is therefore FASTER than the
 * real reference would be.
 */
#include "uw_oracle.h"

#include <cfloat>
#include <chrono>
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>

namespace {

// Runs fn(c) for c in [0,nt) on nt host threads (nt==1: inline, like the reference).
template <typename F>
void parallel_chunks(int nt, F fn) {
  if (nt <= 1) {
    fn(0);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nt - 1);
  for (int c = 1; c < nt; ++c) th.emplace_back([&fn, c] { fn(c); });
  fn(0);
  for (auto& t : th) t.join();
}

struct Pose {  // Sophus storage order (se3.hpp:469-472): quaternion x,y,z,w then t
  float q[4];
  float t[3];
};

inline Pose pose_from7(const float* p) {
  Pose r;
  for (int i = 0; i < 4; ++i) r.q[i] = p[i];
  for (int i = 0; i < 3; ++i) r.t[i] = p[4 + i];
  return r;
}
inline void pose_to7(const Pose& p, float* o) {
  for (int i = 0; i < 4; ++i) o[i] = p.q[i];
  for (int i = 0; i < 3; ++i) o[4 + i] = p.t[i];
}

// Eigen redux (no-vectorisation unroller): halves -> (x^2+y^2)+(z^2+w^2)
inline float quat_sqnorm(const float* q) {
  return (q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]);
}

// Eigen Quaternion::toRotationMatrix (U7)
inline void quat_to_R(const float* q, float R[9]) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0f - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1.0f - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1.0f - (txx + tyy);
}

inline void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// Eigen QuaternionBase::_transformVector, used by so3.hpp:320-322
inline void quat_rotate(const float* q, const float* v, float* o) {
  float uv[3], c[3];
  cross3(q, v, uv);
  for (int i = 0; i < 3; ++i) uv[i] = uv[i] + uv[i];
  cross3(q, uv, c);
  for (int i = 0; i < 3; ++i) o[i] = (v[i] + q[3] * uv[i]) + c[i];
}

// Hamilton product, scalar form (U7)
inline void quat_mul(const float* a, const float* b, float* o) {
  const float ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const float bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx;
}

// SE3Base::operator*=, se3.hpp:317-321 + SO3Base::operator*=, so3.hpp:338-355
inline Pose se3_mul(const Pose& a, const Pose& b) {
  Pose r;
  float rt[3];
  quat_rotate(a.q, b.t, rt);
  for (int i = 0; i < 3; ++i) r.t[i] = a.t[i] + rt[i];
  quat_mul(a.q, b.q, r.q);
  const float sn = quat_sqnorm(r.q);
  if (sn != 1.0f) {
    const float s = 2.0f / (1.0f + sn);
    for (int i = 0; i < 4; ++i) r.q[i] = r.q[i] * s;
  }
  return r;
}

// SE3::exp, se3.hpp:723-744 with SO3::expAndTheta, so3.hpp:534-568.
// Transcendentals are evaluated in fp64 and rounded to f32 (U5).
inline Pose se3_exp(const float* a) {
  const float eps = 1e-5f;  // common.hpp:155-158
  const float ox = a[3], oy = a[4], oz = a[5];
  const float theta_sq = ox * ox + (oy * oy + oz * oz);
  const float theta = std::sqrt(theta_sq);
  const float half_theta = 0.5f * theta;
  float imag, real;
  if (theta < eps) {
    const float theta_po4 = theta_sq * theta_sq;
    imag = (0.5f - (float)(1.0 / 48.0) * theta_sq) + (float)(1.0 / 3840.0) * theta_po4;
    real = (1.0f - (float)(1.0 / 8.0) * theta_sq) + (float)(1.0 / 384.0) * theta_po4;
  } else {
    const float s = (float)std::sin((double)half_theta);
    imag = s / theta;
    real = (float)std::cos((double)half_theta);
  }
  Pose r;
  r.q[0] = imag * ox;
  r.q[1] = imag * oy;
  r.q[2] = imag * oz;
  r.q[3] = real;
  // hat(omega), so3.hpp:618-627
  const float O[9] = {0.0f, -oz, oy, oz, 0.0f, -ox, -oy, ox, 0.0f};
  float Osq[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Osq[i * 3 + j] = (O[i * 3 + 0] * O[0 * 3 + j] + O[i * 3 + 1] * O[1 * 3 + j]) +
                       O[i * 3 + 2] * O[2 * 3 + j];
  float V[9];
  if (theta < eps) {
    quat_to_R(r.q, V);
  } else {
    const float tsq = theta * theta;
    const float ca = (1.0f - (float)std::cos((double)theta)) / tsq;
    const float cb = (theta - (float)std::sin((double)theta)) / (tsq * theta);
    for (int i = 0; i < 9; ++i) {
      const float I = (i == 0 || i == 4 || i == 8) ? 1.0f : 0.0f;
      V[i] = (I + ca * O[i]) + cb * Osq[i];
    }
  }
  for (int i = 0; i < 3; ++i)
    r.t[i] = (V[i * 3 + 0] * a[0] + V[i * 3 + 1] * a[1]) + V[i * 3 + 2] * a[2];
  return r;
}

// Tracker.cpp:580-590: q.xyz *= 2; SE3(q,t) -> SO3(quat) normalises (so3.hpp:270-276)
inline Pose se3_scale_level(const Pose& p) {
  Pose r = p;
  r.q[0] = r.q[0] * 2.0f;
  r.q[1] = r.q[1] * 2.0f;
  r.q[2] = r.q[2] * 2.0f;
  const float len = std::sqrt(quat_sqnorm(r.q));
  for (int i = 0; i < 4; ++i) r.q[i] = r.q[i] / len;
  return r;
}

// OpenCV hal::LU32f (LUImpl<float>) on [A | B], partial pivoting, eps = 10*FLT_EPSILON.
// Every multiply and add rounds separately (no FMA).  Returns 0 if singular.
int lu_impl(float* A, int m, float* B, int n) {
  const float eps = 1.1920929e-07f * 10.0f;
  for (int i = 0; i < m; ++i) {
    int k = i;
    for (int j = i + 1; j < m; ++j)
      if (std::fabs(A[j * m + i]) > std::fabs(A[k * m + i])) k = j;
    if (std::fabs(A[k * m + i]) < eps) return 0;
    if (k != i) {
      for (int j = i; j < m; ++j) std::swap(A[i * m + j], A[k * m + j]);
      for (int j = 0; j < n; ++j) std::swap(B[i * n + j], B[k * n + j]);
    }
    const float d = -1.0f / A[i * m + i];
    for (int j = i + 1; j < m; ++j) {
      const float alpha = A[j * m + i] * d;
      for (int c = i + 1; c < m; ++c) A[j * m + c] = A[j * m + c] + alpha * A[i * m + c];
      for (int c = 0; c < n; ++c) B[j * n + c] = B[j * n + c] + alpha * B[i * n + c];
    }
  }
  for (int i = m - 1; i >= 0; --i)
    for (int j = 0; j < n; ++j) {
      float s = B[i * n + j];
      for (int c = i + 1; c < m; ++c) s = s - A[i * m + c] * B[c * n + j];
      B[i * n + j] = s / A[i * m + i];
    }
  return 1;
}

inline int iround_half_away(float v) { return (int)std::round(v); }

template <typename ACC>
void accumulate(const float* J, const float* r50, const unsigned char* valid, int n, int threads,
                double* A21, double* b6) {
  // A = J^T J (upper triangle), b = -J^T (50 r): products of two f32 are exact in fp64;
  // accumulation is sequential in ACC (U3).
  int nt = threads < 1 ? 1 : threads;
  std::vector<ACC> partial((size_t)nt * 27, ACC(0));
  parallel_chunks(nt, [&](int t) {
    const int lo = (int)((long long)n * t / nt), hi = (int)((long long)n * (t + 1) / nt);
    ACC acc[27];
    for (int i = 0; i < 27; ++i) acc[i] = ACC(0);
    for (int i = lo; i < hi; ++i) {
      if (!valid[i]) continue;
      const float* Ji = J + (size_t)i * 6;
      int idx = 0;
      for (int a = 0; a < 6; ++a)
        for (int c = a; c < 6; ++c) acc[idx++] += (ACC)((double)Ji[a] * (double)Ji[c]);
      for (int a = 0; a < 6; ++a) acc[21 + a] += (ACC)((double)Ji[a] * (double)r50[i]);
    }
    for (int i = 0; i < 27; ++i) partial[(size_t)t * 27 + i] = acc[i];
  });
  for (int i = 0; i < 27; ++i) {
    ACC s = ACC(0);
    for (int t = 0; t < nt; ++t) s += partial[(size_t)t * 27 + i];
    if (i < 21)
      A21[i] = (double)s;
    else
      b6[i - 21] = (double)s;
  }
}

// saturate_cast<uchar>(float): cvRound (ties to even) then clamp, as Mat::convertTo(CV_8UC1)
inline int sat_u8(float v) {
  const int i = (int)std::nearbyintf(v);
  return i < 0 ? 0 : (i > 255 ? 255 : i);
}

// Tracker::MedianMat, Tracker.cpp:1571-1594
float median_mat(const float* v, int n) {
  float hist[256];  // calcHist returns CV_32F counts
  {
    int cnt[256] = {0};
    for (int i = 0; i < n; ++i) ++cnt[sat_u8(v[i])];  // convertTo + calcHist, :1573,1585
    for (int i = 0; i < 256; ++i) hist[i] = (float)cnt[i];
  }
  const float m = (float)(n / 2);  // :1575 integer division, stored float
  int bin = 0;
  float med = -1.0f;
  for (int i = 0; i < 256 && med < 0.0f; ++i) {  // :1587-1591
    bin += (int)std::nearbyintf(hist[i]);        // cvRound
    if ((float)bin > m && med < 0.0f) med = (float)i;
  }
  return med;
}

// Tracker::MedianAbsoluteDeviation, Tracker.cpp:1607-1619
float mad_scale(const float* v, int n) {
  const float c = 1.4826f;
  const float median = median_mat(v, n);
  std::vector<float> dev(n);
  for (int i = 0; i < n; ++i) dev[i] = std::fabs(v[i] - median);  // :1613
  const float MAD = median_mat(dev.data(), n);                    // :1616
  return c * MAD;
}

// Tracker::TukeyFunctionWeights, Tracker.cpp:1626-1654
void tukey_weights(const float* v, int n, float* w) {
  const float b = 4.6851f;
  float MAD = mad_scale(v, n);
  if (MAD == 0) MAD = 1;              // :1634-1637
  const float inv_MAD = 1.0 / MAD;    // :1638 (double division, stored float)
  const float inv_b2 = 1.0 / (b * b); // :1639
  for (int i = 0; i < n; ++i) {
    const float x = v[i] * inv_MAD;   // :1643
    if (std::fabs(x) <= b) {
      const float tukey = (1.0 - (x * x) * inv_b2);  // :1646 (float product, double subtraction)
      w[i] = tukey * tukey;
    } else {
      w[i] = 0.0f;
    }
  }
}

// Huber weights (north-star option, ARITHMETIC.md R4)
void huber_weights(const float* v, int n, float delta, float* w) {
  for (int i = 0; i < n; ++i) {
    const float a = std::fabs(v[i]);
    w[i] = (a <= delta) ? 1.0f : delta / a;
  }
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

void uwo_default_params(uwo_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->width = 640;
  p->height = 480;
  p->fx = 525.0f;  // calibration/calibrationTUM.xml:20
  p->fy = 525.0f;
  p->cx = 319.5f;
  p->cy = 239.5f;
  p->levels = 5;
  p->first_level = 4;
  p->last_level = 1;
  p->max_iterations = 50;
  p->epsilon = 0.001f;
  p->residual_scale = 50.0f;
  p->gradient_threshold = 20.0;
  p->solve_mode = UWO_SOLVE_LU;
  p->accum_mode = UWO_ACCUM_LONGDOUBLE;
  p->threads = 1;
  p->weight_mode = UWO_WEIGHT_IDENTITY;  // Tracker.cpp:495
  p->huber_delta = 10.0f;
  p->lm_lambda = 0.2f;  // the value in the reference's commented DSO-way block, Tracker.cpp:546
  p->sampling = 0;      // Tracker.cpp:472: nearest
}

void uwo_pyr_down(const uint8_t* src, int w, int h, uint8_t* dst) {
  // cv::resize(.., Size(), 0.5, 0.5) with INTER_LINEAR on an exact half scale runs the
  // 2x2 area-average fast path: (a+b+c+d+2)>>2  (verified against cv2 in tests).
  const int w2 = w / 2, h2 = h / 2;
  for (int y = 0; y < h2; ++y) {
    const uint8_t* r0 = src + (size_t)(2 * y) * w;
    const uint8_t* r1 = r0 + w;
    for (int x = 0; x < w2; ++x)
      dst[(size_t)y * w2 + x] =
          (uint8_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
  }
}

void uwo_scharr(const uint8_t* img, int w, int h, int16_t* gx, int16_t* gy) {
  // BORDER_DEFAULT == BORDER_REFLECT_101: index -1 -> 1, index n -> n-2.
  auto refl = [](int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); };
  for (int y = 0; y < h; ++y) {
    const uint8_t* rm = img + (size_t)refl(y - 1, h) * w;
    const uint8_t* r0 = img + (size_t)y * w;
    const uint8_t* rp = img + (size_t)refl(y + 1, h) * w;
    for (int x = 0; x < w; ++x) {
      const int xm = refl(x - 1, w), xp = refl(x + 1, w);
      const int vx = 3 * (rm[xp] - rm[xm]) + 10 * (r0[xp] - r0[xm]) + 3 * (rp[xp] - rp[xm]);
      const int vy = 3 * (rp[xm] - rm[xm]) + 10 * (rp[x] - rm[x]) + 3 * (rp[xp] - rm[xp]);
      gx[(size_t)y * w + x] = (int16_t)vx;
      gy[(size_t)y * w + x] = (int16_t)vy;
    }
  }
}

void uwo_sobel(const uint8_t* img, int w, int h, int16_t* gx, int16_t* gy) {
  auto refl = [](int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); };
  for (int y = 0; y < h; ++y) {
    const uint8_t* rm = img + (size_t)refl(y - 1, h) * w;
    const uint8_t* r0 = img + (size_t)y * w;
    const uint8_t* rp = img + (size_t)refl(y + 1, h) * w;
    for (int x = 0; x < w; ++x) {
      const int xm = refl(x - 1, w), xp = refl(x + 1, w);
      gx[(size_t)y * w + x] =
          (int16_t)((rm[xp] - rm[xm]) + 2 * (r0[xp] - r0[xm]) + (rp[xp] - rp[xm]));
      gy[(size_t)y * w + x] =
          (int16_t)((rp[xm] - rm[xm]) + 2 * (rp[x] - rm[x]) + (rp[xp] - rm[xp]));
    }
  }
}

void uwo_gradmag(const int16_t* gx, const int16_t* gy, long long n, uint8_t* g) {
  for (long long i = 0; i < n; ++i) {
    const int ax = std::min(std::abs((int)gx[i]), 255);  // convertScaleAbs
    const int ay = std::min(std::abs((int)gy[i]), 255);
    const int s = ax + ay;                               // addWeighted(.5,.5): s/2,
    g[i] = (uint8_t)((s + ((s >> 1) & 1)) >> 1);         // ties to even (cvRound)
  }
}

int uwo_candidates(const uint8_t* g, int w, int h, double gradient_threshold, float* pts4,
                   double* mean_out, int* ithr_out) {
  unsigned long long S = 0;
  const long long N = (long long)w * h;
  for (long long i = 0; i < N; ++i) S += g[i];
  const double mean = (double)S / (double)N;            // U6
  const float thres = (float)(mean + gradient_threshold);  // Tracker.cpp:1327 (float thres)
  const int ithr = (int)std::floor(thres);              // 8-bit threshold floors
  if (mean_out) *mean_out = mean;
  if (ithr_out) *ithr_out = ithr;
  int n = 0;
  for (int x = 0; x < w; ++x)          // Tracker.cpp:1334: x outer
    for (int y = 0; y < h; ++y)        // Tracker.cpp:1335: y inner
      if ((int)g[(size_t)y * w + x] > ithr) {
        pts4[(size_t)n * 4 + 0] = (float)x;
        pts4[(size_t)n * 4 + 1] = (float)y;
        pts4[(size_t)n * 4 + 2] = 1.0f;  // depth_initialization, Tracker.cpp:1317,1354
        pts4[(size_t)n * 4 + 3] = 1.0f;
        ++n;
      }
  return n;
}

void uwo_depth_pyr_down(const uint16_t* src, int w, int h, uint16_t* dst) {
  // resizeAreaFast for 16U: float sum * 0.25 -> saturate_cast<ushort>(cvRound(.)): the sum of
  // four 16-bit values and its quarter are exact in float/double, so this is s/4 ties-to-even
  const int w2 = w / 2, h2 = h / 2;
  for (int y = 0; y < h2; ++y) {
    const uint16_t* r0 = src + (size_t)(2 * y) * w;
    const uint16_t* r1 = r0 + w;
    for (int x = 0; x < w2; ++x) {
      const unsigned s = (unsigned)r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1];
      unsigned q = s >> 2;
      const unsigned rem = s & 3u;
      if (rem == 3u || (rem == 2u && (q & 1u))) ++q;
      dst[(size_t)y * w2 + x] = (uint16_t)q;
    }
  }
}

int uwo_candidates_depth(const uint8_t* g, const uint16_t* depth, int w, int h,
                         double gradient_threshold, int depth_mode, float* pts4, uint16_t* zsrc) {
  const float factor = 0.0002;  // Tracker.cpp:1316 "Factor of TUM depth images"
  unsigned long long S = 0;
  const long long N = (long long)w * h;
  for (long long i = 0; i < N; ++i) S += g[i];
  const double mean = (double)S / (double)N;               // U6
  const float thres = (float)(mean + gradient_threshold);  // Tracker.cpp:1327
  const int ithr = (int)std::floor(thres);
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(depth);  // at<uchar> on a CV_16U Mat
  int n = 0;
  for (int x = 0; x < w; ++x)    // Tracker.cpp:1334
    for (int y = 0; y < h; ++y) {  // Tracker.cpp:1335
      // Tracker.cpp:1339,1344: at<uchar>(y,x) = data[y * step + x], step = 2 * w bytes
      const int d = depth_mode == UWO_DEPTH_REFERENCE ? (int)bytes[(size_t)y * 2 * w + x]
                                                      : (int)depth[(size_t)y * w + x];
      if (d != 0 && (int)g[(size_t)y * w + x] > ithr) {
        pts4[(size_t)n * 4 + 0] = (float)x;
        pts4[(size_t)n * 4 + 1] = (float)y;
        pts4[(size_t)n * 4 + 2] = d * factor;  // Tracker.cpp:1344
        pts4[(size_t)n * 4 + 3] = 1.0f;
        if (zsrc) zsrc[n] = (uint16_t)d;
        ++n;
      }
    }
  return n;
}

void uwo_init_pyramid(int w, int h, float fx, float fy, float cx, float cy, int levels, int* wl,
                      int* hl, float* fxl, float* fyl, float* cxl, float* cyl, float* invfxl,
                      float* invfyl) {
  wl[0] = w;
  hl[0] = h;
  fxl[0] = fx;
  fyl[0] = fy;
  cxl[0] = cx;
  cyl[0] = cy;
  invfxl[0] = 1 / fxl[0];
  invfyl[0] = 1 / fyl[0];
  for (int l = 1; l < levels; ++l) {
    wl[l] = w >> l;
    hl[l] = h >> l;
    fxl[l] = fxl[l - 1] * 0.5;                       // float*double -> float (exact)
    fyl[l] = fyl[l - 1] * 0.5;
    cxl[l] = (cxl[0] + 0.5) / ((int)1 << l) - 0.5;   // evaluated in double, stored float
    cyl[l] = (cyl[0] + 0.5) / ((int)1 << l) - 0.5;
    invfxl[l] = 1 / fxl[l];
    invfyl[l] = 1 / fyl[l];
  }
}

void uwo_warp(const float* pts4, int n, const float* pose7, float fx, float fy, float cx,
              float cy, float invfx, float invfy, float* out4) {
  // T = pose.matrix() (se3.hpp:253-268): [R t; 0 1]
  float R[9];
  quat_to_R(pose7, R);
  const double T[16] = {R[0], R[1], R[2], pose7[4], R[3], R[4], R[5], pose7[5],
                        R[6], R[7], R[8], pose7[6], 0.0,  0.0,  0.0,  1.0};
  for (int i = 0; i < n; ++i) {
    const float x = pts4[i * 4 + 0], y = pts4[i * 4 + 1], Z = pts4[i * 4 + 2],
                W = pts4[i * 4 + 3];
    const float X = ((x - cx) * invfx) * Z;  // Tracker.cpp:1439-1440 (U4)
    const float Y = ((y - cy) * invfy) * Z;  // Tracker.cpp:1443-1444
    // rigid * points.t() (Tracker.cpp:1450) is cv::gemm(GEMM_2_T): double accumulators,
    // one rounding to f32 (verified against cv2.gemm, tests/test_oracle_vs_cv2.py).
    // Products of f32-valued doubles are exact, so the association below only moves
    // roundings at the 1e-16 level; this order is the canonical one (ARITHMETIC.md U4).
    float o[4];
    for (int r = 0; r < 4; ++r)
      o[r] = (float)(T[r * 4 + 0] * (double)X +
                     (T[r * 4 + 1] * (double)Y +
                      (T[r * 4 + 2] * (double)Z + T[r * 4 + 3] * (double)W)));
    // Tracker.cpp:1454-1467; cv::divide yields 0 for a zero divisor
    const float qx = (o[2] != 0.0f) ? (o[0] * fx) / o[2] : 0.0f;
    const float qy = (o[2] != 0.0f) ? (o[1] * fy) / o[2] : 0.0f;
    out4[i * 4 + 0] = (qx + cx) * o[3];
    out4[i * 4 + 1] = (qy + cy) * o[3];
    out4[i * 4 + 2] = o[2];
    out4[i * 4 + 3] = o[3];
  }
}

void uwo_se3_exp(const float* tangent6, float* pose7) { pose_to7(se3_exp(tangent6), pose7); }
void uwo_se3_mul(const float* a7, const float* b7, float* out7) {
  pose_to7(se3_mul(pose_from7(a7), pose_from7(b7)), out7);
}
void uwo_se3_matrix(const float* pose7, float* m16) {
  float R[9];
  quat_to_R(pose7, R);
  const float M[16] = {R[0], R[1], R[2], pose7[4], R[3], R[4], R[5], pose7[5],
                       R[6], R[7], R[8], pose7[6], 0.0f, 0.0f, 0.0f, 1.0f};
  std::memcpy(m16, M, sizeof(M));
}
void uwo_se3_scale_level(const float* pose7, float* out7) {
  pose_to7(se3_scale_level(pose_from7(pose7)), out7);
}

void uwo_chain_pose(const float* previous7, const float* rigid7, float scale, float* out7) {
  // Visualizer.cpp:303-309: t_TUM = 40 * translation; SE3(unit_quaternion, t_TUM) normalises
  Pose cur;
  const float len = std::sqrt(quat_sqnorm(rigid7));
  for (int i = 0; i < 4; ++i) cur.q[i] = rigid7[i] / len;
  for (int i = 0; i < 3; ++i) cur.t[i] = scale * rigid7[4 + i];
  pose_to7(se3_mul(pose_from7(previous7), cur), out7);  // Visualizer.cpp:311
}

int uwo_lu_solve6(const float* A36, const float* b6, float* x6) {
  float A[36], B[6];
  std::memcpy(A, A36, sizeof(A));
  std::memcpy(B, b6, sizeof(B));
  if (!lu_impl(A, 6, B, 1)) {  // cv::solve: "if(!result) dst = Scalar(0)"
    for (int i = 0; i < 6; ++i) x6[i] = 0.0f;
    return 0;
  }
  std::memcpy(x6, B, sizeof(B));
  return 1;
}

int uwo_cholesky_lm_solve6(const float* A36, const float* b6, float lambda, float* x6) {
  // ARITHMETIC.md S2: A_ii + lambda * A_ii, L L^T by columns, then L y = b and L^T x = y;
  // every float operation rounded on its own, sums left to right
  float L[36];
  for (int i = 0; i < 6; ++i) x6[i] = 0.0f;
  for (int j = 0; j < 6; ++j) {
    float s = A36[j * 6 + j] + lambda * A36[j * 6 + j];
    for (int k = 0; k < j; ++k) s = s - L[j * 6 + k] * L[j * 6 + k];
    if (!(s > 0.0f)) return 0;
    const float d = std::sqrt(s);
    L[j * 6 + j] = d;
    for (int i = j + 1; i < 6; ++i) {
      float t = A36[i * 6 + j];
      for (int k = 0; k < j; ++k) t = t - L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = t / d;
    }
  }
  float y[6];
  for (int i = 0; i < 6; ++i) {
    float t = b6[i];
    for (int k = 0; k < i; ++k) t = t - L[i * 6 + k] * y[k];
    y[i] = t / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    float t = y[i];
    for (int k = i + 1; k < 6; ++k) t = t - L[k * 6 + i] * x6[k];
    x6[i] = t / L[i * 6 + i];
  }
  return 1;
}

int uwo_lu_invert6(const float* A36, float* Ainv36) {
  float A[36], B[36];
  std::memcpy(A, A36, sizeof(A));
  for (int i = 0; i < 36; ++i) B[i] = (i % 7 == 0) ? 1.0f : 0.0f;
  if (!lu_impl(A, 6, B, 6)) {  // cv::invert: singular -> zero matrix (U8)
    for (int i = 0; i < 36; ++i) Ainv36[i] = 0.0f;
    return 0;
  }
  std::memcpy(Ainv36, B, sizeof(B));
  return 1;
}

// ---- calibration / undistortion front-end (CameraModel.cpp:84-103, System.cpp:148-191, 232-239)
// cv::undistortPoints on one pixel point with P = R = identity, 5 fixed-point iterations
// (OpenCV 4.x cvUndistortPointsInternal with the (ITER, 5) criteria used by
// getOptimalNewCameraMatrix); k = k1 k2 p1 p2, all higher coefficients zero.
static void undistort_point_normalized(double u, double v, double fx, double fy, double cx,
                                       double cy, const double* k, double* xo, double* yo) {
  const double ifx = 1. / fx, ify = 1. / fy;
  double x = (u - cx) * ifx, y = (v - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((0. * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) {
      x = (u - cx) * ifx;
      y = (v - cy) * ify;
      break;
    }
    const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
    const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + 0. * r2 + 0. * r2 * r2;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  *xo = x;
  *yo = y;
}

void uwo_optimal_new_camera_matrix(const float* K9, const float* dist4, int in_w, int in_h,
                                   double alpha, int out_w, int out_h, float* newK9) {
  // getUndistortRectangles: 9 x 9 grid of image points -> normalised undistorted coordinates
  const int N = 9;
  const double k[4] = {dist4[0], dist4[1], dist4[2], dist4[3]};
  double iX0 = -FLT_MAX, iX1 = FLT_MAX, iY0 = -FLT_MAX, iY1 = FLT_MAX;
  double oX0 = FLT_MAX, oX1 = -FLT_MAX, oY0 = FLT_MAX, oY1 = -FLT_MAX;
  for (int y = 0; y < N; ++y)
    for (int x = 0; x < N; ++x) {
      double px, py;
      undistort_point_normalized((double)x * (in_w - 1) / (N - 1), (double)y * (in_h - 1) / (N - 1),
                                 K9[0], K9[4], K9[2], K9[5], k, &px, &py);
      oX0 = std::min(oX0, px); oX1 = std::max(oX1, px);
      oY0 = std::min(oY0, py); oY1 = std::max(oY1, py);
      if (x == 0) iX0 = std::max(iX0, px);
      if (x == N - 1) iX1 = std::min(iX1, px);
      if (y == 0) iY0 = std::max(iY0, py);
      if (y == N - 1) iY1 = std::min(iY1, py);
    }
  const double iw = iX1 - iX0, ih = iY1 - iY0, ow = oX1 - oX0, oh = oY1 - oY0;
  // projections mapping the inner / outer rectangle to the viewport, blended by alpha
  const double fx0 = (out_w - 1) / iw, fy0 = (out_h - 1) / ih;
  const double cx0 = -fx0 * iX0, cy0 = -fy0 * iY0;
  const double fx1 = (out_w - 1) / ow, fy1 = (out_h - 1) / oh;
  const double cx1 = -fx1 * oX0, cy1 = -fy1 * oY0;
  double M[9];
  for (int i = 0; i < 9; ++i) M[i] = K9[i];
  M[0] = fx0 * (1 - alpha) + fx1 * alpha;
  M[4] = fy0 * (1 - alpha) + fy1 * alpha;
  M[2] = cx0 * (1 - alpha) + cx1 * alpha;
  M[5] = cy0 * (1 - alpha) + cy1 * alpha;
  for (int i = 0; i < 9; ++i) newK9[i] = (float)M[i];
}

void uwo_init_undistort_rectify_map(const float* K9, const float* dist4, const float* newK9,
                                    int out_w, int out_h, int16_t* map1, uint16_t* map2) {
  // iR = (newK * I)^-1 : cv::invert of a 3 x 3 CV_64F matrix uses the cofactor formula
  double S[9], ir[9];
  for (int i = 0; i < 9; ++i) S[i] = newK9[i];
  double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
             S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (d != 0.) {
    d = 1. / d;
    ir[0] = (S[4] * S[8] - S[5] * S[7]) * d;
    ir[1] = (S[2] * S[7] - S[1] * S[8]) * d;
    ir[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    ir[3] = (S[5] * S[6] - S[3] * S[8]) * d;
    ir[4] = (S[0] * S[8] - S[2] * S[6]) * d;
    ir[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    ir[6] = (S[3] * S[7] - S[4] * S[6]) * d;
    ir[7] = (S[1] * S[6] - S[0] * S[7]) * d;
    ir[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  } else {
    for (int i = 0; i < 9; ++i) ir[i] = 0.;
  }
  const double u0 = K9[2], v0 = K9[5], fx = K9[0], fy = K9[4];
  const double k1 = dist4[0], k2 = dist4[1], p1 = dist4[2], p2 = dist4[3];
  for (int i = 0; i < out_h; ++i) {
    double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
    for (int j = 0; j < out_w; ++j, _x += ir[0], _y += ir[3], _w += ir[6]) {
      const double w = 1. / _w, x = _x * w, y = _y * w;
      const double x2 = x * x, y2 = y * y;
      const double r2 = x2 + y2, _2xy = 2 * x * y;
      const double kr = (1 + ((0. * r2 + k2) * r2 + k1) * r2) / (1 + ((0. * r2 + 0.) * r2 + 0.) * r2);
      const double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + 0. * r2 + 0. * r2 * r2);
      const double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + 0. * r2 + 0. * r2 * r2);
      const double u = fx * xd + u0, v = fy * yd + v0;
      // saturate_cast<int>(u * INTER_TAB_SIZE): cvRound
      const int iu = (int)std::nearbyint(u * 32), iv = (int)std::nearbyint(v * 32);
      const int sx = iu >> 5, sy = iv >> 5;
      map1[((size_t)i * out_w + j) * 2 + 0] =
          (int16_t)(sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx));
      map1[((size_t)i * out_w + j) * 2 + 1] =
          (int16_t)(sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy));
      map2[(size_t)i * out_w + j] = (uint16_t)((iv & 31) * 32 + (iu & 31));
    }
  }
}

void uwo_remap_bilinear(const uint8_t* src, int sw, int sh, const int16_t* map1,
                        const uint16_t* map2, int dw, int dh, uint8_t* dst) {
  // remapBilinear<FixedPtCast<int, uchar, 15>>: integer weights (32-a)(32-b)*32 out of 2^15,
  // samples outside the image read the constant border value 0
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      const int sx = map1[((size_t)y * dw + x) * 2], sy = map1[((size_t)y * dw + x) * 2 + 1];
      const int f = map2[(size_t)y * dw + x] & 1023;
      const int a = f & 31, b = f >> 5;
      auto S = [&](int yy, int xx) -> int {
        return (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sw + xx] : 0;
      };
      const int v = S(sy, sx) * ((32 - a) * (32 - b) * 32) + S(sy, sx + 1) * (a * (32 - b) * 32) +
                    S(sy + 1, sx) * ((32 - a) * b * 32) + S(sy + 1, sx + 1) * (a * b * 32);
      const int o = (v + (1 << 14)) >> 15;
      dst[(size_t)y * dw + x] = (uint8_t)(o < 0 ? 0 : (o > 255 ? 255 : o));
    }
}

int uwo_calculate_roi(const uint8_t* und, int w, int h, int* roi4) {
  // System.cpp:155-190
  const int x_middle = (int)((w - 1) * 0.5), y_middle = (int)((h - 1) * 0.5);
  int p1x = 0, p1y = 0, p2x = w - 1, p2y = h - 1;
  while (und[(size_t)y_middle * w + p1x] == 0) if (++p1x >= w) return -1;
  while (und[(size_t)y_middle * w + p2x] == 0) if (--p2x < 0) return -1;
  while (und[(size_t)p1y * w + x_middle] == 0) if (++p1y >= h) return -1;
  while (und[(size_t)p2y * w + x_middle] == 0) if (--p2y < 0) return -1;
  p1x += 5; p2x -= 5; p1y += 5; p2y -= 5;  // error margin
  roi4[0] = p1x;
  roi4[1] = p1y;
  roi4[2] = p2x - p1x;  // w_ = p2.x - p1.x, also Rect(p1, p2).width
  roi4[3] = p2y - p1y;
  return 0;
}

float uwo_median_mat(const float* v, int n) { return median_mat(v, n); }
float uwo_mad(const float* v, int n) { return mad_scale(v, n); }
void uwo_tukey_weights(const float* v, int n, float* w) { tukey_weights(v, n, w); }
void uwo_huber_weights(const float* v, int n, float delta, float* w) { huber_weights(v, n, delta, w); }

void uwo_build_pyramid(const uwo_params* p, uint8_t* const* images) {
  for (int l = 1; l < p->levels; ++l)
    uwo_pyr_down(images[l - 1], p->width >> (l - 1), p->height >> (l - 1), images[l]);
}

// One residual sweep (Tracker.cpp:422-490 + the sums of :559-562) over candidate rows
// [lo, hi) of one level at pose7.  sums32: 21 upper-triangular J^T J terms, 6 J^T (50 r)
// terms (b is minus these), sum r^2, N_valid, [29] = sum r (r w) when weighted, 2 x 0.  The structure is the reference's: a
// separate WarpFunction pass, then materialised Jacobian / residual arrays, then the products.
int uwo_sweep_range(const uwo_params* p, int lvl, const uint8_t* I1, const uint8_t* I2,
                    const int16_t* gx, const int16_t* gy, const float* cand, int lo, int hi,
                    const float* pose7, double* sums32) {
  int wl[UWO_MAX_LEVELS], hl[UWO_MAX_LEVELS];
  float fxl[UWO_MAX_LEVELS], fyl[UWO_MAX_LEVELS], cxl[UWO_MAX_LEVELS], cyl[UWO_MAX_LEVELS],
      ifx[UWO_MAX_LEVELS], ify[UWO_MAX_LEVELS];
  if (p->levels > UWO_MAX_LEVELS || lvl < 0 || lvl >= p->levels || lo < 0 || hi < lo) return -1;
  uwo_init_pyramid(p->width, p->height, p->fx, p->fy, p->cx, p->cy, p->levels, wl, hl, fxl, fyl,
                   cxl, cyl, ifx, ify);
  const int n = hi - lo;
  const float* pts = cand + (size_t)lo * 4;
  const int cols = wl[lvl], rows = hl[lvl];
  const float fx = fxl[lvl], fy = fyl[lvl];
  const int threads = p->threads < 1 ? 1 : p->threads;
  std::vector<float> warped((size_t)n * 4), J((size_t)n * 6), r50(n), res(n);
  std::vector<unsigned char> valid(n);
  // --- WarpFunction (separate pass over all points, Tracker.cpp:422) ---
  parallel_chunks(threads, [&](int c) {
    const int a = (int)((long long)n * c / threads), b = (int)((long long)n * (c + 1) / threads);
    uwo_warp(pts + (size_t)a * 4, b - a, pose7, fx, fy, cxl[lvl], cyl[lvl], ifx[lvl], ify[lvl],
             warped.data() + (size_t)a * 4);
  });
  // --- residuals and Jacobian rows (Tracker.cpp:432-490) ---
  std::vector<long long> part_r2(threads, 0);
  std::vector<int> part_nv(threads, 0);
  parallel_chunks(threads, [&](int c) {
    const int a = (int)((long long)n * c / threads), b = (int)((long long)n * (c + 1) / threads);
    long long sum_r2 = 0;
    int n_valid = 0;
    for (int i = a; i < b; ++i) {
      const float x1 = pts[(size_t)i * 4 + 0], y1 = pts[(size_t)i * 4 + 1];
      const float x2 = warped[(size_t)i * 4 + 0], y2 = warped[(size_t)i * 4 + 1],
                  z2 = warped[(size_t)i * 4 + 2];
      valid[i] = 0;
      if (y2 > 0 && y2 < rows && x2 > 0 && x2 < cols && z2 != 0) {  // Tracker.cpp:450-451
        float inv_z2 = 1 / z2;                                      // Tracker.cpp:447
        if (inv_z2 < 0) inv_z2 = 0;                                 // Tracker.cpp:452-453
        float Jw0[6], Jw1[6];                                       // Tracker.cpp:455-467
        Jw0[0] = fx * inv_z2;
        Jw0[1] = 0.0f;
        Jw0[2] = -(fx * x2 * inv_z2 * inv_z2);
        Jw0[3] = -(fx * x2 * y2 * inv_z2 * inv_z2);
        Jw0[4] = (fx * (1 + x2 * x2 * inv_z2 * inv_z2));
        Jw0[5] = -fx * y2 * inv_z2;
        Jw1[0] = 0.0f;
        Jw1[1] = fy * inv_z2;
        Jw1[2] = -(fy * y2 * inv_z2 * inv_z2);
        Jw1[3] = -(fy * (1 + y2 * y2 * inv_z2 * inv_z2));
        Jw1[4] = fy * x2 * y2 * inv_z2 * inv_z2;
        Jw1[5] = fy * x2 * inv_z2;
        // nearest sample at round-half-away, clamped to the image (U1)
        int xi = iround_half_away(x2), yi = iround_half_away(y2);
        if (xi > cols - 1) xi = cols - 1;
        if (yi > rows - 1) yi = rows - 1;
        const int i1 = I1[(size_t)(int)y1 * cols + (int)x1];  // Tracker.cpp:471
        const int i2 = I2[(size_t)yi * cols + xi];            // Tracker.cpp:472
        const int r = i2 - i1;                                // Tracker.cpp:474
        float rf = (float)r;
        if (p->sampling == 1) {
          // north-star option (ARITHMETIC.md B1): float bilinear interpolation at (x2, y2)
          const int ix = (int)x2, iy = (int)y2;
          const float ax = x2 - (float)ix, ay = y2 - (float)iy;
          const int ix1 = ix + 1 > cols - 1 ? cols - 1 : ix + 1;
          const int iy1 = iy + 1 > rows - 1 ? rows - 1 : iy + 1;
          const float a = I2[(size_t)iy * cols + ix], b = I2[(size_t)iy * cols + ix1];
          const float c = I2[(size_t)iy1 * cols + ix], d = I2[(size_t)iy1 * cols + ix1];
          const float top = a + ax * (b - a), bot = c + ax * (d - c);
          const float v = top + ay * (bot - top);
          rf = v - (float)i1;
        }
        const float jlx = gx[(size_t)(int)y1 * cols + (int)x1];  // Tracker.cpp:476
        const float jly = gy[(size_t)(int)y1 * cols + (int)x1];  // Tracker.cpp:477
        // Jl * Jw (Tracker.cpp:479) is cv::gemm: double accumulators, one rounding
        for (int q = 0; q < 6; ++q)
          J[(size_t)i * 6 + q] =
              (float)((double)jlx * (double)Jw0[q] + (double)jly * (double)Jw1[q]);
        res[i] = rf;                         // Residuals row, Tracker.cpp:487
        r50[i] = rf * p->residual_scale;     // Tracker.cpp:559 (exact for integer residuals)
        valid[i] = 1;
        if (p->sampling == 0) sum_r2 += (long long)r * r;
        ++n_valid;
      }
    }
    part_r2[c] = sum_r2;
    part_nv[c] = n_valid;
  });
  long long sum_r2 = 0;
  int n_valid = 0;
  for (int c = 0; c < threads; ++c) {
    sum_r2 += part_r2[c];
    n_valid += part_nv[c];
  }
  for (int i = 0; i < 32; ++i) sums32[i] = 0.0;
  if (p->sampling == 1 && p->weight_mode == UWO_WEIGHT_IDENTITY) {
    // float residuals: error = inv_num * Residuals^T Residuals with the fp64 accumulation of U3
    long double esum = 0.0L;
    for (int i = 0; i < n; ++i)
      if (valid[i]) esum += (long double)((double)res[i] * (double)res[i]);
    sums32[29] = (double)esum;
  }
  if (p->weight_mode != UWO_WEIGHT_IDENTITY) {
    // Tracker.cpp:496: W = TukeyFunctionWeights(Residuals) on the valid rows only, in order
    std::vector<float> rv, wv;
    rv.reserve(n_valid);
    for (int i = 0; i < n; ++i)
      if (valid[i]) rv.push_back(res[i]);
    wv.resize(rv.size());
    if (p->weight_mode == UWO_WEIGHT_TUKEY)
      tukey_weights(rv.data(), (int)rv.size(), wv.data());
    else
      huber_weights(rv.data(), (int)rv.size(), p->huber_delta, wv.data());
    long double esum = 0.0L;
    size_t j = 0;
    for (int i = 0; i < n; ++i) {
      if (!valid[i]) continue;
      const float w = wv[j++];
      // Tracker.cpp:500-501: error = inv_num * Residuals^T (Residuals .* W)   (U3: fp64 sum)
      const float rw = res[i] * w;
      esum += (long double)((double)res[i] * (double)rw);
      // Tukey (Kerl-way, Tracker.cpp:554-562): rows of J and the scaled residuals are both
      // multiplied by w;  Huber: by sqrt(w), i.e. A = sum w J J^T, b = -sum w J (50 r)
      const float s = (p->weight_mode == UWO_WEIGHT_TUKEY) ? w : std::sqrt(w);
      for (int q = 0; q < 6; ++q) J[(size_t)i * 6 + q] = s * J[(size_t)i * 6 + q];
      r50[i] = r50[i] * s;
    }
    sums32[29] = (double)esum;
  }
  // Tracker.cpp:559-562: A = J^T J, b = -J^T (50 r)
  if (p->accum_mode == UWO_ACCUM_LONGDOUBLE)
    accumulate<long double>(J.data(), r50.data(), valid.data(), n, threads, sums32, sums32 + 21);
  else
    accumulate<double>(J.data(), r50.data(), valid.data(), n, threads, sums32, sums32 + 21);
  sums32[27] = (double)sum_r2;
  sums32[28] = (double)n_valid;
  return 0;
}

// Tracker.cpp:495-574 on the (reduced) sums of one sweep: break test, solve, pose update.
// pose7 / last_error are updated in place.  Returns 1 when the level is finished, else 0.
int uwo_gn_update(const uwo_params* p, const double* sums32, int k, float* pose7,
                  float* last_error, uwo_iter_trace* tr) {
  const long long sum_r2 = (long long)sums32[27];
  const int n_valid = (int)sums32[28];
  if (tr) {
    tr->n_valid = n_valid;
    tr->sum_r2 = sum_r2;
  }
  if (n_valid == 0) {  // U2: nothing to optimise on this level
    if (tr) tr->broke = 1;
    return 1;
  }
  // Tracker.cpp:499-502: error = (1/N) r^T r  (U3)
  const float inv_num = 1.0 / n_valid;
  const float error = (p->weight_mode == UWO_WEIGHT_IDENTITY && p->sampling == 0)
                          ? (float)((double)inv_num * (double)sum_r2)
                          : (float)((double)inv_num * sums32[29]);
  if (tr) tr->error = error;
  // Tracker.cpp:508: break test (the update that led here is kept)
  if (error >= *last_error || k == p->max_iterations - 1 ||
      std::fabs(error - *last_error) < p->epsilon) {
    if (tr) tr->broke = 1;
    return 1;
  }
  *last_error = error;  // Tracker.cpp:529
  float A[36], b[6], delta[6];
  {
    int idx = 0;
    for (int a = 0; a < 6; ++a)
      for (int c = a; c < 6; ++c) {
        A[a * 6 + c] = A[c * 6 + a] = (float)sums32[idx];
        ++idx;
      }
    for (int a = 0; a < 6; ++a) b[a] = (float)(-sums32[21 + a]);
  }
  // Tracker.cpp:564: deltaMat = A.inv() * b
  if (p->solve_mode == UWO_SOLVE_CHOLESKY_LM) {
    if (!uwo_cholesky_lm_solve6(A, b, p->lm_lambda, delta))
      for (int i = 0; i < 6; ++i) delta[i] = 0.0f;
  } else if (p->solve_mode == UWO_SOLVE_LU) {
    uwo_lu_solve6(A, b, delta);
  } else {
    float Ainv[36];
    uwo_lu_invert6(A, Ainv);
    for (int a = 0; a < 6; ++a) {
      double s = 0.0;
      for (int c = 0; c < 6; ++c) s += (double)Ainv[a * 6 + c] * (double)b[c];
      delta[a] = (float)s;
    }
  }
  // Tracker.cpp:574: current_pose = current_pose * SE3::exp(delta)
  pose_to7(se3_mul(pose_from7(pose7), se3_exp(delta)), pose7);
  if (tr) {
    std::memcpy(tr->A, A, sizeof(A));
    std::memcpy(tr->b, b, sizeof(b));
    std::memcpy(tr->delta, delta, sizeof(delta));
  }
  return 0;
}

int uwo_estimate_pose(const uwo_params* p, const uint8_t* const* prev_images,
                      const uint8_t* const* cur_images, const int16_t* const* gxs,
                      const int16_t* const* gys, const float* const* cand, const int* ncand,
                      const float* init_pose7, float* out_pose7, uwo_stats* stats,
                      uwo_iter_trace* trace, int trace_cap, int* n_trace) {
  const int L = p->levels;
  if (L > UWO_MAX_LEVELS || p->first_level >= L || p->last_level < 0) return -1;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  int nt = 0;
  // Tracker.cpp:385: identity start (init_pose7 is an extension; NULL = reference)
  float pose7[7];
  if (init_pose7) {
    std::memcpy(pose7, init_pose7, sizeof(pose7));
  } else {
    const float zero6[6] = {0, 0, 0, 0, 0, 0};
    pose_to7(se3_exp(zero6), pose7);
  }
  for (int lvl = p->first_level; lvl >= p->last_level; --lvl) {  // Tracker.cpp:389
    float last_error = 50000.0f;                                  // Tracker.cpp:393
    if (stats) stats->n_points[lvl] = ncand[lvl];
    for (int k = 0; k < p->max_iterations; ++k) {  // Tracker.cpp:414
      double sums[32];
      int rc = uwo_sweep_range(p, lvl, prev_images[lvl], cur_images[lvl], gxs[lvl], gys[lvl],
                               cand[lvl], 0, ncand[lvl], pose7, sums);
      if (rc) return rc;
      if (stats) stats->evaluations[lvl] = k + 1;
      uwo_iter_trace* tr = (trace && nt < trace_cap) ? &trace[nt] : nullptr;
      if (tr) {
        std::memset(tr, 0, sizeof(*tr));
        tr->level = lvl;
        tr->k = k;
      }
      const float err_before = last_error;
      const int brk = uwo_gn_update(p, sums, k, pose7, &last_error, tr);
      if (tr) {
        std::memcpy(tr->pose, pose7, sizeof(pose7));
        ++nt;
      }
      if (stats && sums[28] > 0) {
        const float inv_num = 1.0 / (int)sums[28];
        stats->final_error[lvl] =
            (p->weight_mode == UWO_WEIGHT_IDENTITY && p->sampling == 0)
                ? (float)((double)inv_num * (double)(long long)sums[27])
                : (float)((double)inv_num * sums[29]);
        if (!brk) stats->iterations[lvl] = k + 1;
      }
      (void)err_before;
      if (brk) break;
    }
    if (lvl != 0) pose_to7(se3_scale_level(pose_from7(pose7)), pose7);  // Tracker.cpp:580-590
  }
  std::memcpy(out_pose7, pose7, sizeof(pose7));  // Tracker.cpp:595
  if (n_trace) *n_trace = nt;
  return 0;
}

int uwo_track_pair(const uwo_params* p, const uint8_t* prev0, const uint8_t* cur0,
                   float* out_pose7, uwo_stats* stats, double* seconds4) {
  const int L = p->levels;
  std::vector<std::vector<uint8_t>> pi(L), ci(L), g(L);
  std::vector<std::vector<int16_t>> gx(L), gy(L);
  std::vector<std::vector<float>> cand(L);
  std::vector<int> ncand(L, 0);
  const uint8_t* pptr[UWO_MAX_LEVELS];
  const uint8_t* cptr[UWO_MAX_LEVELS];
  const int16_t* gxp[UWO_MAX_LEVELS];
  const int16_t* gyp[UWO_MAX_LEVELS];
  const float* candp[UWO_MAX_LEVELS];
  double t0 = now_s();
  for (int l = 0; l < L; ++l) {
    const size_t n = (size_t)(p->width >> l) * (p->height >> l);
    pi[l].resize(n);
    ci[l].resize(n);
    if (l == 0) {
      std::memcpy(pi[0].data(), prev0, n);
      std::memcpy(ci[0].data(), cur0, n);
    } else {
      uwo_pyr_down(pi[l - 1].data(), p->width >> (l - 1), p->height >> (l - 1), pi[l].data());
      uwo_pyr_down(ci[l - 1].data(), p->width >> (l - 1), p->height >> (l - 1), ci[l].data());
    }
    pptr[l] = pi[l].data();
    cptr[l] = ci[l].data();
  }
  double t1 = now_s();
  for (int l = 0; l < L; ++l) {  // ApplyGradient(prev): all levels, Tracker.cpp:1129
    const int w = p->width >> l, h = p->height >> l;
    gx[l].resize((size_t)w * h);
    gy[l].resize((size_t)w * h);
    g[l].resize((size_t)w * h);
    uwo_scharr(pi[l].data(), w, h, gx[l].data(), gy[l].data());
    uwo_gradmag(gx[l].data(), gy[l].data(), (long long)w * h, g[l].data());
    gxp[l] = gx[l].data();
    gyp[l] = gy[l].data();
  }
  double t2 = now_s();
  for (int l = 0; l < L; ++l) {  // ObtainCandidatePoints(prev): all levels, Tracker.cpp:1319
    const int w = p->width >> l, h = p->height >> l;
    cand[l].resize((size_t)w * h * 4);
    ncand[l] = uwo_candidates(g[l].data(), w, h, p->gradient_threshold, cand[l].data(), nullptr,
                              nullptr);
    candp[l] = cand[l].data();
  }
  double t3 = now_s();
  int rc = uwo_estimate_pose(p, pptr, cptr, gxp, gyp, candp, ncand.data(), nullptr, out_pose7,
                             stats, nullptr, 0, nullptr);
  double t4 = now_s();
  if (seconds4) {
    seconds4[0] = t1 - t0;
    seconds4[1] = t2 - t1;
    seconds4[2] = t3 - t2;
    seconds4[3] = t4 - t3;
  }
  return rc;
}


// One reference-shaped tracking loop over a frame sequence (System::AddFrame +
// System::Tracking in the direct order), as a stateful stream: the object holds
// what `System` keeps between frames (the previous uw::Frame with its pyramid, gradients and
// candidate points).  uwo_stream_create prepares frame 0; every uwo_stream_track costs exactly
// one "track": pyramid(i), EstimatePose(i-1, i), ApplyGradient(i), ObtainCandidatePoints(i)
// (main_uw_slam.cpp:139-151 / System.cpp:193-223).
}  // extern "C"

namespace {
struct StreamFrame {
  std::vector<std::vector<uint8_t>> img, g;
  std::vector<std::vector<int16_t>> gx, gy;
  std::vector<std::vector<float>> cand;
  std::vector<int> ncand;
};
void stream_pyramid(const uwo_params& p, StreamFrame& f, const uint8_t* l0) {
  const int L = p.levels;
  f.img.resize(L);
  for (int l = 0; l < L; ++l) {
    const size_t n = (size_t)(p.width >> l) * (p.height >> l);
    f.img[l].resize(n);
    if (l == 0)
      std::memcpy(f.img[0].data(), l0, n);
    else
      uwo_pyr_down(f.img[l - 1].data(), p.width >> (l - 1), p.height >> (l - 1), f.img[l].data());
  }
}
void stream_gradient_candidates(const uwo_params& p, StreamFrame& f) {
  const int L = p.levels;
  f.g.resize(L); f.gx.resize(L); f.gy.resize(L); f.cand.resize(L); f.ncand.assign(L, 0);
  for (int l = 0; l < L; ++l) {
    const int w = p.width >> l, h = p.height >> l;
    f.gx[l].resize((size_t)w * h);
    f.gy[l].resize((size_t)w * h);
    f.g[l].resize((size_t)w * h);
    uwo_scharr(f.img[l].data(), w, h, f.gx[l].data(), f.gy[l].data());
    uwo_gradmag(f.gx[l].data(), f.gy[l].data(), (long long)w * h, f.g[l].data());
  }
  for (int l = 0; l < L; ++l) {
    const int w = p.width >> l, h = p.height >> l;
    f.cand[l].resize((size_t)w * h * 4);
    f.ncand[l] = uwo_candidates(f.g[l].data(), w, h, p.gradient_threshold, f.cand[l].data(),
                                nullptr, nullptr);
  }
}
}  // namespace

struct uwo_stream {
  uwo_params p;
  StreamFrame a, b;
  StreamFrame* prev;
  StreamFrame* cur;
};

extern "C" {

uwo_stream* uwo_stream_create(const uwo_params* p, const uint8_t* frame0) {
  uwo_stream* s = new (std::nothrow) uwo_stream();
  if (!s) return nullptr;
  s->p = *p;
  s->prev = &s->a;
  s->cur = &s->b;
  stream_pyramid(s->p, *s->prev, frame0);
  stream_gradient_candidates(s->p, *s->prev);
  return s;
}

void uwo_stream_destroy(uwo_stream* s) { delete s; }

int uwo_stream_track(uwo_stream* s, const uint8_t* frame, float* out_pose7, uwo_stats* stats) {
  const uwo_params* p = &s->p;
  const int L = p->levels;
  stream_pyramid(*p, *s->cur, frame);
  const uint8_t* pp[UWO_MAX_LEVELS];
  const uint8_t* cp[UWO_MAX_LEVELS];
  const int16_t* gxp[UWO_MAX_LEVELS];
  const int16_t* gyp[UWO_MAX_LEVELS];
  const float* cd[UWO_MAX_LEVELS];
  for (int l = 0; l < L; ++l) {
    pp[l] = s->prev->img[l].data();
    cp[l] = s->cur->img[l].data();
    gxp[l] = s->prev->gx[l].data();
    gyp[l] = s->prev->gy[l].data();
    cd[l] = s->prev->cand[l].data();
  }
  int rc = uwo_estimate_pose(p, pp, cp, gxp, gyp, cd, s->prev->ncand.data(), nullptr, out_pose7,
                             stats, nullptr, 0, nullptr);
  if (rc) return rc;
  stream_gradient_candidates(*p, *s->cur);
  std::swap(s->prev, s->cur);
  return 0;
}

// The same loop over a whole sequence held in memory.  poses_out: (n_frames-1) x 7;
// seconds_out (optional): per-track wall seconds, (n_frames-1) entries (frame 0's preparation
// is outside every one of them).
int uwo_track_sequence(const uwo_params* p, const uint8_t* frames, int n_frames,
                       float* poses_out, double* seconds_out, uwo_stats* stats_out) {
  const size_t fsz = (size_t)p->width * p->height;
  uwo_stream* s = uwo_stream_create(p, frames);
  if (!s) return -1;
  int rc = 0;
  for (int i = 1; i < n_frames && rc == 0; ++i) {
    const double t0 = now_s();
    rc = uwo_stream_track(s, frames + (size_t)i * fsz, poses_out + (size_t)(i - 1) * 7,
                          stats_out ? stats_out + (i - 1) : nullptr);
    if (seconds_out) seconds_out[i - 1] = now_s() - t0;
  }
  uwo_stream_destroy(s);
  return rc;
}

}  // extern "C"
