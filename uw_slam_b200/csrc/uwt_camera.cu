// uwt_camera.cu -- calibration / undistortion front-end.
//
// The reference prepares rectification once on the host and applies it to every frame:
//   CameraModel::GetCameraModel, rectify branch   /root/reference/src/CameraModel.cpp:84-98
//       getOptimalNewCameraMatrix(K, dist, in_size, 1.0, out_size, nullptr, false)
//       initUndistortRectifyMap(K, dist, Mat(), newK, out_size, CV_16SC2, map1, map2)
//   CameraModel::Undistort / System::AddFrame     CameraModel.cpp:101-103, System.cpp:232-235
//       remap(img, out, map1, map2, INTER_LINEAR);  out(ROI)
//   System::CalculateROI                          System.cpp:148-191
// The two one-off OpenCV calls are host code here as well (double precision, a few hundred
// thousand evaluations), written after OpenCV 4.x -- the release whose results can be checked
// in this environment; the per-frame remap + crop runs on the GPU, fused into the pyramid
// kernel (uwt_image_kernels.cu) or stand-alone (uwt_undistort_image, used for the ROI search).
#include <cfloat>
#include <cmath>
#include <cstring>

#include "uwt_internal.cuh"

namespace {

struct Brown4 {  // k1 k2 p1 p2: the four coefficients of calibration/*.xml <rectification>
  double k1, k2, p1, p2;
};

// Normalised camera coordinates of pixel (u, v): inverse of the distortion model by the
// fixed-point iteration of cv::undistortPoints (5 iterations, no early exit).
inline void normalise_undistorted(double u, double v, const double K[4], const Brown4& c,
                                  double& xn, double& yn) {
  const double ifx = 1. / K[0], ify = 1. / K[1];
  const double xd = (u - K[2]) * ifx, yd = (v - K[3]) * ify;
  double x = xd, y = yd;
  for (int it = 0; it < 5; ++it) {
    const double r2 = x * x + y * y;
    // numerator and denominator of OpenCV's rational model with k3..k6 = 0, evaluated in the
    // same Horner form so that the roundings agree
    const double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((0. * r2 + c.k2) * r2 + c.k1) * r2);
    if (icdist < 0) {  // the model folds over: keep the distorted coordinates
      x = xd;
      y = yd;
      break;
    }
    const double dx = 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
    const double dy = c.p1 * (r2 + 2 * y * y) + 2 * c.p2 * x * y + 0. * r2 + 0. * r2 * r2;
    x = (xd - dx) * icdist;
    y = (yd - dy) * icdist;
  }
  xn = x;
  yn = y;
}

struct Box {
  double x0, y0, x1, y1;
};

}  // namespace

extern "C" {

int uwt_camera_optimal_matrix(const float* K9, const float* dist4, int in_w, int in_h,
                              double alpha, int out_w, int out_h, float* newK9) {
  if (!K9 || !dist4 || !newK9 || in_w < 2 || in_h < 2 || out_w < 2 || out_h < 2)
    return UWT_E_INVALID;
  const double K[4] = {K9[0], K9[4], K9[2], K9[5]};
  const Brown4 c = {dist4[0], dist4[1], dist4[2], dist4[3]};
  // image border sampled on a 9 x 9 grid -> where it lands in normalised coordinates:
  // `outer` bounds all of it, `inner` is the largest axis-aligned box inside
  constexpr int N = 9;
  Box outer = {FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX}, inner = {-FLT_MAX, -FLT_MAX, FLT_MAX, FLT_MAX};
  for (int gy = 0; gy < N; ++gy)
    for (int gx = 0; gx < N; ++gx) {
      double x, y;
      normalise_undistorted((double)gx * (in_w - 1) / (N - 1), (double)gy * (in_h - 1) / (N - 1), K,
                            c, x, y);
      outer.x0 = fmin(outer.x0, x); outer.x1 = fmax(outer.x1, x);
      outer.y0 = fmin(outer.y0, y); outer.y1 = fmax(outer.y1, y);
      if (gx == 0) inner.x0 = fmax(inner.x0, x);
      if (gx == N - 1) inner.x1 = fmin(inner.x1, x);
      if (gy == 0) inner.y0 = fmax(inner.y0, y);
      if (gy == N - 1) inner.y1 = fmin(inner.y1, y);
    }
  // projection that maps a box onto the output viewport
  auto fit = [&](const Box& b, double out[4]) {
    out[0] = (out_w - 1) / (b.x1 - b.x0);
    out[1] = (out_h - 1) / (b.y1 - b.y0);
    out[2] = -out[0] * b.x0;
    out[3] = -out[1] * b.y0;
  };
  double pi[4], po[4];
  fit(inner, pi);
  fit(outer, po);
  for (int i = 0; i < 9; ++i) newK9[i] = K9[i];
  newK9[0] = (float)(pi[0] * (1 - alpha) + po[0] * alpha);
  newK9[4] = (float)(pi[1] * (1 - alpha) + po[1] * alpha);
  newK9[2] = (float)(pi[2] * (1 - alpha) + po[2] * alpha);
  newK9[5] = (float)(pi[3] * (1 - alpha) + po[3] * alpha);
  return UWT_OK;
}

int uwt_camera_undistort_maps(const float* K9, const float* dist4, const float* newK9, int out_w,
                              int out_h, int16_t* map1, uint16_t* map2) {
  if (!K9 || !dist4 || !newK9 || !map1 || !map2 || out_w < 1 || out_h < 1) return UWT_E_INVALID;
  // inverse of the new camera matrix by cofactors (what cv::invert does for 3 x 3 doubles)
  double m[9], inv[9];
  for (int i = 0; i < 9; ++i) m[i] = newK9[i];
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[3] * m[8] - m[5] * m[6],
               c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 - m[1] * c01 + m[2] * c02;
  if (det == 0.) return UWT_E_INVALID;
  det = 1. / det;
  inv[0] = c00 * det;
  inv[1] = (m[2] * m[7] - m[1] * m[8]) * det;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) * det;
  inv[3] = (m[5] * m[6] - m[3] * m[8]) * det;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) * det;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) * det;
  inv[6] = c02 * det;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) * det;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) * det;
  const double fx = K9[0], fy = K9[4], u0 = K9[2], v0 = K9[5];
  const Brown4 c = {dist4[0], dist4[1], dist4[2], dist4[3]};
  auto clamp16 = [](int v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); };
  for (int row = 0; row < out_h; ++row) {
    // homogeneous ray of the first pixel of the row, then one increment per column (the
    // running sums are part of the result: they decide the last bit of the fixed-point maps)
    double hx = row * inv[1] + inv[2], hy = row * inv[4] + inv[5], hw = row * inv[7] + inv[8];
    int16_t* o1 = map1 + (size_t)row * out_w * 2;
    uint16_t* o2 = map2 + (size_t)row * out_w;
    for (int col = 0; col < out_w; ++col, hx += inv[0], hy += inv[3], hw += inv[6]) {
      const double w = 1. / hw, x = hx * w, y = hy * w;
      const double x2 = x * x, y2 = y * y, r2 = x2 + y2, xy2 = 2 * x * y;
      const double kr = (1 + ((0. * r2 + c.k2) * r2 + c.k1) * r2) / (1 + ((0. * r2 + 0.) * r2 + 0.) * r2);
      const double xd = (x * kr + c.p1 * xy2 + c.p2 * (r2 + 2 * x2) + 0. * r2 + 0. * r2 * r2);
      const double yd = (y * kr + c.p1 * (r2 + 2 * y2) + c.p2 * xy2 + 0. * r2 + 0. * r2 * r2);
      // source position in 1/32 pixel units, ties to even
      const int iu = (int)nearbyint((fx * xd + u0) * 32), iv = (int)nearbyint((fy * yd + v0) * 32);
      o1[2 * col] = clamp16(iu >> 5);
      o1[2 * col + 1] = clamp16(iv >> 5);
      o2[col] = (uint16_t)(((iv & 31) << 5) | (iu & 31));
    }
  }
  return UWT_OK;
}

int uwt_calculate_roi(const uint8_t* und, int w, int h, size_t stride, int* roi4) {
  if (!und || !roi4 || w < 1 || h < 1 || stride < (size_t)w) return UWT_E_INVALID;
  // System.cpp:155-190: walk inwards from the four border mid-points while the pixel is black
  const int xm = (int)((w - 1) * 0.5), ym = (int)((h - 1) * 0.5);
  int left = 0, right = w - 1, top = 0, bottom = h - 1;
  const uint8_t* mid_row = und + (size_t)ym * stride;
  while (left < w && mid_row[left] == 0) ++left;
  while (right >= 0 && mid_row[right] == 0) --right;
  while (top < h && und[(size_t)top * stride + xm] == 0) ++top;
  while (bottom >= 0 && und[(size_t)bottom * stride + xm] == 0) --bottom;
  if (left >= w || right < 0 || top >= h || bottom < 0) return UWT_E_INVALID;  // all black
  roi4[0] = left + 5;                          // error margin, System.cpp:179-183
  roi4[1] = top + 5;
  roi4[2] = (right - 5) - (left + 5);          // w_ = p2.x - p1.x, System.cpp:188
  roi4[3] = (bottom - 5) - (top + 5);
  return (roi4[2] > 0 && roi4[3] > 0) ? UWT_OK : UWT_E_INVALID;
}

int uwt_undistort_image(int device, const uint8_t* src, int in_w, int in_h, size_t src_stride,
                        const int16_t* map1, const uint16_t* map2, int out_w, int out_h,
                        uint8_t* dst) {
  if (!src || !map1 || !map2 || !dst || in_w < 1 || in_h < 1 || out_w < 1 || out_h < 1 ||
      src_stride < (size_t)in_w)
    return UWT_E_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return UWT_E_CUDA;
  const size_t n_out = (size_t)out_w * out_h;
  uint8_t *d_src = nullptr, *d_dst = nullptr;
  short2* d_m1 = nullptr;
  uint16_t* d_m2 = nullptr;
  cudaStream_t st = nullptr;
  int rc = UWT_E_CUDA;
  do {
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) break;
    if (cudaMalloc(&d_src, (size_t)in_w * in_h) != cudaSuccess) break;
    if (cudaMalloc(&d_dst, n_out) != cudaSuccess) break;
    if (cudaMalloc(&d_m1, n_out * sizeof(short2)) != cudaSuccess) break;
    if (cudaMalloc(&d_m2, n_out * sizeof(uint16_t)) != cudaSuccess) break;
    if (cudaMemcpy2DAsync(d_src, in_w, src, src_stride, in_w, in_h, cudaMemcpyHostToDevice, st) !=
        cudaSuccess)
      break;
    cudaMemcpyAsync(d_m1, map1, n_out * sizeof(short2), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_m2, map2, n_out * sizeof(uint16_t), cudaMemcpyHostToDevice, st);
    if (uwt::launch_remap(d_src, in_w, in_w, in_h, d_m1, d_m2, out_w, out_h, d_dst, st) < 0) break;
    cudaMemcpyAsync(dst, d_dst, n_out, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) break;
    rc = UWT_OK;
  } while (0);
  cudaFree(d_src);
  cudaFree(d_dst);
  cudaFree(d_m1);
  cudaFree(d_m2);
  if (st) cudaStreamDestroy(st);
  return rc;
}

}  // extern "C"
