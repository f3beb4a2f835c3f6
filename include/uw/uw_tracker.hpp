// uw_tracker.hpp -- header-only C++ facade over the C ABI (include/uwtrack.h) with the
// reference's class surface, so code written against uw::Tracker / uw::CameraModel /
// uw::Frame ports by search-and-replace.  Plain buffers replace cv::Mat (OpenCV is not a
// dependency of this library).
//
//   reference                                         here
//   uw::CameraModel   include/CameraModel.h:42-145    uw::CameraModel (pinhole + rectify branch)
//   System::CalculateROI src/System.cpp:148-191       uw::CalculateROI / uw::AlignROI
//   uw::Frame         include/System.h:63-103         uw::Frame (a device-side frame slot)
//   uw::Tracker       include/Tracker.h:90-531        uw::Tracker (direct photometric path)
//   SE3               include/Options.h (Sophus::SE3f) uw::SE3f (7 floats, Sophus order)
//
// Errors: the reference exit(0)s; here every failure throws uw::Error (code + message).
#pragma once
#include <array>
#include <cstdint>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../uwtrack.h"

namespace uw {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// Sophus::SE3f storage: unit quaternion (x, y, z, w) then translation (se3.hpp:469-472).
struct SE3f {
  std::array<float, 7> data{{0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f}};
  // 4x4 row-major homogeneous matrix (se3.hpp:253-268, Eigen toRotationMatrix).
  std::array<float, 16> matrix() const {
    const float x = data[0], y = data[1], z = data[2], w = data[3];
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    return {{1.f - (tyy + tzz), txy - twz, txz + twy, data[4],
             txy + twz, 1.f - (txx + tzz), tyz - twx, data[5],
             txz - twy, tyz + twx, 1.f - (txx + tyy), data[6],
             0.f, 0.f, 0.f, 1.f}};
  }
};

// uw::CameraModel: reads the reference's calibration XML
// (calibration/*.xml: in/out size, "calibration_values" fx fy cx cy, "rectification").
class CameraModel {
 public:
  void GetCameraModel(const std::string& path) {  // src/CameraModel.cpp:30-99
    std::ifstream f(path);
    if (!f) throw Error(UWT_E_INVALID, "calibration file not found: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    in_width_ = (int)Number(s, "in_width");
    in_height_ = (int)Number(s, "in_height");
    out_width_ = (int)Number(s, "out_width");
    out_height_ = (int)Number(s, "out_height");
    Matrix4(s, "calibration_values", calib_);
    Matrix4(s, "rectification", dist_);
    if (calib_[2] < 1 && calib_[3] < 1) {  // normalised intrinsics, CameraModel.cpp:61-68
      calib_[0] *= in_width_;
      calib_[1] *= in_height_;
      calib_[2] *= in_width_;
      calib_[3] *= in_height_;
    }
    valid_ = dist_[0] != 0.f;  // CameraModel.cpp:78-83
    if (valid_) Rectify();
  }
  // The rectify branch without an XML file.
  void SetDistorted(int in_w, int in_h, int out_w, int out_h, float fx, float fy, float cx,
                    float cy, const float dist4[4]) {
    in_width_ = in_w; in_height_ = in_h; out_width_ = out_w; out_height_ = out_h;
    calib_[0] = fx; calib_[1] = fy; calib_[2] = cx; calib_[3] = cy;
    for (int i = 0; i < 4; ++i) dist_[i] = dist4[i];
    valid_ = dist_[0] != 0.f;
    if (valid_) Rectify();
  }
  void SetPinhole(int w, int h, float fx, float fy, float cx, float cy) {
    in_width_ = out_width_ = w;
    in_height_ = out_height_ = h;
    calib_[0] = fx; calib_[1] = fy; calib_[2] = cx; calib_[3] = cy;
    valid_ = false;
  }
  // 3x3 row-major K (CameraModel::GetK, src/CameraModel.cpp:113-115): output_intrinsic_camera_,
  // i.e. the new camera matrix when rectifying, else the original one (CameraModel.cpp:78-98)
  std::array<float, 9> GetK() const { return valid_ ? new_K_ : GetOriginalK(); }
  std::array<float, 9> GetOriginalK() const {
    return {{calib_[0], 0.f, calib_[2], 0.f, calib_[1], calib_[3], 0.f, 0.f, 1.f}};
  }
  // Fixed-point undistortion maps (CV_16SC2 / CV_16UC1 layout), empty when !IsValid()
  const std::vector<int16_t>& GetMap1() const { return map1_; }
  const std::vector<uint16_t>& GetMap2() const { return map2_; }
  // CameraModel::Undistort (src/CameraModel.cpp:101-103): in_w x in_h -> out_w x out_h
  std::vector<uint8_t> Undistort(const uint8_t* image, size_t row_stride = 0, int device = 0) const {
    std::vector<uint8_t> out((size_t)out_width_ * out_height_);
    const int rc = uwt_undistort_image(device, image, in_width_, in_height_,
                                       row_stride ? row_stride : (size_t)in_width_, map1_.data(),
                                       map2_.data(), out_width_, out_height_, out.data());
    if (rc != UWT_OK) throw Error(rc, "uwt_undistort_image failed");
    return out;
  }
  int GetOutputWidth() const { return out_width_; }
  int GetOutputHeight() const { return out_height_; }
  int GetInputWidth() const { return in_width_; }
  int GetInputHeight() const { return in_height_; }
  bool IsValid() const { return valid_; }

 private:
  static std::string Body(const std::string& s, const std::string& tag) {
    const size_t a = s.find("<" + tag);
    if (a == std::string::npos) throw Error(UWT_E_INVALID, "missing <" + tag + ">");
    const size_t b = s.find('>', a), c = s.find("</" + tag + ">", b);
    if (b == std::string::npos || c == std::string::npos)
      throw Error(UWT_E_INVALID, "malformed <" + tag + ">");
    return s.substr(b + 1, c - b - 1);
  }
  static double Number(const std::string& s, const std::string& tag) {
    return std::stod(Body(s, tag));
  }
  static void Matrix4(const std::string& s, const std::string& tag, float* out) {
    std::stringstream d(Body(Body(s, tag), "data"));
    for (int i = 0; i < 4; ++i)
      if (!(d >> out[i])) throw Error(UWT_E_INVALID, "<" + tag + "> needs 4 values");
  }
  void Rectify() {  // src/CameraModel.cpp:84-98
    const std::array<float, 9> K = GetOriginalK();
    int rc = uwt_camera_optimal_matrix(K.data(), dist_, in_width_, in_height_, 1.0, out_width_,
                                       out_height_, new_K_.data());
    if (rc != UWT_OK) throw Error(rc, "getOptimalNewCameraMatrix: bad calibration");
    map1_.resize((size_t)out_width_ * out_height_ * 2);
    map2_.resize((size_t)out_width_ * out_height_);
    rc = uwt_camera_undistort_maps(K.data(), dist_, new_K_.data(), out_width_, out_height_,
                                   map1_.data(), map2_.data());
    if (rc != UWT_OK) throw Error(rc, "initUndistortRectifyMap: singular camera matrix");
  }
  int in_width_ = 0, in_height_ = 0, out_width_ = 0, out_height_ = 0;
  std::array<float, 9> new_K_{{0, 0, 0, 0, 0, 0, 0, 0, 1}};
  std::vector<int16_t> map1_;
  std::vector<uint16_t> map2_;
  float calib_[4] = {0, 0, 0, 0};
  float dist_[4] = {0, 0, 0, 0};
  bool valid_ = false;
};

// System::CalculateROI (src/System.cpp:148-191) on the first undistorted image.
struct Rect {
  int x = 0, y = 0, width = 0, height = 0;
};
inline Rect CalculateROI(const std::vector<uint8_t>& undistorted, int w, int h) {
  int r[4];
  const int rc = uwt_calculate_roi(undistorted.data(), w, h, (size_t)w, r);
  if (rc != UWT_OK) throw Error(rc, "CalculateROI: the undistorted image is black on a mid line");
  Rect o;
  o.x = r[0]; o.y = r[1]; o.width = r[2]; o.height = r[3];
  return o;
}
// The reference takes w_, h_ from the ROI as found, which breaks its own pyramid sizing
// (w>>l vs cv::resize rounding); this library needs width % 16 == 0 and both sides divisible
// by 2^(levels-1): shrink the ROI to the largest compliant size, keeping the corner.
inline Rect AlignROI(Rect r, int levels = 5) {
  const int dv = 1 << (levels - 1), dw = dv > 16 ? dv : 16;
  r.width = r.width / dw * dw;
  r.height = r.height / dv * dv;
  return r;
}

class Tracker;

// uw::Frame: a handle on one device-side frame slot.  The per-level members of the reference
// (images_, gradientX_, gradientY_, gradient_, candidatePoints_) are read back on demand.
class Frame {
 public:
  int slot = -1;
  bool obtained_gradients_ = false;
  bool obtained_candidatePoints_ = false;
  bool depth_available_ = false;
  SE3f rigid_transformation_;

  std::vector<uint8_t> images(int lvl) const;
  std::vector<int16_t> gradientX(int lvl) const;
  std::vector<int16_t> gradientY(int lvl) const;
  std::vector<uint8_t> gradient(int lvl) const;
  std::vector<float> candidatePoints(int lvl) const;  // N x 4, rows [x y Z 1] (Z = 1 in mono)
  std::vector<uint16_t> depths(int lvl) const;

 private:
  friend class Tracker;
  Tracker* tracker_ = nullptr;
};

class Tracker {
 public:
  explicit Tracker(bool depth_available) {  // include/Tracker.h:97
    uwt_default_config(&cfg_);
    // depth_available_ (src/Tracker.cpp:275): candidates need depth != 0 and carry Z = d * 0.0002;
    // UWT_DEPTH_REFERENCE reads the depth like the reference does (at<uchar> on the 16-bit image),
    // set config().depth_mode = UWT_DEPTH_U16 for the at<ushort> reading
    if (depth_available) cfg_.depth_mode = UWT_DEPTH_REFERENCE;
  }
  ~Tracker() {
    if (h_) uwt_destroy(h_);
  }
  Tracker(const Tracker&) = delete;
  Tracker& operator=(const Tracker&) = delete;

  uwt_config& config() { return cfg_; }  // edit before InitializePyramid (levels, slots ...)

  // Tracker::InitializePyramid(int, int, Mat K), src/Tracker.cpp:297; K row-major 3x3.
  void InitializePyramid(int width, int height, const std::array<float, 9>& K) {
    cfg_.width = width;
    cfg_.height = height;
    cfg_.fx = K[0]; cfg_.fy = K[4]; cfg_.cx = K[2]; cfg_.cy = K[5];
    if (h_) uwt_destroy(h_);
    h_ = nullptr;
    src_w_ = src_h_ = 0;
    const int rc = uwt_create(&cfg_, &h_);
    if (rc != UWT_OK) throw Error(rc, uwt_last_error(nullptr));
    w_.clear(); h__.clear(); fx_.clear(); fy_.clear(); cx_.clear(); cy_.clear();
    for (int l = 0; l < cfg_.levels; ++l) {
      uwt_level_info li;
      Check(uwt_get_level_info(h_, l, &li));
      w_.push_back(li.width); h__.push_back(li.height);
      fx_.push_back(li.fx); fy_.push_back(li.fy); cx_.push_back(li.cx); cy_.push_back(li.cy);
    }
  }
  // System::AddFrame `remap(..); images_[0] = distortion(ROI)` (src/System.cpp:232-235): after
  // this, AddFrame takes DISTORTED in_w x in_h frames; remap + crop run inside the pyramid kernel.
  void SetUndistortion(const CameraModel& cam, int roi_x, int roi_y) {
    Check(uwt_set_undistortion(h_, cam.GetMap1().data(), cam.GetMap2().data(),
                               cam.GetOutputWidth(), cam.GetOutputHeight(), cam.GetInputWidth(),
                               cam.GetInputHeight(), roi_x, roi_y));
    src_w_ = cam.GetInputWidth();
    src_h_ = cam.GetInputHeight();
  }
  void InitializeMasks() {}  // dead weight in the reference (Tracker.cpp:342-359): no-op

  // System::AddFrame (src/System.cpp:225-262): gray 8-bit frame -> slot, builds the pyramid.
  Frame AddFrame(int slot, const uint8_t* gray, size_t row_stride = 0) {
    const size_t rs = row_stride ? row_stride : (size_t)(src_w_ ? src_w_ : cfg_.width);
    Check(uwt_upload_frames(h_, 1, &slot, gray, rs, rs * (src_h_ ? src_h_ : cfg_.height)));
    Frame f;
    f.slot = slot;
    f.tracker_ = this;
    return f;
  }
  // System::AddFrame depth part (src/System.cpp:241-250): 16-bit depth frame of the same slot
  void AddDepth(Frame* f, const uint16_t* depth, size_t row_stride_bytes = 0) {
    const size_t rs = row_stride_bytes ? row_stride_bytes : (size_t)cfg_.width * 2;
    Check(uwt_upload_depth_frames(h_, 1, &f->slot, depth, rs, rs * cfg_.height));
    f->depth_available_ = true;
  }
  void ApplyGradient(Frame* f) {  // include/Tracker.h:137
    Check(uwt_apply_gradient(h_, 1, &f->slot));
    f->obtained_gradients_ = true;
  }
  void ObtainCandidatePoints(Frame* f) {  // include/Tracker.h:145
    Check(uwt_select_candidates(h_, 1, &f->slot));
    f->obtained_candidatePoints_ = true;
  }
  // include/Tracker.h:153 (src/Tracker.cpp:1259): every pixel with depth > 0 is a point.  The
  // point rule is part of the tracker's configuration here: set
  // config().depth_mode = UWT_DEPTH_ALL_POINTS before InitializePyramid.
  void ObtainAllPoints(Frame* f) {
    if (cfg_.depth_mode != UWT_DEPTH_ALL_POINTS)
      throw std::runtime_error("ObtainAllPoints: config().depth_mode != UWT_DEPTH_ALL_POINTS");
    ObtainCandidatePoints(f);
  }
  // include/Tracker.h:122: writes prev->rigid_transformation_ (Tracker.cpp:595)
  void EstimatePose(Frame* prev, Frame* cur, uwt_track_stats* stats = nullptr) {
    Check(uwt_estimate_pose(h_, 1, &prev->slot, &cur->slot, nullptr,
                            prev->rigid_transformation_.data.data(), stats));
  }
  // include/Tracker.h:193: N x 4 points [x y Z W] -> N x 4 [x2 y2 Z' W']
  std::vector<float> WarpFunction(const std::vector<float>& pts4, const SE3f& T, int lvl) {
    std::vector<float> out(pts4.size());
    Check(uwt_warp_points(h_, pts4.data(), (int)(pts4.size() / 4), T.data.data(), lvl,
                          out.data()));
    return out;
  }

  uwt_tracker* handle() { return h_; }
  // per-level members of the reference (Tracker.h: w_, h_, fx_, fy_, cx_, cy_)
  std::vector<int> w_, h__;
  std::vector<float> fx_, fy_, cx_, cy_;

 private:
  friend class Frame;
  void Check(int rc) const {
    if (rc != UWT_OK) throw Error(rc, uwt_last_error(h_));
  }
  uwt_config cfg_;
  uwt_tracker* h_ = nullptr;
  int src_w_ = 0, src_h_ = 0;  // distorted input size when undistortion is on
};

inline std::vector<uint8_t> Frame::images(int lvl) const {
  std::vector<uint8_t> v((size_t)tracker_->w_[lvl] * tracker_->h__[lvl]);
  tracker_->Check(uwt_get_image(tracker_->h_, slot, lvl, v.data()));
  return v;
}
inline std::vector<int16_t> Frame::gradientX(int lvl) const {
  std::vector<int16_t> v((size_t)tracker_->w_[lvl] * tracker_->h__[lvl]);
  tracker_->Check(uwt_get_gradients(tracker_->h_, slot, lvl, v.data(), nullptr, nullptr));
  return v;
}
inline std::vector<int16_t> Frame::gradientY(int lvl) const {
  std::vector<int16_t> v((size_t)tracker_->w_[lvl] * tracker_->h__[lvl]);
  tracker_->Check(uwt_get_gradients(tracker_->h_, slot, lvl, nullptr, v.data(), nullptr));
  return v;
}
inline std::vector<uint8_t> Frame::gradient(int lvl) const {
  std::vector<uint8_t> v((size_t)tracker_->w_[lvl] * tracker_->h__[lvl]);
  tracker_->Check(uwt_get_gradients(tracker_->h_, slot, lvl, nullptr, nullptr, v.data()));
  return v;
}
inline std::vector<uint16_t> Frame::depths(int lvl) const {
  std::vector<uint16_t> v((size_t)tracker_->w_[lvl] * tracker_->h__[lvl]);
  tracker_->Check(uwt_get_depth(tracker_->h_, slot, lvl, v.data()));
  return v;
}
inline std::vector<float> Frame::candidatePoints(int lvl) const {
  int n = 0;
  tracker_->Check(uwt_get_candidate_count(tracker_->h_, slot, lvl, &n));
  std::vector<float> v((size_t)n * 4);
  if (n) tracker_->Check(uwt_get_candidates(tracker_->h_, slot, lvl, v.data(), n, &n));
  return v;
}

}  // namespace uw
