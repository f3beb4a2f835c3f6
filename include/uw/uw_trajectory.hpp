// uw_trajectory.hpp -- caller-side pose chaining and trajectory I/O without ROS
// (SURVEY.md 8-f row 4).  Header-only, host-only, no dependency on libuwtrack.
//
//   reference                                                   here
//   Visualizer::UpdateMessages pose composition                 uw::Trajectory::Update
//       src/Visualizer.cpp:303-325
//   Visualizer::ReadGroundTruthTUM / ReadGroundTruthEUROC       uw::ReadGroundTruthTUM / EUROC
//       src/Visualizer.cpp:449-505
//   ground-truth stepping (ground_truth_step_/index_)           uw::GroundTruthCursor
//       src/Visualizer.cpp:475-477, 502-504, 367
//   outputFile CSV row (camera pose, ground-truth pose)         uw::Trajectory::WriteCsvRow
//       src/Visualizer.cpp:383-397
// Extensions (not in the reference): TUM-format trajectory file, ATE / RPE against ground truth.
//
// Float arithmetic follows docs/ARITHMETIC.md U7 (Sophus SE3f product as scalar Hamilton product
// + Eigen _transformVector, no FMA) so that a chained trajectory is reproducible bit for bit.
#pragma once
#include <array>
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace uw {

// 7 floats in Sophus storage order: qx qy qz qw tx ty tz (thirdparty/sophus/se3.hpp:469-472)
using Pose7 = std::array<float, 7>;

namespace detail {
inline void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
// SE3Base::operator* (se3.hpp:285-321): t = a.t + a.q * b.t; q = a.q * b.q, renormalised by
// 2 / (1 + |q|^2) when |q|^2 != 1 (so3.hpp:338-355)
inline Pose7 se3_mul(const Pose7& a, const Pose7& b) {
  Pose7 r;
  float uv[3], c[3];
  cross3(a.data(), b.data() + 4, uv);
  for (int i = 0; i < 3; ++i) uv[i] = uv[i] + uv[i];
  cross3(a.data(), uv, c);
  for (int i = 0; i < 3; ++i) r[4 + i] = a[4 + i] + ((b[4 + i] + a[3] * uv[i]) + c[i]);
  const float ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const float bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[3] = aw * bw - ax * bx - ay * by - az * bz;
  r[0] = aw * bx + ax * bw + ay * bz - az * by;
  r[1] = aw * by + ay * bw + az * bx - ax * bz;
  r[2] = aw * bz + az * bw + ax * by - ay * bx;
  const float sn = (r[0] * r[0] + r[1] * r[1]) + (r[2] * r[2] + r[3] * r[3]);
  if (sn != 1.0f) {
    const float s = 2.0f / (1.0f + sn);
    for (int i = 0; i < 4; ++i) r[i] = r[i] * s;
  }
  return r;
}
}  // namespace detail

// SE3(Quaternion, Point) (se3.hpp:446-448): the quaternion is normalised (so3.hpp:270-276,434-440)
inline Pose7 MakeSE3(const float q[4], const float t[3]) {
  const float len = std::sqrt((q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]));
  return {{q[0] / len, q[1] / len, q[2] / len, q[3] / len, t[0], t[1], t[2]}};
}

// One ground-truth sample: 7 doubles after the timestamp, in file order
// (TUM: tx ty tz qx qy qz qw; EuRoC: px py pz qw qx qy qz).
using GroundTruthRow = std::array<double, 7>;

namespace detail {
inline std::vector<GroundTruthRow> read_rows(const std::string& path, int skip_lines, char sep) {
  std::ifstream file(path);
  if (!file.is_open()) throw std::runtime_error("Could not read file " + path);
  std::string line;
  for (int i = 0; i < skip_lines; ++i) std::getline(file, line);
  std::vector<GroundTruthRow> rows;
  while (std::getline(file, line)) {
    if (line.empty()) continue;
    std::stringstream iss(line);
    std::string val;
    std::getline(iss, val, sep);  // timestamp
    GroundTruthRow r;
    for (int i = 0; i < 7; ++i) {
      if (!std::getline(iss, val, sep)) throw std::runtime_error("short ground-truth row: " + line);
      r[i] = std::stod(val);
    }
    rows.push_back(r);
  }
  return rows;
}
}  // namespace detail

// Visualizer::ReadGroundTruthTUM (src/Visualizer.cpp:449-478): three header lines, ' ' separated
inline std::vector<GroundTruthRow> ReadGroundTruthTUM(const std::string& path) {
  return detail::read_rows(path, 3, ' ');
}
// Visualizer::ReadGroundTruthEUROC (src/Visualizer.cpp:480-505): one header line, ',' separated
inline std::vector<GroundTruthRow> ReadGroundTruthEUROC(const std::string& path) {
  return detail::read_rows(path, 1, ',');
}

// ground_truth_step_ / ground_truth_index_ (src/Visualizer.cpp:475-477, 502-504): the ground
// truth is sub-sampled with the integer step num_poses / num_images; EuRoC starts 600 samples in.
struct GroundTruthCursor {
  int step = 1, index = 0;
  GroundTruthCursor(int num_poses, int num_images, int start_index, bool euroc) {
    step = num_poses / num_images;
    index = start_index * step + (euroc ? 600 : 0);
  }
  int Advance() { return index += step; }  // src/Visualizer.cpp:367
};

// Position + orientation (x y z w) of one ground-truth row in the marker convention of
// src/Visualizer.cpp:340-357 (EuRoC rows store qw first).
inline std::array<double, 7> GroundTruthPose(const GroundTruthRow& r, bool euroc) {
  if (euroc) return {{r[0], r[1], r[2], r[4], r[5], r[6], r[3]}};
  return {{r[0], r[1], r[2], r[3], r[4], r[5], r[6]}};
}

class Trajectory {
 public:
  // translation_scale: the reference multiplies every per-frame translation by 40 before
  // chaining ("corrected movement of camera", src/Visualizer.cpp:303-307)
  explicit Trajectory(float translation_scale = 40.0f) : scale_(translation_scale) {}

  void SetInitialPose(const Pose7& p) { previous_pose_ = p; }

  // Visualizer::UpdateMessages (src/Visualizer.cpp:303-325) for frame->rigid_transformation_.
  // Returns final_pose and appends it to the trajectory.
  const Pose7& Update(const Pose7& rigid_transformation) {
    const float t[3] = {scale_ * rigid_transformation[4], scale_ * rigid_transformation[5],
                        scale_ * rigid_transformation[6]};
    const Pose7 current = MakeSE3(rigid_transformation.data(), t);
    previous_pose_ = detail::se3_mul(previous_pose_, current);
    poses_.push_back(previous_pose_);
    return previous_pose_;
  }

  // camera_pose_.pose.position in the Rviz frame (src/Visualizer.cpp:316-318)
  static std::array<float, 3> CameraPosition(const Pose7& p) { return {{-p[6], -p[4], -p[5]}}; }

  const std::vector<Pose7>& poses() const { return poses_; }
  const Pose7& current() const { return previous_pose_; }

  // One row of the reference's output CSV (src/Visualizer.cpp:383-397): camera position (Rviz
  // frame) and orientation, then the ground-truth position and orientation.
  static void WriteCsvRow(std::ostream& os, const Pose7& p, const std::array<double, 7>& gt) {
    const auto c = CameraPosition(p);
    os << c[0] << "," << c[1] << "," << c[2] << "," << p[0] << "," << p[1] << "," << p[2] << ","
       << p[3] << "," << gt[0] << "," << gt[1] << "," << gt[2] << "," << gt[3] << "," << gt[4]
       << "," << gt[5] << "," << gt[6] << "\n";
  }

  // Extension: the chained poses in the TUM trajectory format "stamp tx ty tz qx qy qz qw".
  void WriteTUM(std::ostream& os, const std::vector<double>& stamps) const {
    os.precision(9);
    for (size_t i = 0; i < poses_.size(); ++i) {
      const Pose7& p = poses_[i];
      os << (i < stamps.size() ? stamps[i] : (double)i) << " " << p[4] << " " << p[5] << " "
         << p[6] << " " << p[0] << " " << p[1] << " " << p[2] << " " << p[3] << "\n";
    }
  }

 private:
  float scale_;
  Pose7 previous_pose_{{0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f}};
  std::vector<Pose7> poses_;
};

// Extension: translational errors of an estimated trajectory against ground-truth positions
// (both N x 3, same frame and scale).  ATE: RMSE of position differences after subtracting the
// first-sample offset; RPE: RMSE of the differences of consecutive displacement vectors.
struct TrajectoryError {
  double ate_rmse = 0, rpe_rmse = 0;
};
inline TrajectoryError Evaluate(const std::vector<std::array<double, 3>>& est,
                                const std::vector<std::array<double, 3>>& gt) {
  TrajectoryError e;
  const size_t n = est.size() < gt.size() ? est.size() : gt.size();
  if (n == 0) return e;
  double sa = 0, sr = 0;
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      const double d = (est[i][k] - est[0][k]) - (gt[i][k] - gt[0][k]);
      sa += d * d;
      if (i > 0) {
        const double r = (est[i][k] - est[i - 1][k]) - (gt[i][k] - gt[i - 1][k]);
        sr += r * r;
      }
    }
  e.ate_rmse = std::sqrt(sa / n);
  e.rpe_rmse = n > 1 ? std::sqrt(sr / (n - 1)) : 0.0;
  return e;
}

}  // namespace uw
