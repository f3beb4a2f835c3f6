// chain_trajectory.cpp -- Visualizer::UpdateMessages pose chaining (src/Visualizer.cpp:303-325)
// on a file of per-frame poses, without ROS.
//   chain_trajectory <poses.txt> [scale] [gt.csv TUM|EUROC num_images start_index]
// poses.txt: one "qx qy qz qw tx ty tz" line per tracked frame (what track_sequence prints).
// Prints the chained pose per frame (%.9g, exact floats); with ground truth, also ATE / RPE of
// the Rviz-frame camera positions against the sub-sampled ground-truth positions on stderr.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "uw/uw_trajectory.hpp"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s poses.txt [scale] [gt TUM|EUROC num_images start]\n", argv[0]);
    return 2;
  }
  try {
    std::ifstream in(argv[1]);
    if (!in) throw std::runtime_error(std::string("cannot open ") + argv[1]);
    uw::Trajectory traj(argc > 2 ? (float)std::atof(argv[2]) : 40.0f);
    uw::Pose7 p;
    while (in >> p[0] >> p[1] >> p[2] >> p[3] >> p[4] >> p[5] >> p[6]) {
      const uw::Pose7& f = traj.Update(p);
      std::printf("%.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", f[0], f[1], f[2], f[3], f[4], f[5], f[6]);
    }
    if (argc >= 7) {
      const bool euroc = std::string(argv[4]) == "EUROC";
      const auto gt = euroc ? uw::ReadGroundTruthEUROC(argv[3]) : uw::ReadGroundTruthTUM(argv[3]);
      uw::GroundTruthCursor cur((int)gt.size(), std::atoi(argv[5]), std::atoi(argv[6]), euroc);
      std::vector<std::array<double, 3>> est, ref;
      for (const uw::Pose7& q : traj.poses()) {
        if (cur.index >= (int)gt.size()) break;
        const auto c = uw::Trajectory::CameraPosition(q);
        est.push_back({{c[0], c[1], c[2]}});
        ref.push_back({{gt[cur.index][0], gt[cur.index][1], gt[cur.index][2]}});
        cur.Advance();
      }
      const uw::TrajectoryError e = uw::Evaluate(est, ref);
      std::fprintf(stderr, "samples %zu ATE %.9g RPE %.9g\n", est.size(), e.ate_rmse, e.rpe_rmse);
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
