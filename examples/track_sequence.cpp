// track_sequence.cpp -- the reference's main loop (src/main_uw_slam.cpp:139-151 with
// System::Tracking in the direct order, src/System.cpp:201-220) on top of the C++ facade.
//
//   track_sequence <calibration.xml> <frames.raw> <n_frames>
// frames.raw: n_frames gray 8-bit frames of the calibration's INPUT size, back to back.  With
// non-zero distortion coefficients the frames are rectified and cropped as in
// System::InitializeSystem / CalculateROI / AddFrame (src/System.cpp:105-119,148-191,232-235).
// Prints one line per tracked frame: the pose "qx qy qz qw tx ty tz" (%.9g, exact floats).
//
// Build: g++ -std=c++17 -Iinclude examples/track_sequence.cpp -Luw_slam_b200 -luwtrack
//        -Wl,-rpath,$ORIGIN/../uw_slam_b200 -o examples/track_sequence
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "uw/uw_tracker.hpp"

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s calibration.xml frames.raw n_frames\n", argv[0]);
    return 2;
  }
  try {
    uw::CameraModel camera;
    camera.GetCameraModel(argv[1]);  // System::Calibration, src/System.cpp:77-89
    int w = camera.GetOutputWidth(), h = camera.GetOutputHeight();
    const int iw = camera.GetInputWidth(), ih = camera.GetInputHeight();
    const int n = std::atoi(argv[3]);
    std::vector<uint8_t> frames((size_t)iw * ih * n);
    std::FILE* f = std::fopen(argv[2], "rb");
    if (!f || std::fread(frames.data(), 1, frames.size(), f) != frames.size()) {
      std::fprintf(stderr, "cannot read %d frames of %dx%d from %s\n", n, iw, ih, argv[2]);
      return 2;
    }
    std::fclose(f);

    uw::Rect roi;
    if (camera.IsValid()) {  // distortion_valid_: CalculateROI on the first image, System.cpp:117-119
      roi = uw::AlignROI(uw::CalculateROI(camera.Undistort(frames.data()), w, h));
      w = roi.width;
      h = roi.height;
      std::fprintf(stderr, "rectifying: ROI %dx%d at (%d,%d)\n", w, h, roi.x, roi.y);
    }
    uw::Tracker tracker(false);                      // src/System.cpp:121
    tracker.config().max_frames = 2;
    tracker.InitializePyramid(w, h, camera.GetK());  // src/System.cpp:122
    if (camera.IsValid()) tracker.SetUndistortion(camera, roi.x, roi.y);

    uw::Frame previous = tracker.AddFrame(0, frames.data());  // System::AddFrame
    for (int i = 1; i < n; ++i) {
      uw::Frame current = tracker.AddFrame(i % 2, frames.data() + (size_t)i * iw * ih);
      if (!previous.obtained_gradients_) tracker.ApplyGradient(&previous);
      if (!previous.obtained_candidatePoints_) tracker.ObtainCandidatePoints(&previous);
      tracker.ApplyGradient(&current);
      tracker.EstimatePose(&previous, &current);
      const auto& p = previous.rigid_transformation_.data;
      std::printf("%.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", p[0], p[1], p[2], p[3], p[4], p[5], p[6]);
      previous = current;
    }
  } catch (const uw::Error& e) {
    std::fprintf(stderr, "uwtrack error %d: %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}
