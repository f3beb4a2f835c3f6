"""Caller-side pose chaining and trajectory I/O without ROS: the Python
mirror of include/uw/uw_trajectory.hpp.

  Visualizer::UpdateMessages pose composition   src/Visualizer.cpp:303-325  -> Trajectory.Update
  Visualizer::ReadGroundTruthTUM / EUROC        src/Visualizer.cpp:449-505  -> ReadGroundTruth*
  ground_truth_step_ / ground_truth_index_      src/Visualizer.cpp:475-477  -> GroundTruthCursor

All pose arithmetic is float32 in the operation order of docs/ARITHMETIC.md U7.
"""
import numpy as np

f32 = np.float32


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def se3_mul(a, b):
    """Sophus SE3f product on 7-float poses (qx qy qz qw tx ty tz), se3.hpp:285-321."""
    a = [f32(v) for v in a]
    b = [f32(v) for v in b]
    uv = _cross(a[:3], b[4:])
    uv = [u + u for u in uv]
    c = _cross(a[:3], uv)
    t = [a[4 + i] + ((b[4 + i] + a[3] * uv[i]) + c[i]) for i in range(3)]
    ax, ay, az, aw = a[:4]
    bx, by, bz, bw = b[:4]
    q = [aw * bx + ax * bw + ay * bz - az * by,
         aw * by + ay * bw + az * bx - ax * bz,
         aw * bz + az * bw + ax * by - ay * bx,
         aw * bw - ax * bx - ay * by - az * bz]
    sn = (q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3])
    if sn != f32(1):
        s = f32(2) / (f32(1) + sn)
        q = [v * s for v in q]
    return np.array(q + t, f32)


def make_se3(q, t):
    """SE3(Quaternion, Point): the quaternion is normalised (so3.hpp:270-276,434-440)."""
    q = [f32(v) for v in q]
    ln = np.sqrt((q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]))
    return np.array([v / ln for v in q] + [f32(v) for v in t], f32)


def _read_rows(path, skip, sep):
    rows = []
    with open(path) as f:
        for i, line in enumerate(f):
            line = line.rstrip("\n")
            if i < skip or not line:
                continue
            parts = line.split(sep)
            rows.append([float(v) for v in parts[1:8]])
    return np.array(rows, np.float64).reshape(-1, 7)


def ReadGroundTruthTUM(path):
    """src/Visualizer.cpp:449-478: three header lines, space separated; tx ty tz qx qy qz qw."""
    return _read_rows(path, 3, " ")


def ReadGroundTruthEUROC(path):
    """src/Visualizer.cpp:480-505: one header line, comma separated; px py pz qw qx qy qz."""
    return _read_rows(path, 1, ",")


class GroundTruthCursor:
    """ground_truth_step_ / ground_truth_index_ (src/Visualizer.cpp:475-477, 502-504, 367)."""

    def __init__(self, num_poses, num_images, start_index, euroc):
        self.step = num_poses // num_images
        self.index = start_index * self.step + (600 if euroc else 0)

    def Advance(self):
        self.index += self.step
        return self.index


def GroundTruthPose(row, euroc):
    r = list(row)
    return np.array([r[0], r[1], r[2], r[4], r[5], r[6], r[3]] if euroc else r, np.float64)


class Trajectory:
    def __init__(self, translation_scale=40.0):
        self.scale = f32(translation_scale)   # src/Visualizer.cpp:303-307
        self.previous_pose_ = np.array([0, 0, 0, 1, 0, 0, 0], f32)
        self.poses = []

    def SetInitialPose(self, p):
        self.previous_pose_ = np.asarray(p, f32).copy()

    def Update(self, rigid_transformation):
        """Visualizer::UpdateMessages (src/Visualizer.cpp:303-325)."""
        r = np.asarray(rigid_transformation, f32)
        cur = make_se3(r[:4], [self.scale * r[4], self.scale * r[5], self.scale * r[6]])
        self.previous_pose_ = se3_mul(self.previous_pose_, cur)
        self.poses.append(self.previous_pose_.copy())
        return self.previous_pose_

    @staticmethod
    def CameraPosition(p):
        """camera_pose_.pose.position in the Rviz frame (src/Visualizer.cpp:316-318)."""
        return np.array([-p[6], -p[4], -p[5]], f32)

    def WriteTUM(self, path, stamps=None):
        with open(path, "w") as f:
            for i, p in enumerate(self.poses):
                s = stamps[i] if stamps is not None else float(i)
                f.write("%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n" %
                        (s, p[4], p[5], p[6], p[0], p[1], p[2], p[3]))


def Evaluate(est_xyz, gt_xyz):
    """ATE (first-sample aligned) and RPE (consecutive displacements), RMSE; N x 3 arrays."""
    e = np.asarray(est_xyz, np.float64)
    g = np.asarray(gt_xyz, np.float64)
    n = min(len(e), len(g))
    if n == 0:
        return 0.0, 0.0
    e, g = e[:n], g[:n]
    d = (e - e[0]) - (g - g[0])
    ate = float(np.sqrt((d * d).sum() / n))
    r = np.diff(e, axis=0) - np.diff(g, axis=0)
    rpe = float(np.sqrt((r * r).sum() / (n - 1))) if n > 1 else 0.0
    return ate, rpe
