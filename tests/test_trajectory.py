"""Caller-side pose chaining and trajectory I/O (Visualizer.cpp:303-325,
449-505): Python mirror and C++ header against the oracle, bit for bit.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_poses(oracle, n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        a = np.empty(6, np.float32)
        a[:3] = rng.normal(size=3) * 4e-3
        a[3:] = rng.normal(size=3) * 5e-3
        out.append(oracle.se3_exp(a))
    return np.array(out, np.float32)


def test_python_chain_matches_oracle(oracle):
    from uw_slam_b200 import trajectory as T
    poses = random_poses(oracle, 300, 1)
    tr = T.Trajectory()
    ref = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    for p in poses:
        ref = oracle.chain_pose(ref, p, 40.0)
        assert np.array_equal(tr.Update(p), ref)
    assert len(tr.poses) == 300
    # the product itself against the oracle's Sophus restatement, incl. non-unit inputs
    rng = np.random.default_rng(2)
    for _ in range(200):
        a, b = poses[rng.integers(300)].copy(), poses[rng.integers(300)].copy()
        a[:4] *= np.float32(1 + rng.normal() * 1e-3)
        assert np.array_equal(T.se3_mul(a, b), oracle.se3_mul(a, b))
    assert np.array_equal(T.Trajectory.CameraPosition(ref), [-ref[6], -ref[4], -ref[5]])


def test_cpp_header_chain_and_ground_truth(tmp_path, oracle):
    from uw_slam_b200 import trajectory as T
    exe = tmp_path / "chain"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off",
                           "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "chain_trajectory.cpp"), "-o", str(exe)])
    poses = random_poses(oracle, 64, 3)
    with open(tmp_path / "p.txt", "w") as f:
        for p in poses:
            f.write(" ".join("%.9g" % v for v in p) + "\n")
    # EuRoC-style ground truth: header + 1000 rows; 600-sample offset, step = rows // images
    rng = np.random.default_rng(4)
    gt = np.cumsum(rng.normal(size=(1000, 3)) * 1e-2, axis=0)
    with open(tmp_path / "gt.csv", "w") as f:
        f.write("#timestamp,p_x,p_y,p_z,q_w,q_x,q_y,q_z\n")
        for i, g in enumerate(gt):
            f.write("%d,%.17g,%.17g,%.17g,1,0,0,0\n" % (i, g[0], g[1], g[2]))
    r = subprocess.run([str(exe), str(tmp_path / "p.txt"), "40", str(tmp_path / "gt.csv"),
                        "EUROC", "200", "0"], capture_output=True, text=True, check=True)
    got = np.array([[np.float32(v) for v in ln.split()] for ln in r.stdout.strip().splitlines()],
                   np.float32)
    ref, chain = np.array([0, 0, 0, 1, 0, 0, 0], np.float32), []
    for p in poses:
        ref = oracle.chain_pose(ref, p, 40.0)
        chain.append(ref)
    assert np.array_equal(got, np.array(chain))
    # same evaluation with the Python mirror
    rows = T.ReadGroundTruthEUROC(str(tmp_path / "gt.csv"))
    assert rows.shape == (1000, 7) and np.array_equal(rows[:, :3], gt)
    cur = T.GroundTruthCursor(len(rows), 200, 0, True)
    assert (cur.step, cur.index) == (5, 600)
    est, refp = [], []
    for p in chain:
        if cur.index >= len(rows):
            break
        est.append(T.Trajectory.CameraPosition(p))
        refp.append(rows[cur.index][:3])
        cur.Advance()
    ate, rpe = T.Evaluate(est, refp)
    fields = r.stderr.split()
    assert int(fields[1]) == len(est)
    assert abs(float(fields[3]) - ate) <= 1e-7 * max(1, ate)
    assert abs(float(fields[5]) - rpe) <= 1e-7 * max(1, rpe)


def test_tum_ground_truth_reader(tmp_path):
    from uw_slam_b200 import trajectory as T
    p = tmp_path / "gt.txt"
    p.write_text("# ground truth trajectory\n# file: x\n# timestamp tx ty tz qx qy qz qw\n"
                 "1.5 1 2 3 0 0 0 1\n2.5 4 5 6 0.5 0.5 0.5 0.5\n")
    rows = T.ReadGroundTruthTUM(str(p))
    assert rows.tolist() == [[1, 2, 3, 0, 0, 0, 1], [4, 5, 6, .5, .5, .5, .5]]
    assert T.GroundTruthPose(rows[1], False).tolist() == [4, 5, 6, .5, .5, .5, .5]
    assert T.GroundTruthPose([1, 2, 3, 9, 5, 6, 7], True).tolist() == [1, 2, 3, 5, 6, 7, 9]
    c = T.GroundTruthCursor(1000, 300, 2, False)
    assert (c.step, c.index) == (3, 6) and c.Advance() == 9
    tr = T.Trajectory()
    tr.Update(np.array([0, 0, 0, 1, 0.01, 0, 0], np.float32))
    tr.WriteTUM(str(tmp_path / "o.txt"))
    assert (tmp_path / "o.txt").read_text().split()[1] == \
        "%.9g" % (np.float32(40) * np.float32(0.01))   # the reference's x40 translation scale
