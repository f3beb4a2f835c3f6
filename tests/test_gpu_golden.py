"""GPU: the CUDA path against the committed golden vectors (no oracle involved), and the C++
facade example run end to end."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from uw_slam_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tracker_for(calib, **cfg):
    import uw_slam_b200 as U
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2, **cfg)
    return t


def check_trace(gold, key, trace):
    t = np.array([[x.level, x.k, x.n_valid, x.broke] for x in trace], np.int32)
    assert np.array_equal(t, gold[key + "_lvl_k_nvalid_broke"])
    assert np.array_equal(np.array([x.sum_r2 for x in trace], np.int64), gold[key + "_sum_r2"])
    for name in ("A", "b", "delta", "pose"):
        got = np.array([getattr(x, name)[:] for x in trace], np.float32)
        assert np.array_equal(got, gold[key + "_" + name]), name


@pytest.mark.parametrize("calib,seed,mode", [("tiny", 0, 0), ("tiny", 1, 1), ("small", 0, 0),
                                             ("small", 1, 1), ("small", 2, 0)])
def test_small_golden_on_gpu(calib, seed, mode):
    from uw_slam_b200 import _lib as L
    gold = np.load(os.path.join(GOLD, "golden_small.npz"))
    key = "%s_%d" % (calib, seed)
    t = tracker_for(calib, flags=L.FLAG_TRACE, solve_mode=mode)
    fp, fc = t.AddFrames([0, 1], np.stack([gold[key + "_prev"], gold[key + "_cur"]]))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    for l in range(5):
        assert np.array_equal(fp.image(l), gold["%s_img%d" % (key, l)])
        gx, gy, g = fp.gradients(l)
        assert np.array_equal(gx, gold["%s_gx%d" % (key, l)])
        assert np.array_equal(gy, gold["%s_gy%d" % (key, l)])
        assert np.array_equal(g, gold["%s_g%d" % (key, l)])
        c = fp.candidatePoints(l)
        assert np.array_equal(c[:, :2].astype(np.uint16), gold["%s_cand%d" % (key, l)])
    pose = t.EstimatePose(fp, fc)[0]
    check_trace(gold, "%s_m%d" % (key, mode), t.get_trace(0))
    assert np.array_equal(pose, gold["%s_m%d_final" % (key, mode)])
    t.close()


@pytest.mark.parametrize("calib,seed", [("tum", 0), ("tum", 1), ("tum", 2), ("tum", 3),
                                        ("euroc", 0), ("euroc", 1), ("tum_mono", 0)])
def test_big_golden_on_gpu(calib, seed):
    from uw_slam_b200 import _lib as L
    gold = np.load(os.path.join(GOLD, "golden_big.npz"))
    key = "%s_%d" % (calib, seed)
    prev, cur, _, _ = synth.render_pair(calib, seed)
    if [sha(prev), sha(cur)] != list(gold[key + "_input_sha"]):
        pytest.skip("numpy on this box renders different input bytes than the fixture")
    t = tracker_for(calib, flags=L.FLAG_TRACE)
    fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    for l in range(5):
        gx, gy, g = fp.gradients(l)
        got = [sha(fp.image(l)), sha(gx), sha(gy), sha(g), sha(fp.candidatePoints(l))]
        assert got == list(gold[key + "_level_sha"][l]), l
    pose, stats = t.EstimatePose(fp, fc, return_stats=True)
    check_trace(gold, key, t.get_trace(0))
    assert np.array_equal(pose[0], gold[key + "_final"])
    assert list(stats[0].iterations)[:5] == list(gold[key + "_iterations"])
    t.close()


@pytest.mark.parametrize("calib,seed", [("tiny", 0), ("small", 1), ("small", 2), ("tum", 0),
                                        ("tum", 4)])
@pytest.mark.parametrize("name,mode", [("tukey", 1), ("huber", 2)])
def test_robust_golden_on_gpu(calib, seed, name, mode):
    # Tukey/MAD (Tracker.cpp:496,1571-1654) and Huber weights
    from uw_slam_b200 import _lib as L
    gold = np.load(os.path.join(GOLD, "golden_robust.npz"))
    key = "%s_%d_%s" % (calib, seed, name)
    prev, cur, _, _ = synth.render_pair(calib, seed)
    if [sha(prev), sha(cur)] != list(gold[key + "_input_sha"]):
        pytest.skip("numpy on this box renders different input bytes than the fixture")
    t = tracker_for(calib, flags=L.FLAG_TRACE, weight_mode=mode, huber_delta=7.5)
    fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    pose, stats = t.EstimatePose(fp, fc, return_stats=True)
    tr = t.get_trace(0)
    check_trace(gold, key, tr)
    assert np.array_equal(np.array([x.error for x in tr], np.float32), gold[key + "_error"])
    assert np.array_equal(pose[0], gold[key + "_final"])
    assert list(stats[0].iterations)[:5] == list(gold[key + "_iterations"])
    t.close()


CALIB_XML = """<?xml version="1.0"?>
<opencv_storage>
<in_width type_id="integer"> {w} </in_width>
<in_height type_id="integer"> {h} </in_height>
<out_width type_id="integer"> {w} </out_width>
<out_height type_id="integer"> {h} </out_height>
<calibration_values type_id="opencv-matrix"><rows>1</rows><cols>4</cols><dt>f</dt>
  <data> {fx} {fy} {cx} {cy} </data></calibration_values>
<rectification type_id="opencv-matrix"><rows>1</rows><cols>4</cols><dt>f</dt>
  <data> {dist} </data></rectification>
</opencv_storage>
"""


def test_cpp_facade_example_tracks_a_sequence(tmp_path, oracle):
    from uw_slam_b200 import build
    so = build.build()
    exe = tmp_path / "track_sequence"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "track_sequence.cpp"),
                           "-L", os.path.dirname(so), "-luwtrack",
                           "-Wl,-rpath," + os.path.dirname(so), "-o", str(exe)])
    calib = "small"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    (tmp_path / "c.xml").write_text(CALIB_XML.format(w=w, h=h, fx=fx, fy=fy, cx=cx, cy=cy,
                                                     dist="0 0 0 1"))
    frames, _, _ = synth.render_sequence(calib, 4, 5, rot=3e-3, trans=3e-3)
    np.stack(frames).tofile(str(tmp_path / "f.raw"))
    out = subprocess.check_output([str(exe), str(tmp_path / "c.xml"), str(tmp_path / "f.raw"),
                                   "5"], text=True)
    poses = np.array([[np.float32(v) for v in ln.split()] for ln in out.strip().splitlines()],
                     np.float32)
    assert poses.shape == (4, 7)
    ref, _, _ = oracle.track_sequence(oracle.default_params(w, h, fx, fy, cx, cy), np.stack(frames))
    assert np.array_equal(poses, ref)


def test_cpp_facade_example_rectifies_and_tracks(tmp_path, oracle):
    # the rectify branch end to end in C++: CameraModel (maps), Undistort, CalculateROI, the
    # fused remap + crop + pyramid, tracking -- against the oracle pipeline on the same frames
    from uw_slam_b200 import build
    so = build.build()
    exe = tmp_path / "track_sequence"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "track_sequence.cpp"),
                           "-L", os.path.dirname(so), "-luwtrack",
                           "-Wl,-rpath," + os.path.dirname(so), "-o", str(exe)])
    w, h, fx, fy, cx, cy = 320, 240, 230.0, 228.0, 158.0, 121.5
    dist = [-0.25, 0.06, 0.0005, -0.0003]
    (tmp_path / "c.xml").write_text(CALIB_XML.format(
        w=w, h=h, fx=fx, fy=fy, cx=cx, cy=cy, dist=" ".join("%.9g" % v for v in dist)))
    synth.CALIB["_facade_und"] = (w, h, fx, fy, cx, cy)
    frames, _, _ = synth.render_sequence("_facade_und", 3, 4, rot=3e-3, trans=3e-3)
    frames = np.maximum(np.stack(frames), 1)
    frames.tofile(str(tmp_path / "f.raw"))
    out = subprocess.check_output([str(exe), str(tmp_path / "c.xml"), str(tmp_path / "f.raw"),
                                   "4"], text=True)
    poses = np.array([[np.float32(v) for v in ln.split()] for ln in out.strip().splitlines()],
                     np.float32)
    assert poses.shape == (3, 7)
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    d = np.array(dist, np.float32)
    nK = oracle.optimal_new_camera_matrix(K, d, (w, h), 1.0, (w, h))
    m1, m2 = oracle.init_undistort_rectify_map(K, d, nK, (w, h))
    und = [oracle.remap_bilinear(f, m1, m2) for f in frames]
    x, y, rw, rh = oracle.calculate_roi(und[0])
    rw, rh = rw // 16 * 16, rh // 16 * 16
    crop = np.stack([u[y:y + rh, x:x + rw] for u in und])
    p = oracle.default_params(rw, rh, float(nK[0, 0]), float(nK[1, 1]), float(nK[0, 2]),
                              float(nK[1, 2]))
    ref, _, _ = oracle.track_sequence(p, crop)
    assert np.array_equal(poses, ref)
