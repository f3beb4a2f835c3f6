"""Calibration / undistortion front-end: CameraModel rectify branch
(CameraModel.cpp:84-103), System::CalculateROI (System.cpp:148-191), remap + ROI crop in
System::AddFrame (System.cpp:232-235).

CPU part: the oracle AND the product's host functions against the real OpenCV (cv2).
GPU part: the stand-alone remap and the remap fused into the pyramid kernel against the oracle
and the committed fixture (tests/golden/golden_undistort.npz, made by make_golden_undistort.py).
"""
import hashlib
import os

import numpy as np
import pytest

from uw_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "golden_undistort.npz")

EUROC = dict(in_size=(752, 480), out_size=(736, 480), fx=458.654, fy=457.296, cx=367.215,
             cy=248.375, dist=[-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
# calibration/calibrationTUM.xml has zero distortion; a TUM-mono-like wide-angle case instead
WIDE = dict(in_size=(1280, 1024), out_size=(1280, 1024), fx=685.72, fy=685.64, cx=630.86,
            cy=511.92, dist=[-0.21, 0.045, -0.0004, 0.0007])
SMALL = dict(in_size=(320, 240), out_size=(304, 240), fx=230.0, fy=228.0, cx=158.0, cy=121.5,
             dist=[-0.25, 0.06, 0.0005, -0.0003])
CASES = {"euroc": EUROC, "wide": WIDE, "small": SMALL}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def K_of(c):
    return np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1]], np.float32)


def make_image(c, seed=0):
    w, h = c["in_size"]
    rng = np.random.default_rng(seed)
    # smooth + noise, never 0 (CalculateROI looks for black = outside the remapped frame)
    ys, xs = np.mgrid[0:h, 0:w]
    img = 120 + 60 * np.sin(xs * 0.05) * np.cos(ys * 0.04) + rng.integers(-40, 40, (h, w))
    return np.clip(img, 1, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# CPU: oracle and product host functions vs cv2
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_front_end_matches_cv2(oracle, name):
    cv2 = pytest.importorskip("cv2")
    c = CASES[name]
    K, d = K_of(c), np.array(c["dist"], np.float32)
    ref_K, _ = cv2.getOptimalNewCameraMatrix(K, d.reshape(4, 1), c["in_size"], 1.0, c["out_size"],
                                             False)
    nK = oracle.optimal_new_camera_matrix(K, d, c["in_size"], 1.0, c["out_size"])
    assert np.array_equal(nK, ref_K)
    r1, r2 = cv2.initUndistortRectifyMap(K, d.reshape(4, 1), None, ref_K, c["out_size"],
                                         cv2.CV_16SC2)
    m1, m2 = oracle.init_undistort_rectify_map(K, d, nK, c["out_size"])
    assert np.array_equal(m1, r1) and np.array_equal(m2, r2)
    img = make_image(c)
    und = cv2.remap(img, r1, r2, cv2.INTER_LINEAR)
    assert np.array_equal(oracle.remap_bilinear(img, m1, m2), und)
    # white noise exercises every weight; the maps reach outside the source (constant border)
    noise = np.random.default_rng(1).integers(0, 256, img.shape, dtype=np.uint8)
    assert np.array_equal(oracle.remap_bilinear(noise, m1, m2),
                          cv2.remap(noise, r1, r2, cv2.INTER_LINEAR))
    assert int(m1.min()) < 0


def test_oracle_optimal_matrix_random_calibrations(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for _ in range(25):
        W, H = int(rng.integers(20, 80)) * 16, int(rng.integers(15, 60)) * 16
        K = np.array([[rng.uniform(0.7, 1.2) * W, 0, W / 2 + rng.normal(0, 8)],
                      [0, rng.uniform(0.7, 1.2) * W, H / 2 + rng.normal(0, 8)], [0, 0, 1]],
                     np.float32)
        # moderate barrel distortion (the model stays monotonic over the image)
        d = np.array([rng.uniform(-0.2, 0.05), rng.uniform(0.0, 0.05), rng.normal(0, 3e-4),
                      rng.normal(0, 3e-4)], np.float32)
        for alpha in (1.0, 0.0, 0.4):
            ref, _ = cv2.getOptimalNewCameraMatrix(K, d.reshape(4, 1), (W, H), alpha, (W - 16, H),
                                                   False)
            assert np.array_equal(oracle.optimal_new_camera_matrix(K, d, (W, H), alpha,
                                                                   (W - 16, H)), ref)


def test_calculate_roi(oracle):
    und = np.zeros((100, 160), np.uint8)
    und[12:90, 20:150] = 9
    assert oracle.calculate_roi(und) == (25, 17, 119, 67)
    und[49, 20:30] = 0    # the scan walks along the middle row / column only
    assert oracle.calculate_roi(und) == (35, 17, 109, 67)
    with pytest.raises(RuntimeError):
        oracle.calculate_roi(np.zeros((10, 10), np.uint8))


@pytest.mark.parametrize("name", list(CASES))
def test_product_host_functions_match_cv2_and_oracle(oracle, name):
    cv2 = pytest.importorskip("cv2")
    import uw_slam_b200 as U
    c = CASES[name]
    cam = U.CameraModel.from_distorted(c["in_size"], c["out_size"], c["fx"], c["fy"], c["cx"],
                                       c["cy"], c["dist"])
    assert cam.IsValid()
    K, d = K_of(c), np.array(c["dist"], np.float32)
    ref_K, _ = cv2.getOptimalNewCameraMatrix(K, d.reshape(4, 1), c["in_size"], 1.0, c["out_size"],
                                             False)
    r1, r2 = cv2.initUndistortRectifyMap(K, d.reshape(4, 1), None, ref_K, c["out_size"],
                                         cv2.CV_16SC2)
    assert np.array_equal(cam.GetK(), ref_K)
    assert np.array_equal(cam.GetMap1(), r1) and np.array_equal(cam.GetMap2(), r2)
    assert np.array_equal(cam.GetK(), oracle.optimal_new_camera_matrix(K, d, c["in_size"], 1.0,
                                                                       c["out_size"]))
    und = cv2.remap(make_image(c), r1, r2, cv2.INTER_LINEAR)
    roi = U.CalculateROI(und)
    assert roi == oracle.calculate_roi(und)
    x, y, w, h = U.AlignROI(roi)
    assert w % 16 == 0 and h % 16 == 0 and w <= roi[2] and h <= roi[3]


def test_reference_calibration_files_parse(tmp_path):
    # the XML schema of calibration/calibrationEUROC.xml (rectify) and calibrationTUM.xml (not)
    import uw_slam_b200 as U
    xml = """<?xml version="1.0"?><opencv_storage>
<in_width type_id="integer"> 752 </in_width><in_height type_id="integer"> 480 </in_height>
<out_width type_id="integer"> 736 </out_width><out_height type_id="integer"> 480 </out_height>
<calibration_values type_id="opencv-matrix"><rows>1</rows><cols>4</cols><dt>f</dt>
<data> 458.654 457.296 367.215 248.375 </data></calibration_values>
<rectification type_id="opencv-matrix"><rows>1</rows><cols>4</cols><dt>f</dt>
<data> {d} </data></rectification></opencv_storage>"""
    p = tmp_path / "c.xml"
    p.write_text(xml.format(d="-0.28340811 0.07395907 0.00019359 1.76187114e-05"))
    cam = U.CameraModel().GetCameraModel(str(p))
    assert cam.IsValid() and cam.GetMap1().shape == (480, 736, 2)
    assert abs(float(cam.GetK()[0, 0]) - 327.32285) < 1e-3
    p.write_text(xml.format(d="0 0 0 0"))
    cam = U.CameraModel().GetCameraModel(str(p))
    assert not cam.IsValid() and cam.GetMap1() is None
    assert float(cam.GetK()[0, 0]) == np.float32(458.654)   # CameraModel.cpp:78-83


def test_fixture_is_current(oracle):
    # the committed fixture equals what the oracle produces today (maps, remap, ROI)
    gold = np.load(GOLD)
    for name, c in CASES.items():
        K, d = K_of(c), np.array(c["dist"], np.float32)
        nK = oracle.optimal_new_camera_matrix(K, d, c["in_size"], 1.0, c["out_size"])
        m1, m2 = oracle.init_undistort_rectify_map(K, d, nK, c["out_size"])
        und = oracle.remap_bilinear(make_image(c), m1, m2)
        assert np.array_equal(nK, gold[name + "_newK"])
        assert [sha(m1), sha(m2), sha(und)] == list(gold[name + "_sha"])
        assert oracle.calculate_roi(und) == tuple(gold[name + "_roi"])


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_undistort_image_matches_fixture_and_oracle(oracle, name):
    import uw_slam_b200 as U
    gold = np.load(GOLD)
    c = CASES[name]
    cam = U.CameraModel.from_distorted(c["in_size"], c["out_size"], c["fx"], c["fy"], c["cx"],
                                       c["cy"], c["dist"])
    assert np.array_equal(cam.GetK(), gold[name + "_newK"])
    assert [sha(cam.GetMap1()), sha(cam.GetMap2())] == list(gold[name + "_sha"][:2])
    img = make_image(c)
    und = cam.Undistort(img)
    assert sha(und) == gold[name + "_sha"][2]
    assert np.array_equal(und, oracle.remap_bilinear(img, cam.GetMap1(), cam.GetMap2()))
    assert U.CalculateROI(und) == tuple(gold[name + "_roi"])
    noise = np.random.default_rng(2).integers(0, 256, img.shape, dtype=np.uint8)
    assert np.array_equal(cam.Undistort(noise),
                          oracle.remap_bilinear(noise, cam.GetMap1(), cam.GetMap2()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["euroc", "small"])
def test_gpu_fused_remap_crop_pyramid_and_track(oracle, name):
    """System::AddFrame + Tracking on DISTORTED frames: remap, ROI crop and pyramid in one kernel,
    then the usual track; everything equal to the oracle run on the oracle-undistorted crop."""
    import uw_slam_b200 as U
    c = CASES[name]
    cam = U.CameraModel.from_distorted(c["in_size"], c["out_size"], c["fx"], c["fy"], c["cx"],
                                       c["cy"], c["dist"])
    iw, ih = c["in_size"]
    # two distorted input frames: a synthetic pair rendered at the input size
    calib = (iw, ih, c["fx"], c["fy"], c["cx"], c["cy"])
    synth.CALIB["_und_" + name] = calib
    prev_d, cur_d, _, _ = synth.render_pair("_und_" + name, 3)
    prev_d, cur_d = np.maximum(prev_d, 1), np.maximum(cur_d, 1)
    x, y, w, h = U.AlignROI(U.CalculateROI(cam.Undistort(prev_d)))
    K = cam.GetK()   # the reference keeps K_ as is after the crop (System.cpp:105-123)
    t = U.Tracker(False)
    t.InitializePyramid(w, h, K, max_frames=2)
    t.SetUndistortion(cam, (x, y))
    fp, fc = t.AddFrames([0, 1], np.stack([prev_d, cur_d]))
    ref_prev = oracle.remap_bilinear(prev_d, cam.GetMap1(), cam.GetMap2())[y:y + h, x:x + w]
    ref_cur = oracle.remap_bilinear(cur_d, cam.GetMap1(), cam.GetMap2())[y:y + h, x:x + w]
    rp, rc = oracle.FrameData(ref_prev), oracle.FrameData(ref_cur, with_candidates=False)
    for lvl in range(5):
        assert np.array_equal(fp.image(lvl), rp.images[lvl]), lvl
        assert np.array_equal(fc.image(lvl), rc.images[lvl]), lvl
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    pose = t.EstimatePose(fp, fc)[0]
    p = oracle.default_params(w, h, float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
    assert np.array_equal(pose, oracle.estimate_pose(p, rp, rc)[0])
    # device-resident distorted frames take the same path
    import ctypes
    t.SetUndistortion(None)
    with pytest.raises(U.UwtError):   # wrong ROI is rejected
        t._check(t._lib.uwt_set_undistortion(
            t._h, cam.GetMap1().ctypes.data_as(U._lib._i16p),
            cam.GetMap2().ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), cam.GetOutputWidth(),
            cam.GetOutputHeight(), iw, ih, cam.GetOutputWidth() - 8, 0))
    # undistortion off again: frames of the cropped size go straight in
    f2 = t.AddFrames([0], ref_prev)[0]
    assert np.array_equal(f2.image(1), rp.images[1])
    t.close()
