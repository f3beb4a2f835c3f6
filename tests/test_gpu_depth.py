"""Depth input: Tracker(depth_available=true), the 16-bit depth pyramid of
System::AddFrame (System.cpp:241-250) and the depth branch of ObtainCandidatePoints
(Tracker.cpp:1338-1347), in both readings of `depths_[lvl].at<uchar>(y,x)`; GPU vs oracle."""
import numpy as np
import pytest

from uw_slam_b200 import synth


def make_depth(shape, seed, holes=True):
    """A smooth 16-bit depth map around 1.5-4 m (TUM factor 5000/m ~ here 1/0.0002), with holes
    (depth 0 = no measurement) and values whose low byte is 0 (matters to the at<uchar> read)."""
    rng = np.random.default_rng(seed)
    h, w = shape
    ys, xs = np.mgrid[0:h, 0:w]
    d = 9000 + 4000 * np.sin(xs * 0.013 + seed) * np.cos(ys * 0.017) + rng.integers(0, 40, shape)
    d = d.astype(np.uint16)
    if holes:
        d[rng.random(shape) < 0.07] = 0
        d[h // 3:h // 3 + 9, w // 4:w // 2] = 0
        d[5::17, 3::11] &= 0xFF00
    return d


def test_depth_pyramid_matches_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    d = make_depth((480, 640), 1)
    for _ in range(4):
        nxt = oracle.depth_pyr_down(d)
        assert np.array_equal(nxt, cv2.resize(d, None, fx=0.5, fy=0.5))   # System.cpp:249
        d = nxt


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("calib,seed,batch", [("small", 3, 1), ("tum", 1, 1), ("euroc", 2, 1),
                                              ("small", 10, 30), ("small", 10, -30)])
def test_depth_tracking_matches_oracle(oracle, calib, seed, batch, mode):
    # batch 30: the dataflow kernel's depth mode; batch -30: the same batch on the cluster kernel
    import uw_slam_b200 as U
    import uw_slam_b200._lib as L
    kernel_flag = L.FLAG_CLUSTER_KERNEL if batch < 0 else 0
    batch = abs(batch)
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    pairs = [synth.render_pair(calib, seed + i)[:2] for i in range(batch)]
    deps = [make_depth((h, w), seed + i) for i in range(batch)]
    t = U.Tracker(True, depth_mode=mode)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2 * batch, flags=L.FLAG_TRACE | kernel_flag)
    fp = t.AddFrames(list(range(batch)), np.stack([p[0] for p in pairs]))
    fc = t.AddFrames(list(range(batch, 2 * batch)), np.stack([p[1] for p in pairs]))
    t.ApplyGradient(fp)
    with pytest.raises(U.UwtError):      # the candidate rule needs the depth frame
        t.ObtainCandidatePoints(fp)
    t.AddDepthFrames([f.slot for f in fp], np.stack(deps))
    t.ObtainCandidatePoints(fp)
    poses, stats = t.EstimatePose(fp, fc, return_stats=True)
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    for i in range(batch):
        rp = oracle.FrameData(pairs[i][0], depth=deps[i], depth_mode=mode)
        rc = oracle.FrameData(pairs[i][1], with_candidates=False)
        if i == 0:
            for lvl in range(5):
                assert np.array_equal(fp[0].depth(lvl), rp.depths[lvl]), lvl
                c = fp[0].candidatePoints(lvl)
                assert c.shape == rp.cand[lvl].shape, (lvl, c.shape, rp.cand[lvl].shape)
                assert np.array_equal(c, rp.cand[lvl]), lvl      # rows [x, y, Z, 1]
            assert 0 < rp.cand[1].shape[0] < oracle.FrameData(pairs[i][0]).cand[1].shape[0]
        opose, ostats, otrace = oracle.estimate_pose(p, rp, rc)
        tr = t.get_trace(i)
        assert [(a.level, a.k, a.n_valid, a.broke, a.sum_r2) for a in tr] == \
            [(b.level, b.k, b.n_valid, b.broke, b.sum_r2) for b in otrace], i
        for a, b in zip(tr, otrace):
            assert np.array_equal(np.array(a.A[:]), np.array(b.A[:])), (i, b.level, b.k)
            assert np.array_equal(np.array(a.delta[:]), np.array(b.delta[:])), (i, b.level, b.k)
            assert np.array_equal(np.array(a.pose[:]), np.array(b.pose[:])), (i, b.level, b.k)
        assert np.array_equal(poses[i], opose), i
    t.close()


def test_all_points_oracle_follows_the_reference_loops(oracle):
    """oracle.all_points_depth against a literal restatement of Tracker.cpp:1264-1300, and the
    property the library relies on: the [0, 0, 1, 0] rows never take part in a sweep, and the
    order of the points does not change any rounded result."""
    rng = np.random.default_rng(5)
    h, w = 24, 40
    d = rng.integers(0, 65536, (h, w)).astype(np.uint16)
    d[rng.random((h, w)) < 0.3] = 0
    for lvl in (0, 2, 4):
        factor_lvl = np.float32(np.float64(np.float32(0.0002)) / 2.0 ** lvl)   # :1266
        rows = []
        for y in range(h):               # Tracker.cpp:1268
            for x in range(w):           # Tracker.cpp:1269
                v = int(d[y, x]) - 65536 if d[y, x] >= 32768 else int(d[y, x])   # at<short>
                if v > 0:
                    rows.append([x, y, np.float32(v) * factor_lvl, 1.0])
                else:
                    rows.append([0.0, 0.0, 1.0, 0.0])
        assert np.array_equal(oracle.all_points_depth(d, lvl), np.array(rows, np.float32))
    calib = "small"
    wd, hd, fx, fy, cx, cy = synth.CALIB[calib]
    a, b = synth.render_pair(calib, 4)[:2]
    dep = make_depth((hd, wd), 4)
    dep[7::13, 5::9] = 0x8000 + 77
    p = oracle.default_params(wd, hd, fx, fy, cx, cy)
    rp = oracle.FrameData(a, depth=dep, depth_mode=oracle.DEPTH_ALL_POINTS)
    rc = oracle.FrameData(b, with_candidates=False)
    pose, st, tr = oracle.estimate_pose(p, rp, rc)
    for lvl in range(5):                 # points only, x-major: what the library keeps
        c = rp.cand[lvl]
        c = c[c[:, 3] == 1.0]
        rp.cand[lvl] = c[np.lexsort((c[:, 1], c[:, 0]))]
    pose2, st2, tr2 = oracle.estimate_pose(p, rp, rc)
    assert np.array_equal(pose, pose2) and len(tr) == len(tr2)
    for x, y in zip(tr, tr2):
        assert (x.level, x.k, x.n_valid, x.sum_r2) == (y.level, y.k, y.n_valid, y.sum_r2)
        assert np.array_equal(np.array(x.A[:]), np.array(y.A[:]))
        assert np.array_equal(np.array(x.b[:]), np.array(y.b[:]))


@pytest.mark.gpu
@pytest.mark.parametrize("calib,seed,batch", [("small", 4, 1), ("tum", 2, 1), ("small", 20, 6)])
def test_all_points_tracking_matches_oracle(oracle, calib, seed, batch):
    """Tracker::ObtainAllPoints (Tracker.cpp:1259-1310) + EstimatePose: every pixel whose depth,
    read as a signed short, is > 0, with Z = depth * 0.0002 / 2^level and no gradient test.  The
    oracle builds the reference's row-major list including its [0, 0, 1, 0] rows; the library
    keeps the points with depth, x-major: same points, and every sweep must give the same bits."""
    import uw_slam_b200 as U
    import uw_slam_b200._lib as L
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    pairs = [synth.render_pair(calib, seed + i)[:2] for i in range(batch)]
    deps = [make_depth((h, w), seed + i) for i in range(batch)]
    for d in deps:
        d[7::13, 5::9] = 0x8000 + 77      # negative as a short: not a point (Tracker.cpp:1273)
    t = U.Tracker(True, depth_mode=L.DEPTH_ALL_POINTS)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2 * batch, flags=L.FLAG_TRACE)
    fp = t.AddFrames(list(range(batch)), np.stack([p[0] for p in pairs]))
    fc = t.AddFrames(list(range(batch, 2 * batch)), np.stack([p[1] for p in pairs]))
    t.ApplyGradient(fp)
    t.AddDepthFrames([f.slot for f in fp], np.stack(deps))
    t.ObtainAllPoints(fp)
    poses, stats = t.EstimatePose(fp, fc, return_stats=True)
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    for i in range(batch):
        rp = oracle.FrameData(pairs[i][0], depth=deps[i], depth_mode=oracle.DEPTH_ALL_POINTS)
        rc = oracle.FrameData(pairs[i][1], with_candidates=False)
        if i == 0:
            for lvl in range(5):
                ref = rp.cand[lvl]
                assert ref.shape[0] == rp.images[lvl].size           # one row per pixel
                ref = ref[ref[:, 3] == 1.0]                          # the rows that are points
                got = fp[0].candidatePoints(lvl)                     # x-major
                assert got.shape == ref.shape, (lvl, got.shape, ref.shape)
                order = np.lexsort((ref[:, 1], ref[:, 0]))           # row-major -> x-major
                assert np.array_equal(got, ref[order]), lvl
                assert list(stats[0].n_points)[lvl] == ref.shape[0] or lvl == 0
        opose, ostats, otrace = oracle.estimate_pose(p, rp, rc)
        tr = t.get_trace(i)
        assert [(a.level, a.k, a.n_valid, a.broke, a.sum_r2) for a in tr] == \
            [(b.level, b.k, b.n_valid, b.broke, b.sum_r2) for b in otrace], i
        for a, b in zip(tr, otrace):
            assert np.array_equal(np.array(a.A[:]), np.array(b.A[:])), (i, b.level, b.k)
            assert np.array_equal(np.array(a.delta[:]), np.array(b.delta[:])), (i, b.level, b.k)
            assert np.array_equal(np.array(a.pose[:]), np.array(b.pose[:])), (i, b.level, b.k)
        assert np.array_equal(poses[i], opose), i
    t.close()


@pytest.mark.gpu
def test_depth_api_rules():
    import uw_slam_b200 as U
    import uw_slam_b200._lib as L
    w, h, fx, fy, cx, cy = synth.CALIB["small"]
    K = U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK()
    t = U.Tracker(False)
    t.InitializePyramid(w, h, K)
    with pytest.raises(U.UwtError):      # mono tracker takes no depth
        t.AddDepthFrames([0], np.ones((h, w), np.uint16))
    t.close()
    for bad in (dict(depth_mode=4), dict(depth_mode=3, weight_mode=1),
                dict(depth_mode=1, weight_mode=1), dict(depth_mode=2, flags=L.FLAG_DMMA_ACCUM)):
        with pytest.raises(U.UwtError):
            U.Tracker(False).InitializePyramid(w, h, K, **bad)
    # all-zero depth: no candidates anywhere, identity pose (ARITHMETIC.md U2)
    t = U.Tracker(True)
    t.InitializePyramid(w, h, K)
    prev, cur = synth.render_pair("small", 0)[:2]
    fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
    t.AddDepthFrames([0], np.zeros((h, w), np.uint16))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    assert all(fp.candidatePoints(l).shape[0] == 0 for l in range(5))
    assert np.allclose(t.EstimatePose(fp, fc)[0], [0, 0, 0, 1, 0, 0, 0])
    t.close()
