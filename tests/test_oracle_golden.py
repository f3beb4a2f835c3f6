"""CPU: the oracle against the committed golden vectors (tests/golden/*.npz, made by
tests/golden/make_golden.py after cv2 validation), and an independent numpy restatement of
every traced Gauss-Newton iteration (a second implementation of Tracker.cpp:414-574)."""
import hashlib
import math
import os

import numpy as np
import pytest

from uw_slam_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(GOLD, "golden_small.npz"))


@pytest.fixture(scope="module")
def big():
    return np.load(os.path.join(GOLD, "golden_big.npz"))


def check_trace(gold, key, trace):
    t = np.array([[x.level, x.k, x.n_valid, x.broke] for x in trace], np.int32)
    assert np.array_equal(t, gold[key + "_lvl_k_nvalid_broke"])
    assert np.array_equal(np.array([x.sum_r2 for x in trace], np.int64), gold[key + "_sum_r2"])
    for name in ("A", "b", "delta", "pose"):
        got = np.array([getattr(x, name)[:] for x in trace], np.float32)
        assert np.array_equal(got, gold[key + "_" + name]), name
    assert np.array_equal(np.array([x.error for x in trace], np.float32), gold[key + "_error"])


@pytest.mark.parametrize("calib,seed", [("tiny", 0), ("tiny", 1), ("small", 0), ("small", 1),
                                        ("small", 2)])
def test_small_golden(oracle, small, calib, seed):
    key = "%s_%d" % (calib, seed)
    prev, cur = small[key + "_prev"], small[key + "_cur"]
    # the generator is deterministic: the committed inputs are what the seed renders
    p2, c2, _, _ = synth.render_pair(calib, seed)
    assert np.array_equal(prev, p2) and np.array_equal(cur, c2)
    fp, fc = oracle.FrameData(prev), oracle.FrameData(cur, with_candidates=False)
    for l in range(5):
        assert np.array_equal(fp.images[l], small["%s_img%d" % (key, l)])
        assert np.array_equal(fp.gx[l], small["%s_gx%d" % (key, l)])
        assert np.array_equal(fp.gy[l], small["%s_gy%d" % (key, l)])
        assert np.array_equal(fp.g[l], small["%s_g%d" % (key, l)])
        assert np.array_equal(fp.cand[l][:, :2].astype(np.uint16), small["%s_cand%d" % (key, l)])
        assert np.all(fp.cand[l][:, 2:] == 1.0)
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    for mode in (0, 1):
        p = oracle.default_params(w, h, fx, fy, cx, cy, solve_mode=mode)
        pose, _, tr = oracle.estimate_pose(p, fp, fc)
        check_trace(small, "%s_m%d" % (key, mode), tr)
        assert np.array_equal(pose, small["%s_m%d_final" % (key, mode)])


@pytest.mark.parametrize("calib,seed", [("tum", 0), ("tum", 3), ("euroc", 1)])
def test_big_golden(oracle, big, calib, seed):
    key = "%s_%d" % (calib, seed)
    prev, cur, _, _ = synth.render_pair(calib, seed)
    assert [sha(prev), sha(cur)] == list(big[key + "_input_sha"])
    fp, fc = oracle.FrameData(prev), oracle.FrameData(cur, with_candidates=False)
    got = [[sha(fp.images[l]), sha(fp.gx[l]), sha(fp.gy[l]), sha(fp.g[l]), sha(fp.cand[l])]
           for l in range(5)]
    assert got == [list(r) for r in big[key + "_level_sha"]]
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    pose, st, tr = oracle.estimate_pose(oracle.default_params(w, h, fx, fy, cx, cy), fp, fc)
    check_trace(big, key, tr)
    assert np.array_equal(pose, big[key + "_final"])
    assert list(st.iterations)[:5] == list(big[key + "_iterations"])


@pytest.fixture(scope="module")
def robust():
    return np.load(os.path.join(GOLD, "golden_robust.npz"))


@pytest.mark.parametrize("calib,seed", [("tiny", 0), ("small", 1), ("small", 2), ("tum", 0),
                                        ("tum", 4)])
@pytest.mark.parametrize("name,mode", [("tukey", 1), ("huber", 2)])
def test_robust_golden(oracle, robust, calib, seed, name, mode):
    key = "%s_%d_%s" % (calib, seed, name)
    prev, cur, _, _ = synth.render_pair(calib, seed)
    assert [sha(prev), sha(cur)] == list(robust[key + "_input_sha"])
    fp, fc = oracle.FrameData(prev), oracle.FrameData(cur, with_candidates=False)
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    p = oracle.default_params(w, h, fx, fy, cx, cy, weight_mode=mode, huber_delta=7.5)
    pose, st, tr = oracle.estimate_pose(p, fp, fc)
    check_trace(robust, key, tr)
    assert np.array_equal(pose, robust[key + "_final"])
    assert list(st.iterations)[:5] == list(robust[key + "_iterations"])


def test_accumulator_modes_agree(oracle):
    # fp64-sequential (timed baseline) and 80-bit (checker) accumulation round to the same f32
    calib = "small"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    prev, cur, _, _ = synth.render_pair(calib, 5)
    fp, fc = oracle.FrameData(prev), oracle.FrameData(cur, with_candidates=False)
    a = oracle.estimate_pose(oracle.default_params(w, h, fx, fy, cx, cy, accum_mode=0), fp, fc)
    b = oracle.estimate_pose(oracle.default_params(w, h, fx, fy, cx, cy, accum_mode=1), fp, fc)
    c = oracle.estimate_pose(oracle.default_params(w, h, fx, fy, cx, cy, threads=3), fp, fc)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[0], c[0])


# ---- independent numpy restatement of one GN iteration ---------------------------------------
f32 = np.float32


def np_quat_to_R(q):
    x, y, z, w = [f32(v) for v in q]
    tx, ty, tz = f32(2) * x, f32(2) * y, f32(2) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = f32(1)
    return np.array([[one - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, one - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, one - (txx + tyy)]], f32)


def np_iteration(pts, I1, I2, gx, gy, pose, K, lvl, weight_fn=None, sqrt_weights=False,
                 bilinear=False):
    """A, b (float32), n_valid, sum_r2 [, weighted error sum] for one sweep; vectorised float32
    numpy, sums by fsum.  weight_fn(residuals f32) -> weights f32 (Tracker.cpp:496)."""
    fx, fy, cx, cy = (f32(K[k][lvl]) for k in ("fx", "fy", "cx", "cy"))
    ifx, ify = f32(K["invfx"][lvl]), f32(K["invfy"][lvl])
    rows, cols = I2.shape
    x1, y1 = pts[:, 0], pts[:, 1]
    Z, W = pts[:, 2].astype(f32), pts[:, 3].astype(f32)   # mono: Z = W = 1 (every product exact)
    X = (((x1 - cx) * ifx).astype(f32) * Z).astype(f32)   # Tracker.cpp:1439-1441
    Y = (((y1 - cy) * ify).astype(f32) * Z).astype(f32)
    R = np_quat_to_R(pose[:4]).astype(np.float64)
    t = pose[4:].astype(np.float64)
    Xd, Yd, Zd, Wd = (v.astype(np.float64) for v in (X, Y, Z, W))
    o = [(R[r, 0] * Xd + (R[r, 1] * Yd + (R[r, 2] * Zd + t[r] * Wd))).astype(f32)
         for r in range(3)]
    with np.errstate(divide="ignore", invalid="ignore"):
        x2 = (((o[0] * fx) / o[2] + cx).astype(f32) * W).astype(f32)   # Tracker.cpp:1460-1467
        y2 = (((o[1] * fy) / o[2] + cy).astype(f32) * W).astype(f32)
    z2 = o[2]
    valid = (y2 > 0) & (y2 < f32(rows)) & (x2 > 0) & (x2 < f32(cols)) & (z2 != 0)
    x1, y1, x2, y2, z2 = x1[valid], y1[valid], x2[valid], y2[valid], z2[valid]
    iz = (f32(1) / z2).astype(f32)
    iz = np.where(iz < 0, f32(0), iz)
    one = f32(1)
    Jw0 = [fx * iz, None, -(fx * x2 * iz * iz), -(fx * x2 * y2 * iz * iz),
           fx * (one + x2 * x2 * iz * iz), -fx * y2 * iz]
    Jw1 = [None, fy * iz, -(fy * y2 * iz * iz), -(fy * (one + y2 * y2 * iz * iz)),
           fy * x2 * y2 * iz * iz, fy * x2 * iz]
    rnd = lambda v: np.where(v - np.floor(v) >= 0.5, np.floor(v) + 1, np.floor(v)).astype(np.int64)
    xi = np.minimum(rnd(x2), cols - 1)
    yi = np.minimum(rnd(y2), rows - 1)
    xs, ys = x1.astype(np.int64), y1.astype(np.int64)
    r = I2[yi, xi].astype(np.int64) - I1[ys, xs].astype(np.int64)
    jx = gx[ys, xs].astype(np.float64)
    jy = gy[ys, xs].astype(np.float64)
    J = np.zeros((x1.size, 6), np.float64)
    for c in range(6):
        a = Jw0[c].astype(np.float64) if Jw0[c] is not None else 0.0
        b = Jw1[c].astype(np.float64) if Jw1[c] is not None else 0.0
        J[:, c] = (jx * a + jy * b).astype(f32).astype(np.float64)
    r50 = (r.astype(f32) * f32(50)).astype(np.float64)
    esum = None
    if bilinear:   # ARITHMETIC.md B1: float interpolation of the four neighbours of (x2, y2)
        ix, iy = x2.astype(np.int64), y2.astype(np.int64)
        ax, ay = (x2 - ix.astype(f32)).astype(f32), (y2 - iy.astype(f32)).astype(f32)
        ix1, iy1 = np.minimum(ix + 1, cols - 1), np.minimum(iy + 1, rows - 1)
        a, b_ = I2[iy, ix].astype(f32), I2[iy, ix1].astype(f32)
        c_, d = I2[iy1, ix].astype(f32), I2[iy1, ix1].astype(f32)
        top = (a + (ax * (b_ - a)).astype(f32)).astype(f32)
        bot = (c_ + (ax * (d - c_)).astype(f32)).astype(f32)
        v = (top + (ay * (bot - top)).astype(f32)).astype(f32)
        rf = (v - I1[ys, xs].astype(f32)).astype(f32)
        r50 = (rf * f32(50)).astype(f32).astype(np.float64)
        esum = math.fsum(rf.astype(np.float64) ** 2)
        r = np.zeros_like(r)           # the integer side-sum is reported as 0
    if weight_fn is not None:
        rf = r.astype(f32)
        wgt = weight_fn(rf).astype(f32)
        esum = math.fsum(rf.astype(np.float64) * (rf * wgt).astype(np.float64))  # Tracker.cpp:500
        sc = np.sqrt(wgt).astype(f32) if sqrt_weights else wgt
        J = (sc[:, None] * J.astype(f32)).astype(f32).astype(np.float64)    # Tracker.cpp:554-557
        r50 = ((rf * f32(50)) * sc).astype(f32).astype(np.float64)          # Tracker.cpp:559,562
    A = np.zeros((6, 6), f32)
    b = np.zeros(6, f32)
    for i in range(6):
        for j in range(i, 6):
            A[i, j] = A[j, i] = f32(math.fsum(J[:, i] * J[:, j]))   # exactly rounded sums
        b[i] = f32(-math.fsum(J[:, i] * r50))
    if weight_fn is not None or bilinear:
        return A, b, int(valid.sum()), int((r * r).sum()), esum
    return A, b, int(valid.sum()), int((r * r).sum())


@pytest.mark.parametrize("calib,seed,mode,depth_mode", [
    ("small", 0, 0, 0), ("tum", 2, 0, 0), ("small", 1, 1, 0), ("tum", 2, 1, 0), ("small", 1, 2, 0),
    ("small", 3, 0, 1), ("small", 3, 0, 2), ("small", 4, 0, 3), ("tum", 2, 0, 3),
    ("small", 2, 0, -1), ("tum", 3, 0, -1)])        # depth_mode -1: mono, bilinear sampling (B1)
def test_every_traced_iteration_matches_numpy_restatement(oracle, calib, seed, mode, depth_mode):
    bilinear = depth_mode < 0
    depth_mode = max(depth_mode, 0)
    cv2 = pytest.importorskip("cv2")
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    prev, cur, _, _ = synth.render_pair(calib, seed)
    depth = None
    if depth_mode:   # per-point depth (ARITHMETIC.md D2-D4): [x, y, Z, W] rows from the oracle
        from test_gpu_depth import make_depth
        depth = make_depth((h, w), seed)
        depth[7::13, 5::9] = 0x8000 + 77
    fp = oracle.FrameData(prev, depth=depth, depth_mode=depth_mode)
    fc = oracle.FrameData(cur, with_candidates=False)
    p = oracle.default_params(w, h, fx, fy, cx, cy, weight_mode=mode, huber_delta=7.5,
                              sampling=1 if bilinear else 0)
    _, _, tr = oracle.estimate_pose(p, fp, fc)
    weight_fn = None
    if mode == 1:   # Tukey / MAD evaluated with the real OpenCV (tests/test_oracle_vs_cv2.py)
        from test_oracle_vs_cv2 import cv_tukey
        weight_fn = lambda r: cv_tukey(r)[0]
    elif mode == 2:
        def weight_fn(r):
            a = np.abs(r)
            with np.errstate(divide="ignore"):
                return np.where(a <= f32(7.5), f32(1), f32(7.5) / a).astype(f32)
    K = oracle.init_pyramid(w, h, fx, fy, cx, cy, 5)
    pose = np.array([0, 0, 0, 1, 0, 0, 0], f32)
    checked = 0
    for i, t in enumerate(tr):
        if t.k == 0 and i > 0:  # new level: previous pose went through the level transition
            pose = oracle.se3_scale_level(pose)
        lvl = t.level
        res = np_iteration(fp.cand[lvl], fp.images[lvl], fc.images[lvl], fp.gx[lvl],
                           fp.gy[lvl], pose, K, lvl, weight_fn, sqrt_weights=(mode == 2),
                           bilinear=bilinear)
        A, b, nv, sr2 = res[:4]
        assert (nv, sr2) == (t.n_valid, t.sum_r2), (lvl, t.k)
        err = f32(np.float64(f32(1.0 / nv)) *
                  np.float64(sr2 if (mode == 0 and not bilinear) else res[4]))
        assert err == f32(t.error)
        if not t.broke:
            assert np.array_equal(A, np.array(t.A[:], f32).reshape(6, 6)), (lvl, t.k)
            assert np.array_equal(b, np.array(t.b[:], f32)), (lvl, t.k)
            ok, delta = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_LU)  # the real OpenCV
            assert np.array_equal(delta.ravel(), np.array(t.delta[:], f32))
            pose = oracle.se3_mul(pose, oracle.se3_exp(delta.ravel()))
            checked += 1
        assert np.array_equal(pose, np.array(t.pose[:], f32))
    assert checked >= 3


def test_se3_pieces_against_scipy(oracle):
    # sanity of the Sophus restatement against an independent fp64 implementation
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(3)
    for _ in range(50):
        a = (rng.normal(size=6) * [0.05, 0.05, 0.05, 0.3, 0.3, 0.3]).astype(f32)
        T = oracle.se3_matrix(oracle.se3_exp(a)).astype(np.float64)
        R = Rot.from_rotvec(a[3:].astype(np.float64)).as_matrix()
        assert np.allclose(T[:3, :3], R, atol=3e-7)
        th = np.linalg.norm(a[3:].astype(np.float64))
        O = np.array([[0, -a[5], a[4]], [a[5], 0, -a[3]], [-a[4], a[3], 0]], np.float64)
        V = np.eye(3) + (1 - math.cos(th)) / th**2 * O + (th - math.sin(th)) / th**3 * O @ O
        assert np.allclose(T[:3, 3], V @ a[:3].astype(np.float64), atol=1e-6)
    p = oracle.se3_exp(np.array([0.01, 0.02, 0.03, 0.1, -0.2, 0.05], f32))
    q = oracle.se3_exp(np.array([-0.02, 0.01, 0.0, -0.05, 0.1, 0.2], f32))
    M = oracle.se3_matrix(oracle.se3_mul(p, q)).astype(np.float64)
    assert np.allclose(M, oracle.se3_matrix(p).astype(np.float64) @
                       oracle.se3_matrix(q).astype(np.float64), atol=1e-6)
    # identity through the tiny-angle branch, and the level transition doubles small angles
    assert np.array_equal(oracle.se3_exp(np.zeros(6, f32)), np.array([0, 0, 0, 1, 0, 0, 0], f32))
    s = oracle.se3_scale_level(p)
    assert abs(np.linalg.norm(s[:4]) - 1) < 1e-6 and np.array_equal(s[4:], p[4:])
