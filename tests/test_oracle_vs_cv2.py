"""Pins the CPU oracle's OpenCV primitives bit-for-bit against the real OpenCV (cv2).

The reference has no tests or golden vectors, and its OpenCV 3.2
C++ build cannot be reproduced here; python cv2 runs the same core/imgproc algorithms, so
every primitive on the hot path is checked against it.  CPU only.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

SIZES = [(64, 48), (160, 128), (640, 480), (752, 480), (1280, 1024)]


def rand_img(rng, w, h, smooth=False):
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    if smooth:
        img = cv2.GaussianBlur(img, (0, 0), 2.0)
    return img


@pytest.mark.parametrize("w,h", SIZES + [(3840, 2160)])
def test_pyr_down_matches_cv_resize(oracle, w, h):
    # System.cpp:247: resize(prev, next, Size(), 0.5, 0.5) (default INTER_LINEAR)
    rng = np.random.default_rng(w * 7 + h)
    img = rand_img(rng, w, h)
    ref = cv2.resize(img, None, fx=0.5, fy=0.5)
    assert np.array_equal(oracle.pyr_down(img), ref)


@pytest.mark.parametrize("w,h", SIZES)
def test_scharr_matches_cv(oracle, w, h):
    # Tracker.cpp:1133-1134
    rng = np.random.default_rng(w + h)
    img = rand_img(rng, w, h)
    gx, gy = oracle.scharr(img)
    assert np.array_equal(gx, cv2.Scharr(img, cv2.CV_16S, 1, 0, scale=1, delta=0,
                                         borderType=cv2.BORDER_DEFAULT))
    assert np.array_equal(gy, cv2.Scharr(img, cv2.CV_16S, 0, 1, scale=1, delta=0,
                                         borderType=cv2.BORDER_DEFAULT))


def test_gradmag_exhaustive(oracle):
    # Tracker.cpp:1139-1142: convertScaleAbs x2 then addWeighted(.5,.5,0); all |g| pairs
    v = np.arange(-300, 301, dtype=np.int16)
    gx, gy = np.meshgrid(v, v)
    gx = np.ascontiguousarray(gx)
    gy = np.ascontiguousarray(gy)
    ref = cv2.addWeighted(cv2.convertScaleAbs(gx, alpha=1.0, beta=0.0), 0.5,
                          cv2.convertScaleAbs(gy, alpha=1.0, beta=0.0), 0.5, 0)
    assert np.array_equal(oracle.gradmag(gx, gy), ref)
    big = np.array([[4080, -4080, 255, -256]], np.int16)
    ref = cv2.addWeighted(cv2.convertScaleAbs(big), 0.5, cv2.convertScaleAbs(big[:, ::-1].copy()),
                          0.5, 0)
    assert np.array_equal(oracle.gradmag(big, big[:, ::-1].copy()), ref)


@pytest.mark.parametrize("w,h", SIZES[:4])
def test_candidates_match_cv(oracle, w, h):
    # Tracker.cpp:1325-1357: meanStdDev, thres = mean + 20 (float), threshold, x-major scan
    rng = np.random.default_rng(w * 3 + h)
    img = rand_img(rng, w, h, smooth=True)
    gx, gy = oracle.scharr(img)
    g = oracle.gradmag(gx, gy)
    pts, mean, ithr = oracle.candidates(g, 20.0)
    m, _ = cv2.meanStdDev(g)
    thres = np.float32(m[0, 0] + 20.0)
    _, filt = cv2.threshold(g, float(thres), 255, cv2.THRESH_BINARY)
    xs, ys = np.nonzero(filt.T)  # transposed: x outer, y inner
    ref = np.stack([xs, ys, np.ones_like(xs), np.ones_like(xs)], 1).astype(np.float32)
    assert abs(mean - m[0, 0]) <= 1e-12 * max(1.0, mean)
    assert pts.shape == ref.shape and np.array_equal(pts, ref)
    assert 0 < pts.shape[0] < w * h


def test_threshold_floors_float_threshold(oracle):
    g = np.arange(256, dtype=np.uint8).reshape(16, 16)
    for thr in [20.0, 20.4, 20.999, 147.5 - 20.0]:
        _, filt = cv2.threshold(g, float(np.float32(127.5 + thr)), 255, cv2.THRESH_BINARY)
        pts, mean, ithr = oracle.candidates(g, thr)
        assert mean == 127.5
        assert pts.shape[0] == int(np.count_nonzero(filt))


def cond_matrix(rng):
    # J^T J-like SPD matrices with the wide dynamic range of the tracker's normal equations
    J = rng.normal(size=(400, 6)) * np.array([50, 50, 3e3, 4e5, 4e5, 2e4])
    A = (J.T @ J).astype(np.float32)
    b = (J.T @ rng.normal(size=400) * 50).astype(np.float32)
    return A, b


def test_lu_solve_matches_cv_solve(oracle):
    # Tracker.cpp:564: A.inv()*b is folded by cv::MatExpr into cv::solve(A,b,DECOMP_LU)
    rng = np.random.default_rng(5)
    for i in range(300):
        A, b = cond_matrix(rng) if i % 2 else (rng.normal(size=(6, 6)).astype(np.float32),
                                               rng.normal(size=6).astype(np.float32))
        ok, ref = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_LU)
        x, ok2 = oracle.lu_solve6(A, b)
        assert bool(ok) == bool(ok2)
        assert np.array_equal(x, ref.ravel()), i


def test_lu_invert_matches_cv_invert(oracle):
    rng = np.random.default_rng(6)
    for i in range(300):
        A = cond_matrix(rng)[0] if i % 2 else rng.normal(size=(6, 6)).astype(np.float32)
        _, ref = cv2.invert(A, flags=cv2.DECOMP_LU)
        Ai, ok = oracle.lu_invert6(A)
        assert ok == 1
        assert np.array_equal(Ai, ref), i


def test_lu_singular_gives_zero(oracle):
    A = np.zeros((6, 6), np.float32)
    A[0, 0] = 1.0
    b = np.ones(6, np.float32)
    ok, ref = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_LU)
    x, ok2 = oracle.lu_solve6(A, b)
    assert not ok and ok2 == 0 and np.all(x == 0) and np.all(ref == 0)
    Ai, ok3 = oracle.lu_invert6(A)
    r, refi = cv2.invert(A, flags=cv2.DECOMP_LU)
    assert ok3 == 0 and r == 0.0 and np.all(Ai == 0) and np.all(refi == 0)


def test_warp_matches_cv_gemm(oracle):
    # Tracker.cpp:1417-1471 with the 4x4 * (N x 4)^T product done by the real cv2.gemm
    rng = np.random.default_rng(11)
    K = oracle.init_pyramid(640, 480, 525.0, 525.0, 319.5, 239.5, 5)
    for lvl in [1, 4]:
        w, h = K["w"][lvl], K["h"][lvl]
        n = 20000
        pts = np.ones((n, 4), np.float32)
        pts[:, 0] = rng.integers(0, w, n)
        pts[:, 1] = rng.integers(0, h, n)
        pose = oracle.se3_exp(np.array([4e-3, -3e-3, 2e-3, 2e-3, -3e-3, 4e-3], np.float32))
        fx, fy, cx, cy = (K[k][lvl] for k in ("fx", "fy", "cx", "cy"))
        ifx, ify = K["invfx"][lvl], K["invfy"][lvl]
        out = oracle.warp(pts, pose, fx, fy, cx, cy, ifx, ify)
        P = pts.copy()
        P[:, 0] = ((P[:, 0] - cx) * ifx) * P[:, 2]
        P[:, 1] = ((P[:, 1] - cy) * ify) * P[:, 2]
        T = oracle.se3_matrix(pose)
        D = cv2.gemm(T, P, 1.0, None, 0.0, flags=cv2.GEMM_2_T)
        D[0] = cv2.divide(D[0] * fx, D[2]).ravel() + cx
        D[1] = cv2.divide(D[1] * fy, D[2]).ravel() + cy
        D[0] *= D[3]
        D[1] *= D[3]
        assert np.array_equal(out, D.T)


def test_init_pyramid_matches_source_formulas(oracle):
    # Tracker.cpp:297-340 evaluated with numpy scalars of the same widths
    for (w, h, fx, fy, cx, cy) in [(640, 480, 525.0, 525.0, 319.5, 239.5),
                                   (752, 480, 458.654, 457.296, 367.215, 248.375),
                                   (1280, 1024, 685.72, 685.64, 630.86, 511.92)]:
        K = oracle.init_pyramid(w, h, fx, fy, cx, cy, 5)
        f = np.float32
        efx, ecx = f(fx), f(cx)
        for l in range(1, 5):
            efx = f(np.float64(efx) * 0.5)
            ecx = f((np.float64(f(cx)) + 0.5) / (1 << l) - 0.5)
            assert K["fx"][l] == efx and K["cx"][l] == ecx
            assert K["invfx"][l] == f(1) / efx
            assert K["w"][l] == w >> l and K["h"][l] == h >> l


# ---- robust weights (Tracker.cpp:1571-1594, 1607-1654) --------------
def cv_median_mat(v):
    """Tracker::MedianMat with the real OpenCV: convertTo(CV_8UC1) + calcHist + the loop."""
    v = np.asarray(v, np.float32).reshape(-1, 1)
    # Mat::convertTo(CV_8UC1) == saturate_cast<uchar>(cvRound(x)); cv2.add with dtype runs the
    # same saturating conversion kernel
    ch = cv2.add(v, np.zeros_like(v), dtype=cv2.CV_8U)
    hist = cv2.calcHist([ch], [0], None, [256], [0, 256])
    m = np.float32((ch.shape[0] * ch.shape[1]) // 2)
    bin_, med = 0, -1.0
    for i in range(256):
        if med >= 0.0:
            break
        bin_ += int(np.rint(hist[i, 0]))
        if np.float32(bin_) > m and med < 0.0:
            med = float(i)
    return med


def cv_tukey(v):
    v = np.asarray(v, np.float32)
    med = np.float32(cv_median_mat(v))
    dev = cv2.absdiff(v.reshape(-1, 1), np.full((v.size, 1), med, np.float32)).ravel()
    MAD = raw_mad = np.float32(1.4826) * np.float32(cv_median_mat(dev))
    if MAD == 0:
        MAD = np.float32(1)
    b = np.float32(4.6851)
    inv_MAD = np.float32(1.0 / np.float64(MAD))
    inv_b2 = np.float32(1.0 / np.float64(b * b))
    x = v * inv_MAD
    t = (1.0 - ((x * x) * inv_b2).astype(np.float64)).astype(np.float32)
    return np.where(np.abs(x) <= b, t * t, np.float32(0)).astype(np.float32), raw_mad


@pytest.mark.parametrize("scale", [0.0, 1.0, 4.0, 25.0, 90.0, 300.0])
def test_median_mad_tukey_match_cv(oracle, scale):
    rng = np.random.default_rng(int(scale * 10) + 1)
    for trial in range(12):
        n = int(rng.integers(1, 6000))
        v = np.clip(np.rint(rng.normal(rng.normal(0, 3), scale, n)), -255, 255).astype(np.float32)
        if trial == 0:
            v[:] = 3.0  # MAD == 0 -> replaced by 1 (Tracker.cpp:1634-1637)
        assert oracle.median_mat(v) == cv_median_mat(v)
        w_ref, mad_ref = cv_tukey(v)
        assert np.float32(oracle.mad(v)) == mad_ref
        assert np.array_equal(oracle.tukey_weights(v), w_ref)


def test_median_mat_saturation_and_ties(oracle):
    # negatives clamp to 0, .5 rounds to even, > 255 clamps (saturate_cast<uchar>(cvRound))
    v = np.array([-7, -0.5, 0.5, 1.5, 2.5, 254.5, 255.5, 400], np.float32)
    assert oracle.median_mat(v) == cv_median_mat(v)
    for n in (1, 2, 3, 4, 5, 8, 9):
        v = np.arange(n, dtype=np.float32) * 3
        assert oracle.median_mat(v) == cv_median_mat(v)


def test_huber_weights(oracle):
    v = np.arange(-255, 256, dtype=np.float32)
    for d in (0.5, 7.5, 10.0, 300.0):
        a = np.abs(v)
        with np.errstate(divide="ignore"):
            ref = np.where(a <= np.float32(d), np.float32(1), np.float32(d) / a).astype(np.float32)
        assert np.array_equal(oracle.huber_weights(v, d), ref)


def test_cholesky_lm_solver(oracle):
    # north-star solver option (ARITHMETIC.md S2): against an fp64 solve of the damped system,
    # and against cv2.solve(DECOMP_CHOLESKY) for lambda = 0
    rng = np.random.default_rng(8)
    for _ in range(50):
        J = rng.normal(size=(40, 6)).astype(np.float32)
        A = (J.T @ J).astype(np.float32)
        b = rng.normal(size=6).astype(np.float32)
        for lam in (0.0, 0.2, 3.0):
            x, ok = oracle.cholesky_lm_solve6(A, b, lam)
            assert ok == 1
            Ad = A.astype(np.float64) + lam * np.diag(np.diag(A).astype(np.float64))
            ref = np.linalg.solve(Ad, b.astype(np.float64))
            assert np.abs(x - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())
        ok_cv, xc = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_CHOLESKY)
        x0, _ = oracle.cholesky_lm_solve6(A, b, 0.0)
        assert ok_cv and np.abs(x0 - xc.ravel()).max() <= 1e-4 * max(1.0, np.abs(xc).max())
    x, ok = oracle.cholesky_lm_solve6(-np.eye(6, dtype=np.float32), np.ones(6, np.float32), 0.2)
    assert ok == 0 and not x.any()


@pytest.mark.parametrize("w,h", SIZES[:3])
def test_sobel_option_matches_cv(oracle, w, h):
    # north-star wording (Sobel); the reference itself uses Scharr (Tracker.cpp:1133-1134)
    img = rand_img(np.random.default_rng(w * 3 + h), w, h)
    gx, gy = oracle.sobel(img)
    assert np.array_equal(gx, cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=3))
    assert np.array_equal(gy, cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=3))
