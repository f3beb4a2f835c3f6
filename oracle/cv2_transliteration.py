"""Tracker::EstimatePose transliterated onto the REAL OpenCV (Python cv2), call for call.

TEST INFRASTRUCTURE ONLY (tests/ and the fixture generator import it; never the product).

The reference cannot be compiled here (OpenCV 3.2 C++ / Eigen / Ceres / ROS are absent), so the
oracle (oracle/uw_oracle.cpp) fixes a canonical arithmetic where OpenCV's own is build-dependent:
rule U3 rounds the fp64 sums of J^T J, J^T r and r^T r once to f32, whereas the reference runs
`Jacobians.t() * Jacobians`, `-Jacobians.t() * Residuals` and `A.inv() * b` through cv::gemm /
cv::invert on CV_32F matrices (src/Tracker.cpp:500-501, 559-564), i.e. whatever float
accumulation the installed OpenCV build uses.  This file runs exactly those cv2 calls
(cv2.gemm with GEMM_1_T, cv2.invert DECOMP_LU, cv2.gemm for the 4x4 * 4xN warp of
WarpFunction, src/Tracker.cpp:1417-1471, cv2.multiply / divide / add for its row operations)
so that the DISTANCE between "what a real OpenCV produces" and the oracle can be measured and
committed (tests/golden/cv2_distance.json).  It is the only reference-side number obtainable in
this container: cv2 4.13 instead of the reference's 3.2, Sophus pieces from the oracle.

Per-element float expressions (Jw rows, src/Tracker.cpp:455-467) are evaluated with numpy
float32 arrays in the source's left-to-right order (one rounding per operation, as the C++
float code does on x86-64/SSE); `Jl * Jw` (a 1x2 by 2x6 cv::gemm per point, :479) is
evaluated in its vectorised form, checked against real per-point cv2.gemm calls on a sample.
"""
import numpy as np

f32 = np.float32


def warp_function(cv2, pts4, pose_matrix44, fx, fy, cx, cy, invfx, invfy, folded=False):
    """Tracker::WarpFunction (src/Tracker.cpp:1417-1471).  `folded`: evaluate
    `(col - cx) * invfx` the way cv::MatExpr folds it, col * alpha + beta in float, instead of
    the literal two-step order (docs/ARITHMETIC.md U4 takes the literal order)."""
    P = np.ascontiguousarray(pts4, f32).copy()
    for c, (cc, inv) in enumerate(((cx, invfx), (cy, invfy))):
        col = np.ascontiguousarray(P[:, c:c + 1])
        if folded:
            col = (col * f32(inv) + f32(f32(-cc) * f32(inv))).astype(f32)
        else:
            col = cv2.multiply(cv2.subtract(col, float(cc)), float(inv))
        P[:, c:c + 1] = cv2.multiply(col.reshape(-1, 1), np.ascontiguousarray(P[:, 2:3]))
    # projected_points = rigid * projected_points.t()
    Q = cv2.gemm(np.ascontiguousarray(pose_matrix44, f32), np.ascontiguousarray(P.T), 1.0, None,
                 0.0)
    row2 = np.ascontiguousarray(Q[2:3, :])
    for r, (f, c) in enumerate(((fx, cx), (fy, cy))):
        row = cv2.multiply(np.ascontiguousarray(Q[r:r + 1, :]), float(f))   # row *= f
        row = cv2.divide(row, row2)                                        # row /= row(2)
        row = cv2.add(row, float(c))                                       # row += c
        Q[r:r + 1, :] = cv2.multiply(row, np.ascontiguousarray(Q[3:4, :]))  # .mul(row(3))
    return np.ascontiguousarray(Q.T)


def jacobian_rows(x2, y2, iz, fx, fy, gx, gy):
    """Jw (src/Tracker.cpp:455-467, z_factor = angle_factor = 1) and Jl * Jw (:476-479)."""
    fx, fy, one = f32(fx), f32(fy), f32(1)
    w00 = fx * iz
    w02 = -(fx * x2 * iz * iz)
    w03 = -(fx * x2 * y2 * iz * iz)
    w04 = fx * (one + x2 * x2 * iz * iz)
    w05 = -fx * y2 * iz
    w11 = fy * iz
    w12 = -(fy * y2 * iz * iz)
    w13 = -(fy * (one + y2 * y2 * iz * iz))
    w14 = fy * x2 * y2 * iz * iz
    w15 = fy * x2 * iz
    zero = np.zeros_like(w00)
    Jw0 = np.stack([w00, zero, w02, w03, w04, w05], 1)
    Jw1 = np.stack([zero, w11, w12, w13, w14, w15], 1)
    # cv::gemm of a 1x2 by a 2x6 float matrix: double accumulator, one rounding
    J = (gx[:, None].astype(np.float64) * Jw0.astype(np.float64) +
         gy[:, None].astype(np.float64) * Jw1.astype(np.float64)).astype(f32)
    return J, Jw0, Jw1


def estimate_pose(cv2, O, params, prev, cur, folded=False, gemm_check=64, rng=None):
    """The loop of src/Tracker.cpp:362-597 on two oracle FrameData (their images, gradients and
    candidate lists are cv2-verified).  Returns (pose7, iterations per level, per-sweep records
    [(level, k, n_valid, error, A, b, delta)])."""
    w, h = params.width, params.height
    K = O.init_pyramid(w, h, params.fx, params.fy, params.cx, params.cy, params.levels)
    pose = O.se3_exp(np.zeros(6, f32))            # SE3(SO3::exp(0), 0), Tracker.cpp:385
    iters = [0] * params.levels
    trace = []
    rng = rng or np.random.default_rng(0)
    for lvl in range(params.first_level, params.last_level - 1, -1):
        last_error = f32(50000.0)
        img1, img2 = prev.images[lvl], cur.images[lvl]
        rows, cols = img2.shape
        pts = np.ascontiguousarray(prev.cand[lvl], f32)
        gX, gY = prev.gx[lvl], prev.gy[lvl]
        fx, fy, cx, cy = (K[k][lvl] for k in ("fx", "fy", "cx", "cy"))
        for k in range(params.max_iterations):
            if pts.shape[0] == 0:
                break                               # ARITHMETIC.md U2
            wp = warp_function(cv2, pts, O.se3_matrix(pose), fx, fy, cx, cy, K["invfx"][lvl],
                               K["invfy"][lvl], folded)
            x1, y1 = pts[:, 0].astype(np.int64), pts[:, 1].astype(np.int64)
            x2, y2, z2 = wp[:, 0], wp[:, 1], wp[:, 2]
            with np.errstate(divide="ignore", invalid="ignore"):
                iz = (f32(1) / z2).astype(f32)
            valid = (y2 > 0) & (y2 < rows) & (x2 > 0) & (x2 < cols) & (z2 != 0)
            if not valid.any():
                break                               # ARITHMETIC.md U2
            iz = np.where(iz < 0, f32(0), iz)[valid]
            x2v, y2v = x2[valid], y2[valid]
            # image2.at<uchar>(round(y2), round(x2)): C round() = half away from zero; the index
            # is clamped to the image like the oracle does (ARITHMETIC.md U1)
            xi = np.minimum(np.floor(x2v.astype(np.float64) + 0.5).astype(np.int64), cols - 1)
            yi = np.minimum(np.floor(y2v.astype(np.float64) + 0.5).astype(np.int64), rows - 1)
            r = (img2[yi, xi].astype(np.int64) - img1[y1[valid], x1[valid]].astype(np.int64))
            R = r.astype(f32).reshape(-1, 1)
            gx = gX[y1[valid], x1[valid]].astype(f32)
            gy = gY[y1[valid], x1[valid]].astype(f32)
            J, Jw0, Jw1 = jacobian_rows(x2v, y2v, iz, fx, fy, gx, gy)
            if gemm_check:                          # real per-point cv2.gemm on a sample
                for i in rng.integers(0, J.shape[0], min(gemm_check, J.shape[0])):
                    Jl = np.array([[gx[i], gy[i]]], f32)
                    Jw = np.ascontiguousarray(np.stack([Jw0[i], Jw1[i]]), f32)
                    assert np.array_equal(cv2.gemm(Jl, Jw, 1.0, None, 0.0).ravel(), J[i])
            n = R.shape[0]
            W = np.ones((n, 1), f32)                # IdentityWeights, Tracker.cpp:495
            RW = cv2.multiply(R, W)
            inv_n = f32(1.0 / n)
            # errorMat = inv_num_residuals * Residuals.t() * ResidualsW  -> one gemm with alpha
            err = cv2.gemm(R, RW, float(inv_n), None, 0.0, flags=cv2.GEMM_1_T)[0, 0]
            iters[lvl] = k
            rec = [lvl, k, int(n), f32(err), None, None, None]
            trace.append(rec)
            if err >= last_error or k == params.max_iterations - 1 or \
                    abs(f32(err - last_error)) < f32(params.epsilon):
                break
            last_error = f32(err)
            Jn = cv2.multiply(J, np.ones_like(J))   # Jacobians.row(i) = wi * row, wi = 1
            R50 = cv2.multiply(R, float(params.residual_scale))
            A = cv2.gemm(Jn, Jn, 1.0, None, 0.0, flags=cv2.GEMM_1_T)
            b = cv2.gemm(Jn, cv2.multiply(R50, W), -1.0, None, 0.0, flags=cv2.GEMM_1_T)
            _, Ainv = cv2.invert(A, flags=cv2.DECOMP_LU)
            delta = cv2.gemm(Ainv, b, 1.0, None, 0.0)
            rec[4:] = [A.copy(), b.ravel().copy(), delta.ravel().copy()]
            pose = O.se3_mul(pose, O.se3_exp(delta.ravel()))
        if lvl != 0:
            pose = O.se3_scale_level(pose)          # Tracker.cpp:580-590
    return pose, iters, trace


def pose_distance(a, b):
    """(rotation angle between the two quaternions in rad, |dt| / |t_b|, |dt|)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    dot = abs(float(np.dot(a[:4], b[:4]))) / (np.linalg.norm(a[:4]) * np.linalg.norm(b[:4]))
    ang = 2.0 * np.arccos(min(1.0, dot))
    dt = float(np.linalg.norm(a[4:] - b[4:]))
    return ang, dt / max(float(np.linalg.norm(b[4:])), 1e-30), dt
