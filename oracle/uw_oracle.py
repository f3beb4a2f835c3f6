"""ctypes binding of the CPU oracle (oracle/libuw_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (uw_slam_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libuw_oracle.so")
MAX_LEVELS = 8
SOLVE_LU, SOLVE_INVERSE, SOLVE_CHOLESKY_LM = 0, 1, 2
ACCUM_DOUBLE, ACCUM_LONGDOUBLE = 0, 1
WEIGHT_IDENTITY, WEIGHT_TUKEY, WEIGHT_HUBER = 0, 1, 2
DEPTH_NONE, DEPTH_REFERENCE, DEPTH_U16, DEPTH_ALL_POINTS = 0, 1, 2, 3


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("levels", C.c_int), ("first_level", C.c_int), ("last_level", C.c_int),
        ("max_iterations", C.c_int), ("epsilon", C.c_float), ("residual_scale", C.c_float),
        ("gradient_threshold", C.c_double),
        ("solve_mode", C.c_int), ("accum_mode", C.c_int), ("threads", C.c_int),
        ("weight_mode", C.c_int), ("huber_delta", C.c_float), ("lm_lambda", C.c_float),
        ("sampling", C.c_int),
    ]


class IterTrace(C.Structure):
    _fields_ = [
        ("level", C.c_int), ("k", C.c_int), ("n_valid", C.c_int), ("broke", C.c_int),
        ("sum_r2", C.c_longlong), ("error", C.c_float),
        ("A", C.c_float * 36), ("b", C.c_float * 6), ("delta", C.c_float * 6),
        ("pose", C.c_float * 7),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("iterations", C.c_int * MAX_LEVELS), ("evaluations", C.c_int * MAX_LEVELS),
        ("n_points", C.c_int * MAX_LEVELS), ("final_error", C.c_float * MAX_LEVELS),
    ]


def build(force=False):
    """Compile the oracle with its Makefile (gcc; no GPU needed)."""
    src = os.path.join(_HERE, "uw_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, i16p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_float)
        L.uwo_default_params.argtypes = [C.POINTER(Params)]
        L.uwo_pyr_down.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.uwo_scharr.argtypes = [u8p, C.c_int, C.c_int, i16p, i16p]
        L.uwo_gradmag.argtypes = [i16p, i16p, C.c_longlong, u8p]
        L.uwo_candidates.argtypes = [u8p, C.c_int, C.c_int, C.c_double, f32p,
                                     C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.uwo_candidates.restype = C.c_int
        L.uwo_warp.argtypes = [f32p, C.c_int, f32p] + [C.c_float] * 6 + [f32p]
        L.uwo_se3_exp.argtypes = [f32p, f32p]
        L.uwo_se3_mul.argtypes = [f32p, f32p, f32p]
        L.uwo_se3_matrix.argtypes = [f32p, f32p]
        L.uwo_se3_scale_level.argtypes = [f32p, f32p]
        L.uwo_lu_solve6.argtypes = [f32p, f32p, f32p]
        L.uwo_lu_solve6.restype = C.c_int
        L.uwo_lu_invert6.argtypes = [f32p, f32p]
        L.uwo_lu_invert6.restype = C.c_int
        L.uwo_median_mat.argtypes = [f32p, C.c_int]
        L.uwo_median_mat.restype = C.c_float
        L.uwo_mad.argtypes = [f32p, C.c_int]
        L.uwo_mad.restype = C.c_float
        L.uwo_tukey_weights.argtypes = [f32p, C.c_int, f32p]
        L.uwo_huber_weights.argtypes = [f32p, C.c_int, C.c_float, f32p]
        L.uwo_estimate_pose.restype = C.c_int
        L.uwo_track_pair.restype = C.c_int
        L.uwo_track_pair.argtypes = [C.POINTER(Params), u8p, u8p, f32p, C.POINTER(Stats),
                                     C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def default_params(width=640, height=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5, **kw):
    p = Params()
    lib().uwo_default_params(C.byref(p))
    p.width, p.height, p.fx, p.fy, p.cx, p.cy = width, height, fx, fy, cx, cy
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def median_mat(v):
    v = np.ascontiguousarray(v, np.float32).ravel()
    return float(lib().uwo_median_mat(_p(v, C.c_float), v.size))


def mad(v):
    v = np.ascontiguousarray(v, np.float32).ravel()
    return float(lib().uwo_mad(_p(v, C.c_float), v.size))


def tukey_weights(v):
    v = np.ascontiguousarray(v, np.float32).ravel()
    w = np.empty(v.size, np.float32)
    lib().uwo_tukey_weights(_p(v, C.c_float), v.size, _p(w, C.c_float))
    return w


def huber_weights(v, delta):
    v = np.ascontiguousarray(v, np.float32).ravel()
    w = np.empty(v.size, np.float32)
    lib().uwo_huber_weights(_p(v, C.c_float), v.size, float(delta), _p(w, C.c_float))
    return w


def optimal_new_camera_matrix(K, dist, in_size, alpha, out_size):
    K = np.ascontiguousarray(K, np.float32)
    d = np.ascontiguousarray(dist, np.float32).ravel()
    o = np.empty((3, 3), np.float32)
    f = lib().uwo_optimal_new_camera_matrix
    f.argtypes = [C.POINTER(C.c_float)] * 2 + [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                                C.POINTER(C.c_float)]
    f.restype = None
    f(_p(K, C.c_float), _p(d, C.c_float), in_size[0], in_size[1], float(alpha), out_size[0],
      out_size[1], _p(o, C.c_float))
    return o


def init_undistort_rectify_map(K, dist, newK, out_size):
    K = np.ascontiguousarray(K, np.float32)
    d = np.ascontiguousarray(dist, np.float32).ravel()
    nK = np.ascontiguousarray(newK, np.float32)
    w, h = out_size
    m1 = np.empty((h, w, 2), np.int16)
    m2 = np.empty((h, w), np.uint16)
    f = lib().uwo_init_undistort_rectify_map
    f.argtypes = [C.POINTER(C.c_float)] * 3 + [C.c_int, C.c_int, C.POINTER(C.c_int16),
                                                C.POINTER(C.c_uint16)]
    f.restype = None
    f(_p(K, C.c_float), _p(d, C.c_float), _p(nK, C.c_float), w, h, _p(m1, C.c_int16),
      _p(m2, C.c_uint16))
    return m1, m2


def remap_bilinear(src, map1, map2):
    src = np.ascontiguousarray(src, np.uint8)
    m1 = np.ascontiguousarray(map1, np.int16)
    m2 = np.ascontiguousarray(map2, np.uint16)
    h, w = m2.shape
    out = np.empty((h, w), np.uint8)
    f = lib().uwo_remap_bilinear
    f.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_int16),
                  C.POINTER(C.c_uint16), C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    f.restype = None
    f(_p(src, C.c_uint8), src.shape[1], src.shape[0], _p(m1, C.c_int16), _p(m2, C.c_uint16), w, h,
      _p(out, C.c_uint8))
    return out


def calculate_roi(undistorted):
    u = np.ascontiguousarray(undistorted, np.uint8)
    roi = (C.c_int * 4)()
    f = lib().uwo_calculate_roi
    f.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_int)]
    f.restype = C.c_int
    if f(_p(u, C.c_uint8), u.shape[1], u.shape[0], roi) != 0:
        raise RuntimeError("CalculateROI scanned past the image border")
    return tuple(roi)


def pyr_down(img):
    h, w = img.shape
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty((h // 2, w // 2), np.uint8)
    lib().uwo_pyr_down(_p(img, C.c_uint8), w, h, _p(out, C.c_uint8))
    return out


def build_pyramid(img, levels=5):
    pyr = [np.ascontiguousarray(img, np.uint8)]
    for _ in range(1, levels):
        pyr.append(pyr_down(pyr[-1]))
    return pyr


def scharr(img):
    h, w = img.shape
    img = np.ascontiguousarray(img, np.uint8)
    gx = np.empty((h, w), np.int16)
    gy = np.empty((h, w), np.int16)
    lib().uwo_scharr(_p(img, C.c_uint8), w, h, _p(gx, C.c_int16), _p(gy, C.c_int16))
    return gx, gy


def sobel(img):
    h, w = img.shape
    img = np.ascontiguousarray(img, np.uint8)
    gx = np.empty((h, w), np.int16)
    gy = np.empty((h, w), np.int16)
    f = lib().uwo_sobel
    f.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_int16),
                  C.POINTER(C.c_int16)]
    f.restype = None
    f(_p(img, C.c_uint8), w, h, _p(gx, C.c_int16), _p(gy, C.c_int16))
    return gx, gy


def gradmag(gx, gy):
    gx = np.ascontiguousarray(gx, np.int16)
    gy = np.ascontiguousarray(gy, np.int16)
    g = np.empty(gx.shape, np.uint8)
    lib().uwo_gradmag(_p(gx, C.c_int16), _p(gy, C.c_int16), gx.size, _p(g, C.c_uint8))
    return g


def candidates(g, gradient_threshold=20.0):
    """Returns (N x 4 f32 points in the reference's x-major order, mean, ithr)."""
    h, w = g.shape
    g = np.ascontiguousarray(g, np.uint8)
    pts = np.empty((h * w, 4), np.float32)
    mean = C.c_double()
    ithr = C.c_int()
    n = lib().uwo_candidates(_p(g, C.c_uint8), w, h, gradient_threshold, _p(pts, C.c_float),
                             C.byref(mean), C.byref(ithr))
    return pts[:n].copy(), mean.value, ithr.value


def init_pyramid(w, h, fx, fy, cx, cy, levels=5):
    wl = (C.c_int * levels)()
    hl = (C.c_int * levels)()
    arrs = [(C.c_float * levels)() for _ in range(6)]
    L = lib()
    L.uwo_init_pyramid.argtypes = [C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int] + \
        [C.POINTER(C.c_int)] * 2 + [C.POINTER(C.c_float)] * 6
    L.uwo_init_pyramid(w, h, fx, fy, cx, cy, levels, wl, hl, *arrs)
    names = ["fx", "fy", "cx", "cy", "invfx", "invfy"]
    out = {"w": list(wl), "h": list(hl)}
    for nme, a in zip(names, arrs):
        out[nme] = np.array(list(a), np.float32)
    return out


def warp(pts4, pose7, fx, fy, cx, cy, invfx, invfy):
    pts4 = np.ascontiguousarray(pts4, np.float32)
    pose7 = np.ascontiguousarray(pose7, np.float32)
    out = np.empty_like(pts4)
    lib().uwo_warp(_p(pts4, C.c_float), pts4.shape[0], _p(pose7, C.c_float),
                   fx, fy, cx, cy, invfx, invfy, _p(out, C.c_float))
    return out


def se3_exp(t6):
    t6 = np.ascontiguousarray(t6, np.float32)
    o = np.empty(7, np.float32)
    lib().uwo_se3_exp(_p(t6, C.c_float), _p(o, C.c_float))
    return o


def se3_mul(a7, b7):
    a7 = np.ascontiguousarray(a7, np.float32)
    b7 = np.ascontiguousarray(b7, np.float32)
    o = np.empty(7, np.float32)
    lib().uwo_se3_mul(_p(a7, C.c_float), _p(b7, C.c_float), _p(o, C.c_float))
    return o


def chain_pose(prev7, rigid7, scale=40.0):
    a = np.ascontiguousarray(prev7, np.float32)
    b = np.ascontiguousarray(rigid7, np.float32)
    o = np.empty(7, np.float32)
    f = lib().uwo_chain_pose
    f.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]
    f.restype = None
    f(_p(a, C.c_float), _p(b, C.c_float), float(scale), _p(o, C.c_float))
    return o


def se3_matrix(p7):
    p7 = np.ascontiguousarray(p7, np.float32)
    o = np.empty(16, np.float32)
    lib().uwo_se3_matrix(_p(p7, C.c_float), _p(o, C.c_float))
    return o.reshape(4, 4)


def se3_scale_level(p7):
    p7 = np.ascontiguousarray(p7, np.float32)
    o = np.empty(7, np.float32)
    lib().uwo_se3_scale_level(_p(p7, C.c_float), _p(o, C.c_float))
    return o


def lu_solve6(A, b):
    A = np.ascontiguousarray(A, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    x = np.empty(6, np.float32)
    ok = lib().uwo_lu_solve6(_p(A, C.c_float), _p(b, C.c_float), _p(x, C.c_float))
    return x, ok


def cholesky_lm_solve6(A, b, lam):
    A = np.ascontiguousarray(A, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    x = np.empty(6, np.float32)
    f = lib().uwo_cholesky_lm_solve6
    f.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]
    f.restype = C.c_int
    ok = f(_p(A, C.c_float), _p(b, C.c_float), float(lam), _p(x, C.c_float))
    return x, ok


def lu_invert6(A):
    A = np.ascontiguousarray(A, np.float32)
    Ai = np.empty((6, 6), np.float32)
    ok = lib().uwo_lu_invert6(_p(A, C.c_float), _p(Ai, C.c_float))
    return Ai, ok


def depth_pyr_down(d):
    h, w = d.shape
    d = np.ascontiguousarray(d, np.uint16)
    out = np.empty((h // 2, w // 2), np.uint16)
    f = lib().uwo_depth_pyr_down
    f.argtypes = [C.POINTER(C.c_uint16), C.c_int, C.c_int, C.POINTER(C.c_uint16)]
    f.restype = None
    f(_p(d, C.c_uint16), w, h, _p(out, C.c_uint16))
    return out


def candidates_depth(g, depth, gradient_threshold=20.0, depth_mode=DEPTH_REFERENCE):
    """Depth branch of ObtainCandidatePoints: (N x 4 points [x, y, Z, 1], integer depths used)."""
    h, w = g.shape
    g = np.ascontiguousarray(g, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    pts = np.empty((h * w, 4), np.float32)
    z = np.empty(h * w, np.uint16)
    f = lib().uwo_candidates_depth
    f.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.c_int, C.c_int, C.c_double,
                  C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint16)]
    f.restype = C.c_int
    n = f(_p(g, C.c_uint8), _p(depth, C.c_uint16), w, h, gradient_threshold, depth_mode,
          _p(pts, C.c_float), _p(z, C.c_uint16))
    return pts[:n].copy(), z[:n].copy()


def all_points_depth(depth, lvl):
    """Tracker::ObtainAllPoints, depth branch (Tracker.cpp:1259-1300): one row per pixel in
    ROW-major order (:1268-1269); [x, y, at<short>(y,x) * (0.0002 / 2^lvl), 1] where the depth
    read as a signed short is > 0 (:1273-1279), [0, 0, 1, 0] elsewhere (:1290-1292)."""
    d = np.ascontiguousarray(depth, np.uint16).view(np.int16)
    h, w = d.shape
    factor_lvl = np.float32(np.float64(np.float32(0.0002)) / np.float64(2.0 ** lvl))  # :1266
    pts = np.zeros((h * w, 4), np.float32)
    pts[:, 2] = 1.0
    ys, xs = np.divmod(np.arange(h * w), w)
    v = d.ravel() > 0
    pts[v, 0] = xs[v]
    pts[v, 1] = ys[v]
    pts[v, 2] = d.ravel()[v].astype(np.float32) * factor_lvl
    pts[v, 3] = 1.0
    return pts


class FrameData:
    """What uw::Frame holds after pyramid + ApplyGradient + ObtainCandidatePoints (or, with
    depth_mode = DEPTH_ALL_POINTS, ObtainAllPoints)."""

    def __init__(self, img, levels=5, gradient_threshold=20.0, with_candidates=True, depth=None,
                 depth_mode=DEPTH_NONE, gradient_op=0):
        self.images = build_pyramid(img, levels)
        self.gx, self.gy, self.g, self.cand, self.mean, self.ithr = [], [], [], [], [], []
        self.depths, self.zsrc = [], []
        if depth is not None and depth_mode != DEPTH_NONE:
            self.depths = [np.ascontiguousarray(depth, np.uint16)]
            for _ in range(1, levels):
                self.depths.append(depth_pyr_down(self.depths[-1]))
        if with_candidates:
            for lvl, im in enumerate(self.images):
                gx, gy = sobel(im) if gradient_op == 1 else scharr(im)
                g = gradmag(gx, gy)
                c, m, t = candidates(g, gradient_threshold)
                if self.depths and depth_mode == DEPTH_ALL_POINTS:
                    c = all_points_depth(self.depths[lvl], lvl)
                elif self.depths:
                    c, z = candidates_depth(g, self.depths[lvl], gradient_threshold, depth_mode)
                    self.zsrc.append(z)
                self.gx.append(gx)
                self.gy.append(gy)
                self.g.append(g)
                self.cand.append(c)
                self.mean.append(m)
                self.ithr.append(t)


def estimate_pose(params, prev, cur, init_pose=None, trace_cap=256):
    """Tracker::EstimatePose on two FrameData.  Returns (pose7, Stats, [IterTrace])."""
    L = params.levels
    u8pp = (C.POINTER(C.c_uint8) * MAX_LEVELS)
    i16pp = (C.POINTER(C.c_int16) * MAX_LEVELS)
    f32pp = (C.POINTER(C.c_float) * MAX_LEVELS)
    pi, ci, gx, gy, cd = u8pp(), u8pp(), i16pp(), i16pp(), f32pp()
    nc = (C.c_int * MAX_LEVELS)()
    keep = []
    for l in range(L):
        a = np.ascontiguousarray(prev.images[l]); keep.append(a); pi[l] = _p(a, C.c_uint8)
        a = np.ascontiguousarray(cur.images[l]); keep.append(a); ci[l] = _p(a, C.c_uint8)
        a = np.ascontiguousarray(prev.gx[l]); keep.append(a); gx[l] = _p(a, C.c_int16)
        a = np.ascontiguousarray(prev.gy[l]); keep.append(a); gy[l] = _p(a, C.c_int16)
        a = np.ascontiguousarray(prev.cand[l], np.float32); keep.append(a)
        cd[l] = _p(a, C.c_float)
        nc[l] = a.shape[0]
    out = np.empty(7, np.float32)
    st = Stats()
    tr = (IterTrace * trace_cap)()
    ntr = C.c_int(0)
    ip = None
    if init_pose is not None:
        ipa = np.ascontiguousarray(init_pose, np.float32)
        keep.append(ipa)
        ip = _p(ipa, C.c_float)
    rc = lib().uwo_estimate_pose(C.byref(params), pi, ci, gx, gy, cd, nc, ip,
                                 _p(out, C.c_float), C.byref(st), tr, trace_cap, C.byref(ntr))
    if rc != 0:
        raise RuntimeError("uwo_estimate_pose failed: %d" % rc)
    return out, st, [tr[i] for i in range(ntr.value)]


def track_pair(params, prev0, cur0):
    """Full track (pyramids, gradient, candidates, estimate).  Returns (pose7, Stats, secs[4])."""
    prev0 = np.ascontiguousarray(prev0, np.uint8)
    cur0 = np.ascontiguousarray(cur0, np.uint8)
    out = np.empty(7, np.float32)
    st = Stats()
    secs = (C.c_double * 4)()
    rc = lib().uwo_track_pair(C.byref(params), _p(prev0, C.c_uint8), _p(cur0, C.c_uint8),
                              _p(out, C.c_float), C.byref(st), secs)
    if rc != 0:
        raise RuntimeError("uwo_track_pair failed: %d" % rc)
    return out, st, list(secs)


def track_sequence(params, frames):
    """frames: u8 [n, H, W].  Returns (poses [n-1, 7], seconds [n-1], [Stats])."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n = frames.shape[0]
    poses = np.empty((n - 1, 7), np.float32)
    secs = np.empty(n - 1, np.float64)
    stats = (Stats * (n - 1))()
    L = lib()
    L.uwo_track_sequence.restype = C.c_int
    L.uwo_track_sequence.argtypes = [C.POINTER(Params), C.POINTER(C.c_uint8), C.c_int,
                                     C.POINTER(C.c_float), C.POINTER(C.c_double),
                                     C.POINTER(Stats)]
    rc = L.uwo_track_sequence(C.byref(params), _p(frames, C.c_uint8), n, _p(poses, C.c_float),
                              _p(secs, C.c_double), stats)
    if rc != 0:
        raise RuntimeError("uwo_track_sequence failed: %d" % rc)
    return poses, secs, list(stats)


class Stream:
    """uwo_stream: the reference's per-frame loop with the previous frame kept between calls
    (create = frame 0 prepared; track = one pose-track)."""

    def __init__(self, params, frame0):
        L = lib()
        L.uwo_stream_create.restype = C.c_void_p
        L.uwo_stream_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_uint8)]
        L.uwo_stream_destroy.argtypes = [C.c_void_p]
        L.uwo_stream_track.restype = C.c_int
        L.uwo_stream_track.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_float),
                                       C.POINTER(Stats)]
        self._L = L
        self._params = params
        f0 = np.ascontiguousarray(frame0, np.uint8)
        self._h = L.uwo_stream_create(C.byref(params), _p(f0, C.c_uint8))
        if not self._h:
            raise MemoryError("uwo_stream_create")

    def track(self, frame, with_stats=False):
        f = np.ascontiguousarray(frame, np.uint8)
        out = np.empty(7, np.float32)
        st = Stats() if with_stats else None
        rc = self._L.uwo_stream_track(self._h, _p(f, C.c_uint8), _p(out, C.c_float),
                                      C.byref(st) if with_stats else None)
        if rc != 0:
            raise RuntimeError("uwo_stream_track failed: %d" % rc)
        return (out, st) if with_stats else out

    def close(self):
        if self._h:
            self._L.uwo_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sweep_range(params, frame_prev, frame_cur, lvl, lo, hi, pose7):
    """uwo_sweep_range on FrameData: the 32 sums of candidate rows [lo, hi) of level lvl."""
    L = lib()
    u8p, i16p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_float)
    L.uwo_sweep_range.restype = C.c_int
    L.uwo_sweep_range.argtypes = [C.POINTER(Params), C.c_int, u8p, u8p, i16p, i16p, f32p, C.c_int,
                                  C.c_int, f32p, C.POINTER(C.c_double)]
    I1 = np.ascontiguousarray(frame_prev.images[lvl])
    I2 = np.ascontiguousarray(frame_cur.images[lvl])
    gx = np.ascontiguousarray(frame_prev.gx[lvl])
    gy = np.ascontiguousarray(frame_prev.gy[lvl])
    cd = np.ascontiguousarray(frame_prev.cand[lvl], np.float32)
    pose = np.ascontiguousarray(pose7, np.float32)
    sums = np.zeros(32, np.float64)
    rc = L.uwo_sweep_range(C.byref(params), lvl, _p(I1, C.c_uint8), _p(I2, C.c_uint8),
                           _p(gx, C.c_int16), _p(gy, C.c_int16), _p(cd, C.c_float), lo, hi,
                           _p(pose, C.c_float), _p(sums, C.c_double))
    if rc != 0:
        raise RuntimeError("uwo_sweep_range failed: %d" % rc)
    return sums


def gn_update(params, sums32, k, pose7, last_error):
    """uwo_gn_update: returns (level_finished, new_pose7, new_last_error)."""
    L = lib()
    L.uwo_gn_update.restype = C.c_int
    L.uwo_gn_update.argtypes = [C.POINTER(Params), C.POINTER(C.c_double), C.c_int,
                                C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
    sums = np.ascontiguousarray(sums32, np.float64)
    pose = np.array(pose7, np.float32)
    le = C.c_float(last_error)
    brk = L.uwo_gn_update(C.byref(params), _p(sums, C.c_double), k, _p(pose, C.c_float),
                          C.byref(le), None)
    return bool(brk), pose, le.value
