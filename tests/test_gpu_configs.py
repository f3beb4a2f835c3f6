"""GPU parity over the configuration space of the C ABI (levels, optimised range, thresholds,
iteration limits, frame sizes up to the maximum): everything against the oracle with the same
parameters, bit for bit."""
import numpy as np
import pytest

from uw_slam_b200 import synth

pytestmark = pytest.mark.gpu


def run_case(oracle, w, h, seed, levels, first, last, batch=1, extra_flags=0, **cfg):
    import uw_slam_b200 as U
    import uw_slam_b200._lib as L
    fx, fy, cx, cy = 0.8 * w, 0.82 * w, w / 2 - 0.5, h / 2 - 0.5
    synth.CALIB["_cfg"] = (w, h, fx, fy, cx, cy)
    pairs = [synth.render_pair("_cfg", seed + i)[:2] for i in range(batch)]
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2 * batch, levels=levels, first_level=first, last_level=last,
                        flags=L.FLAG_TRACE | extra_flags, **cfg)
    fp = t.AddFrames(list(range(batch)), np.stack([p[0] for p in pairs]))
    fc = t.AddFrames(list(range(batch, 2 * batch)), np.stack([p[1] for p in pairs]))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    poses, stats = t.EstimatePose(fp, fc, return_stats=True)
    okw = {k: v for k, v in cfg.items() if k in ("max_iterations", "epsilon", "residual_scale",
                                                 "gradient_threshold", "solve_mode",
                                                 "weight_mode", "huber_delta", "lm_lambda", "sampling")}
    gop = cfg.get("gradient_op", 0)
    p = oracle.default_params(w, h, fx, fy, cx, cy, levels=levels, first_level=first,
                              last_level=last, **okw)
    thr = cfg.get("gradient_threshold", 20.0)
    for i in range(batch):
        rp = oracle.FrameData(pairs[i][0], levels=levels, gradient_threshold=thr, gradient_op=gop)
        rc = oracle.FrameData(pairs[i][1], levels=levels, with_candidates=False)
        if i == 0:
            for lvl in range(levels):
                assert np.array_equal(fp[0].image(lvl), rp.images[lvl]), lvl
                gx, gy, g = fp[0].gradients(lvl)
                assert np.array_equal(gx, rp.gx[lvl]) and np.array_equal(gy, rp.gy[lvl]), lvl
                assert np.array_equal(g, rp.g[lvl]), lvl
                assert np.array_equal(fp[0].candidatePoints(lvl), rp.cand[lvl]), lvl
        opose, ostats, otrace = oracle.estimate_pose(p, rp, rc, trace_cap=1024)
        if batch <= 64:
            tr = t.get_trace(i, cap=1024)
            assert [(a.level, a.k, a.n_valid, a.broke, a.sum_r2) for a in tr] == \
                [(b.level, b.k, b.n_valid, b.broke, b.sum_r2) for b in otrace], i
            for a, b in zip(tr, otrace):
                assert np.array_equal(np.array(a.A[:]), np.array(b.A[:])), (i, b.level, b.k)
                assert np.array_equal(np.array(a.delta[:]), np.array(b.delta[:])), (i, b.level, b.k)
        assert np.array_equal(poses[i], opose), i
        assert list(stats[i].iterations)[:levels] == list(ostats.iterations)[:levels]
    t.close()


@pytest.mark.parametrize("levels,first,last", [(1, 0, 0), (2, 1, 0), (3, 2, 0), (3, 2, 2),
                                               (5, 4, 0), (5, 3, 2), (6, 5, 1), (7, 6, 3)])
def test_level_configurations(oracle, levels, first, last):
    # 320x256 is divisible by 2^6; level 0 can be optimised too (records on level 0)
    run_case(oracle, 320, 256, 11, levels, first, last)


@pytest.mark.parametrize("cfg", [dict(max_iterations=1), dict(max_iterations=3),
                                 dict(epsilon=5.0), dict(residual_scale=12.5),
                                 dict(residual_scale=1.0, max_iterations=8),
                                 dict(gradient_threshold=5.0), dict(gradient_threshold=60.5),
                                 dict(gradient_threshold=400.0),   # nothing passes: U2 everywhere
                                 dict(solve_mode=1, weight_mode=1),
                                 dict(gradient_op=1), dict(gradient_op=1, weight_mode=1),
                                 dict(sampling=1), dict(sampling=1, residual_scale=12.5),
                                 # the north-star wording end to end: Sobel, bilinear, LM/Cholesky
                                 dict(sampling=1, gradient_op=1, solve_mode=2, lm_lambda=0.2),
                                 dict(solve_mode=2), dict(solve_mode=2, lm_lambda=0.0),
                                 dict(solve_mode=2, lm_lambda=7.5, weight_mode=2),
                                 dict(weight_mode=2, huber_delta=3.0, residual_scale=20.0)])
def test_parameter_variations(oracle, cfg):
    run_case(oracle, 160, 128, 5, 5, 4, 1, **cfg)


@pytest.mark.parametrize("cfg", [dict(), dict(levels=4, first=3, last=0),
                                 dict(extra=dict(solve_mode=2, lm_lambda=0.2)),
                                 dict(extra=dict(sampling=1)),
                                 dict(extra=dict(sampling=1, gradient_op=1, solve_mode=2)),
                                 dict(extra=dict(residual_scale=12.5)),   # generic point loop
                                 dict(levels=7, first=6, last=3),         # levels beyond 4
                                 dict(extra=dict(sampling=1), flags=4)])  # 4: cluster kernel
def test_dataflow_kernel_configurations(oracle, cfg):
    # 32 problems -> the dataflow kernel (bilinear sampling included); non-default level ranges,
    # the LM/Cholesky solver, a non-integer residual scale (generic loop), and the same batch
    # on the cluster kernel for comparison
    w, h = (320, 256) if cfg.get("levels", 5) > 5 else (160, 128)
    run_case(oracle, w, h, 40, cfg.get("levels", 5), cfg.get("first", 4), cfg.get("last", 1),
             batch=32, extra_flags=cfg.get("flags", 0), **cfg.get("extra", {}))


@pytest.mark.parametrize("w,h", [(4096, 4096), (4096, 16), (16, 4096), (1936, 1216)])
def test_extreme_frame_sizes(oracle, w, h):
    """The largest supported frame (12-bit record coordinates), degenerate strips and a size
    that leaves ragged tiles in every kernel."""
    import uw_slam_b200 as U
    rng = np.random.default_rng(w + h)
    levels = 5 if min(w, h) >= 32 else 1
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    # smooth it a little so that candidates are a realistic fraction
    img = ((img.astype(np.uint16) + np.roll(img, 1, 0) + np.roll(img, 1, 1) +
            np.roll(img, 2, 1)) // 4).astype(np.uint8)
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, w, w, w / 2, h / 2).GetK(),
                        max_frames=1, levels=levels, first_level=levels - 1,
                        last_level=min(1, levels - 1))
    f = t.AddFrames([0], img)[0]
    t.ApplyGradient(f)
    t.ObtainCandidatePoints(f)
    ref = oracle.FrameData(img, levels=levels)
    for lvl in range(levels):
        assert np.array_equal(f.image(lvl), ref.images[lvl]), lvl
        gx, gy, g = f.gradients(lvl)
        assert np.array_equal(gx, ref.gx[lvl]) and np.array_equal(gy, ref.gy[lvl]), lvl
        assert np.array_equal(g, ref.g[lvl]), lvl
        assert np.array_equal(f.candidatePoints(lvl), ref.cand[lvl]), lvl
    t.close()


def test_rejected_configurations():
    import uw_slam_b200 as U
    K = U.CameraModel.from_intrinsics(640, 480, 500, 500, 320, 240).GetK()
    for bad in (dict(width=4112, height=480), dict(width=648, height=480),   # > 4096, w % 16
                dict(width=640, height=488),                                   # h % 16
                dict(levels=0), dict(levels=8), dict(first_level=5), dict(last_level=-1),
                dict(first_level=1, last_level=2), dict(max_iterations=0), dict(max_frames=0),
                dict(solve_mode=3), dict(solve_mode=2, lm_lambda=-1.0), dict(cluster_size=3),
                dict(gradient_op=2), dict(sampling=2), dict(sampling=1, weight_mode=1),
                dict(gradient_threshold=-1.0),
                dict(device=99)):
        cfg = dict(bad)
        w, h = cfg.pop("width", 640), cfg.pop("height", 480)
        with pytest.raises(U.UwtError):
            U.Tracker(False).InitializePyramid(w, h, K, **cfg)


def test_randomised_small_configurations(oracle):
    """Seeded random sweep over frame sizes (every multiple of 16 the level count allows), level
    ranges, thresholds and weight / solve modes on noise-plus-structure images: image kernels,
    candidate lists and the whole estimate against the oracle."""
    import uw_slam_b200 as U
    rng = np.random.default_rng(2024)
    for case in range(48):
        levels = int(rng.integers(1, 6))
        mult = max(16, 1 << (levels - 1))
        lo = max(2 << (levels - 1), mult)           # coarsest level at least 2 x 2
        w = int(rng.integers(lo // mult, 384 // mult + 1)) * mult
        hm = max(2, 1 << (levels - 1))              # the reference requires even sizes
        h = int(rng.integers(max(1, (2 << (levels - 1)) // hm), 320 // hm + 1)) * hm
        first = int(rng.integers(0, levels))
        last = int(rng.integers(0, first + 1))
        cfg = dict(gradient_threshold=float(rng.choice([2.0, 11.5, 20.0, 35.0])),
                   max_iterations=int(rng.choice([1, 4, 50])),
                   solve_mode=int(rng.integers(0, 3)), weight_mode=int(rng.integers(0, 3)),
                   huber_delta=float(rng.choice([2.5, 9.0])))
        fx, fy = float(rng.uniform(0.6, 1.4) * w), float(rng.uniform(0.6, 1.4) * w)
        cx, cy = w / 2 + float(rng.normal(0, 3)), h / 2 + float(rng.normal(0, 3))
        ys, xs = np.mgrid[0:h, 0:w]
        base = 120 + 70 * np.sin(xs * rng.uniform(0.05, 0.4)) * np.cos(ys * rng.uniform(0.05, 0.4))
        prev = np.clip(base + rng.integers(-25, 25, (h, w)), 0, 255).astype(np.uint8)
        cur = np.clip(np.roll(base, (int(rng.integers(-1, 2)), int(rng.integers(-1, 2))), (0, 1)) +
                      rng.integers(-25, 25, (h, w)), 0, 255).astype(np.uint8)
        t = U.Tracker(False)
        t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                            max_frames=2, levels=levels, first_level=first, last_level=last, **cfg)
        fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
        t.ApplyGradient(fp)
        t.ObtainCandidatePoints(fp)
        pose = t.EstimatePose(fp, fc)[0]
        rp = oracle.FrameData(prev, levels=levels, gradient_threshold=cfg["gradient_threshold"])
        rc = oracle.FrameData(cur, levels=levels, with_candidates=False)
        tag = (case, w, h, levels, first, last, cfg)
        for lvl in range(levels):
            assert np.array_equal(fp.image(lvl), rp.images[lvl]), tag
            gx, gy, g = fp.gradients(lvl)
            assert np.array_equal(gx, rp.gx[lvl]) and np.array_equal(gy, rp.gy[lvl]), tag
            assert np.array_equal(g, rp.g[lvl]), tag
            assert np.array_equal(fp.candidatePoints(lvl), rp.cand[lvl]), tag
        okw = {k: v for k, v in cfg.items() if k != "gradient_threshold"}
        p = oracle.default_params(w, h, fx, fy, cx, cy, levels=levels, first_level=first,
                                  last_level=last, **okw)
        assert np.array_equal(pose, oracle.estimate_pose(p, rp, rc)[0]), tag
        t.close()


@pytest.mark.parametrize("batch", [1, 32])
def test_principal_point_at_zero_uses_exact_division(oracle, batch):
    """cx = 0.5 puts the level-1 principal point at exactly 0 (Tracker.cpp:321: (cx + 0.5) / 2 -
    0.5): the shared-reciprocal shortcut of the fast sweep is not provably exact there
    (docs/ARITHMETIC.md, implementation notes), so the handle takes the generic IEEE division for
    every point (Geom::exact_div).  Same parity bar: every sweep against the oracle."""
    import uw_slam_b200 as U
    import uw_slam_b200._lib as L
    w, h = 160, 128
    fx, fy, cx, cy = 0.8 * w, 0.82 * w, 0.5, h / 2 - 0.5
    synth.CALIB["_cfg0"] = (w, h, fx, fy, cx, cy)
    pairs = [synth.render_pair("_cfg0", 70 + i)[:2] for i in range(batch)]
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2 * batch, flags=L.FLAG_TRACE)
    assert t.level_info(1).cx == 0.0
    fp = t.AddFrames(list(range(batch)), np.stack([p[0] for p in pairs]))
    fc = t.AddFrames(list(range(batch, 2 * batch)), np.stack([p[1] for p in pairs]))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    poses = t.EstimatePose(fp, fc)
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    for i in range(batch):
        rp = oracle.FrameData(pairs[i][0])
        rc = oracle.FrameData(pairs[i][1], with_candidates=False)
        opose, _, otr = oracle.estimate_pose(p, rp, rc)
        tr = t.get_trace(i)
        assert [(a.level, a.k, a.n_valid, a.sum_r2) for a in tr] == \
            [(b.level, b.k, b.n_valid, b.sum_r2) for b in otr], i
        for a, b in zip(tr, otr):
            assert np.array_equal(np.array(a.A[:]), np.array(b.A[:])), (i, b.level, b.k)
        assert np.array_equal(poses[i], opose), i
    t.close()


@pytest.mark.parametrize("batch", [1, 32])
def test_points_outside_the_division_window(oracle, batch):
    """Initial poses that put Z' at 0 or at ~2^-64 (outside the [2^-60, 2^60) window of the
    shared-reciprocal division): the fast sweep defers those points to the generic IEEE path.
    Every quantity must still equal the oracle's (here: all points invalid, rule U2)."""
    import uw_slam_b200 as U
    import uw_slam_b200._lib as L
    calib = "small"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    prev, cur, _, _ = synth.render_pair(calib, 9)
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=max(2, batch), flags=L.FLAG_TRACE)   # n <= max_frames
    t.AddFrames([0, 1], np.stack([prev, cur]))
    t.ApplyGradient([0])
    t.ObtainCandidatePoints([0])
    a = np.float32(1e-19)
    inits = [np.array([0, 0, 0, 1, 0, 0, -1], np.float32),                    # Z' == 0
             np.array([a / 2, 0, 0, 1, 0, 0, -1], np.float32),                # |Z'| ~ 2^-64
             np.array([0, 0, 0, 1, 0.01, 0, 0], np.float32)]                  # a normal one
    init = np.stack([inits[i % 3] for i in range(batch)])
    poses, stats = t.EstimatePose([0] * batch, [1] * batch, init_poses=init, return_stats=True)
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    rp, rc = oracle.FrameData(prev), oracle.FrameData(cur, with_candidates=False)
    for i in range(batch):
        opose, ost, otr = oracle.estimate_pose(p, rp, rc, init_pose=init[i])
        assert np.array_equal(poses[i], opose, equal_nan=True), (i, poses[i], opose)
        assert list(stats[i].evaluations)[:5] == list(ost.evaluations)[:5], i
        tr = t.get_trace(i)
        assert [(x.level, x.k, x.n_valid, x.sum_r2) for x in tr] == \
            [(y.level, y.k, y.n_valid, y.sum_r2) for y in otr], i
    t.close()


@pytest.mark.parametrize("w,h,levels,first,last,thr", [
    (1008, 496, 5, 4, 1, 20.0),    # level widths 504 / 252 / 126 / 63: 3-, 2- and 3-pixel tail groups
    (1008, 496, 5, 4, 0, 0.0),     # level 0 has records too; threshold = mean: about half the pixels
    (272, 144, 4, 3, 1, 20.0),     # 136 / 68 / 34 columns: one strip and an 8-column second strip
    (2064, 80, 3, 2, 1, 5.0),      # 1032 columns = 8 strips + 8 columns, 40 rows < one segment
    (48, 1040, 3, 2, 0, 20.0),     # tall and narrow: 17 row segments, one partial strip
])
def test_candidate_scatter_tile_edges(oracle, w, h, levels, first, last, thr):
    """The scatter kernel's tiles (128 columns x 64 rows staged by tensor copies, dense stencil
    phase on 4-pixel groups, column walk on 8-row masks) against ObtainCandidatePoints on frames
    whose level sizes are not multiples of the tile, of 4 pixels or of 8 rows: white noise, so
    every border, saturation and partial-group path selects something."""
    import uw_slam_b200 as U
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    img[:, : w // 3] //= 4      # a flatter region, so that selection is not uniform
    fx, fy, cx, cy = 0.8 * w, 0.82 * w, w / 2 - 0.5, h / 2 - 0.5
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=3, levels=levels, first_level=first, last_level=last,
                        gradient_threshold=thr)
    # slot 2 (not 0): the per-slot offsets of the tensor maps are part of what is tested
    f = t.AddFrames([2], img)[0]
    t.ApplyGradient(f)
    t.ObtainCandidatePoints(f)
    ref = oracle.FrameData(img, levels=levels, gradient_threshold=thr)
    for lvl in range(levels):
        c = f.candidatePoints(lvl)
        assert c.shape == ref.cand[lvl].shape, (lvl, c.shape, ref.cand[lvl].shape)
        assert np.array_equal(c, ref.cand[lvl]), lvl
        if last <= lvl <= first:
            r = t.get_records(f.slot, lvl)
            xs, ys = ref.cand[lvl][:, 0].astype(int), ref.cand[lvl][:, 1].astype(int)
            assert np.array_equal(r["x"], xs) and np.array_equal(r["y"], ys), lvl
            assert np.array_equal(r["i1"], ref.images[lvl][ys, xs]), lvl
            assert np.array_equal(r["gx"], ref.gx[lvl][ys, xs]), lvl
            assert np.array_equal(r["gy"], ref.gy[lvl][ys, xs]), lvl
    t.close()


def test_large_frames_stay_on_the_dataflow_kernel_beyond_296_problems(oracle):
    """320 problems of 1024x640 (finest optimised level 512x320 = 160 k pixels) in ONE call: the
    batch is served by the dataflow kernel (two launches: ring init + persistent kernel) with
    16384-record tasks, not by one-CTA clusters; 8 distinct pairs replicated 40 times must give
    40 identical copies of the oracle's 8 poses."""
    import uw_slam_b200 as U
    w, h = 1024, 640
    fx, fy, cx, cy = 0.8 * w, 0.82 * w, w / 2 - 0.5, h / 2 - 0.5
    synth.CALIB["_big"] = (w, h, fx, fy, cx, cy)
    distinct, copies = 8, 40
    pairs = [synth.render_pair("_big", 300 + i)[:2] for i in range(distinct)]
    n = distinct * copies
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2 * n)
    prev = np.stack([pairs[i % distinct][0] for i in range(n)])
    cur = np.stack([pairs[i % distinct][1] for i in range(n)])
    fp = t.AddFrames(list(range(n)), prev)
    fc = t.AddFrames(list(range(n, 2 * n)), cur)
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    l0 = t.launch_count()
    poses, stats = t.EstimatePose(fp, fc, return_stats=True)
    assert t.launch_count() - l0 == 2        # flow_init_kernel + estimate_flow_kernel
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    for i in range(distinct):
        opose, ostats, _ = oracle.estimate_pose(p, oracle.FrameData(pairs[i][0]),
                                                oracle.FrameData(pairs[i][1], with_candidates=False))
        for c in range(copies):
            assert np.array_equal(poses[i + c * distinct], opose), (i, c)
            assert list(stats[i + c * distinct].evaluations)[:5] == list(ostats.evaluations)[:5]
    t.close()
