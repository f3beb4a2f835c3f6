"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Integer / index work must be bit-exact; the normal equations, the GN
update and the poses are compared bit-for-bit too (the arithmetic spec makes them so), with
the north-star tolerance (1e-5 rad, 1e-5 of the baseline length) as the stated bar.
"""
import numpy as np
import pytest

from uw_slam_b200 import synth

pytestmark = pytest.mark.gpu

CALIBS = ["tiny", "small", "tum", "euroc"]


def make_tracker(calib_name, **cfg):
    import uw_slam_b200 as U
    w, h, fx, fy, cx, cy = synth.CALIB[calib_name]
    cam = U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy)
    t = U.Tracker(False)
    cfg.setdefault("max_frames", 2)
    t.InitializePyramid(w, h, cam.GetK(), **cfg)
    return t


@pytest.fixture(scope="module")
def pairs():
    cache = {}

    def get(calib, seed):
        if (calib, seed) not in cache:
            cache[(calib, seed)] = synth.render_pair(calib, seed)[:2]
        return cache[(calib, seed)]
    return get


def pose_close(a, b, tol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    # rotation: angle of q_a^-1 q_b ; translation: relative to its length
    dot = abs(float(np.dot(a[:4], b[:4]))) / (np.linalg.norm(a[:4]) * np.linalg.norm(b[:4]))
    ang = 2.0 * np.arccos(min(1.0, dot))
    tl = max(np.linalg.norm(b[4:]), 1e-12)
    return ang <= tol and np.linalg.norm(a[4:] - b[4:]) <= tol * max(tl, 1.0) and \
        np.linalg.norm(a[4:] - b[4:]) / tl <= tol * max(1.0, 1.0 / tl)


@pytest.mark.parametrize("calib", CALIBS + ["tum_mono"])
def test_pyramid_gradients_candidates_bit_exact(oracle, pairs, calib):
    prev, _ = pairs(calib, 3)
    t = make_tracker(calib)
    f = t.AddFrames([0], prev)[0]
    t.ApplyGradient(f)
    t.ObtainCandidatePoints(f)
    ref = oracle.FrameData(prev)
    for lvl in range(5):
        assert np.array_equal(f.image(lvl), ref.images[lvl]), ("image", lvl)
        gx, gy, g = f.gradients(lvl)
        assert np.array_equal(gx, ref.gx[lvl]), ("gx", lvl)
        assert np.array_equal(gy, ref.gy[lvl]), ("gy", lvl)
        assert np.array_equal(g, ref.g[lvl]), ("g", lvl)
        c = f.candidatePoints(lvl)
        assert c.shape == ref.cand[lvl].shape, ("ncand", lvl, c.shape, ref.cand[lvl].shape)
        assert np.array_equal(c, ref.cand[lvl]), ("cand", lvl)
        if 1 <= lvl <= 4:
            # the packed records EstimatePose streams: same order, and I1 / gradientX_ /
            # gradientY_ at every candidate equal the reference planes
            r = t.get_records(f.slot, lvl)
            xs, ys = ref.cand[lvl][:, 0].astype(int), ref.cand[lvl][:, 1].astype(int)
            assert np.array_equal(r["x"], xs) and np.array_equal(r["y"], ys)
            assert np.array_equal(r["i1"], ref.images[lvl][ys, xs])
            assert np.array_equal(r["gx"], ref.gx[lvl][ys, xs])
            assert np.array_equal(r["gy"], ref.gy[lvl][ys, xs])
    t.close()


def test_random_noise_image_bit_exact(oracle):
    # worst case for the integer kernels: white noise (every border / saturation path)
    rng = np.random.default_rng(0)
    w, h = synth.CALIB["euroc"][:2]
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    t = make_tracker("euroc")
    f = t.AddFrames([1], img)[0]
    t.ApplyGradient(f)
    t.ObtainCandidatePoints(f)
    ref = oracle.FrameData(img)
    for lvl in range(5):
        assert np.array_equal(f.image(lvl), ref.images[lvl])
        gx, gy, g = f.gradients(lvl)
        assert np.array_equal(gx, ref.gx[lvl]) and np.array_equal(gy, ref.gy[lvl])
        assert np.array_equal(g, ref.g[lvl])
        assert np.array_equal(f.candidatePoints(lvl), ref.cand[lvl])
    t.close()


def test_flat_image_has_no_candidates(oracle):
    w, h = synth.CALIB["small"][:2]
    img = np.full((h, w), 77, np.uint8)
    t = make_tracker("small")
    f = t.AddFrames([0], img)[0]
    g2 = t.AddFrames([1], img)[0]
    t.ApplyGradient(f)
    t.ObtainCandidatePoints(f)
    for lvl in range(5):
        assert f.candidatePoints(lvl).shape[0] == 0
    pose, stats = t.EstimatePose(f, g2, return_stats=True)
    # no points on any level (ARITHMETIC.md U2): identity, renormalised per level
    assert np.allclose(pose[0], [0, 0, 0, 1, 0, 0, 0])
    assert list(stats[0].iterations)[:5] == [0] * 5
    t.close()


@pytest.mark.parametrize("calib", ["small", "tum"])
def test_warp_function_matches_oracle(oracle, calib):
    t = make_tracker(calib)
    rng = np.random.default_rng(4)
    pose = oracle.se3_exp(np.array([4e-3, -3e-3, 2e-3, 2e-3, -3e-3, 4e-3], np.float32))
    for lvl in [1, 3]:
        i = t.level_info(lvl)
        n = 5000
        pts = np.ones((n, 4), np.float32)
        pts[:, 0] = rng.integers(0, i.width, n)
        pts[:, 1] = rng.integers(0, i.height, n)
        ref = oracle.warp(pts, pose, i.fx, i.fy, i.cx, i.cy, i.invfx, i.invfy)
        assert np.array_equal(t.WarpFunction(pts, pose, lvl), ref)
    t.close()


def run_both(oracle, calib, prev, cur, extra_flags=0, **cfg):
    import uw_slam_b200._lib as L
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    t = make_tracker(calib, flags=L.FLAG_TRACE | extra_flags, **cfg)
    fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    pose, stats = t.EstimatePose(fp, fc, return_stats=True)
    trace = t.get_trace(0)
    t.close()
    p = oracle.default_params(w, h, fx, fy, cx, cy, solve_mode=cfg.get("solve_mode", 0),
                              weight_mode=cfg.get("weight_mode", 0),
                              huber_delta=cfg.get("huber_delta", 10.0))
    rp, rc = oracle.FrameData(prev), oracle.FrameData(cur, with_candidates=False)
    opose, ostats, otrace = oracle.estimate_pose(p, rp, rc)
    return pose[0], stats[0], trace, opose, ostats, otrace


def assert_trace_equal(trace, otrace):
    assert len(trace) == len(otrace)
    for a, b in zip(trace, otrace):
        key = (b.level, b.k)
        assert (a.level, a.k, a.broke) == (b.level, b.k, b.broke), key
        assert a.n_valid == b.n_valid, key            # exact integer
        assert a.sum_r2 == b.sum_r2, key              # exact integer
        assert np.float32(a.error) == np.float32(b.error), key
        assert np.array_equal(np.array(a.A[:]), np.array(b.A[:])), ("A", key)
        assert np.array_equal(np.array(a.b[:]), np.array(b.b[:])), ("b", key)
        assert np.array_equal(np.array(a.delta[:]), np.array(b.delta[:])), ("delta", key)
        assert np.array_equal(np.array(a.pose[:]), np.array(b.pose[:])), ("pose", key)


@pytest.mark.parametrize("calib,seed", [("tiny", 0), ("small", 0), ("small", 1), ("tum", 0),
                                        ("tum", 7), ("euroc", 2), ("tum_mono", 5)])
def test_estimate_pose_matches_oracle(oracle, pairs, calib, seed):
    prev, cur = pairs(calib, seed)
    pose, stats, trace, opose, ostats, otrace = run_both(oracle, calib, prev, cur)
    assert_trace_equal(trace, otrace)
    assert list(stats.iterations)[:5] == list(ostats.iterations)[:5]
    assert list(stats.n_points)[:5] == list(ostats.n_points)[:5]
    assert np.array_equal(pose, opose)
    assert pose_close(pose, opose, 1e-5)


def test_estimate_pose_inverse_solve_mode(oracle, pairs):
    prev, cur = pairs("tum", 1)
    pose, _, trace, opose, _, otrace = run_both(oracle, "tum", prev, cur, solve_mode=1)
    assert_trace_equal(trace, otrace)
    assert np.array_equal(pose, opose)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_cluster_sizes_agree_with_oracle(oracle, pairs, cluster):
    prev, cur = pairs("tum", 3)
    pose, _, trace, opose, _, otrace = run_both(oracle, "tum", prev, cur, cluster_size=cluster)
    assert_trace_equal(trace, otrace)
    assert np.array_equal(pose, opose)


# the robust-weight path (Tracker.cpp:496, 1571-1654) and the Huber option
@pytest.mark.parametrize("mode,calib,seed,cluster", [
    (1, "tiny", 0, 0), (1, "small", 1, 0), (1, "tum", 0, 0), (1, "tum", 7, 1), (1, "tum", 3, 4),
    (1, "euroc", 2, 0), (1, "tum_mono", 5, 0),
    (2, "small", 0, 0), (2, "tum", 2, 0), (2, "tum", 2, 2), (2, "tum_mono", 5, 0)])
def test_robust_weights_match_oracle(oracle, pairs, mode, calib, seed, cluster):
    prev, cur = pairs(calib, seed)
    pose, stats, trace, opose, ostats, otrace = run_both(
        oracle, calib, prev, cur, weight_mode=mode, huber_delta=7.5, cluster_size=cluster)
    assert_trace_equal(trace, otrace)
    assert list(stats.iterations)[:5] == list(ostats.iterations)[:5]
    assert np.array_equal(pose, opose)
    # the weights must actually change the estimate relative to the shipped identity weights
    pose_id = run_both(oracle, calib, prev, cur, cluster_size=cluster)[0]
    assert not np.array_equal(pose, pose_id)


def test_robust_weights_with_outliers(oracle, pairs):
    # an occluder in the second frame: Tukey must down-weight it, and still match the oracle
    prev, cur = pairs("tum", 4)
    cur = cur.copy()
    cur[100:220, 200:360] = 255 - cur[100:220, 200:360]
    for mode in (1, 2):
        pose, _, trace, opose, _, otrace = run_both(oracle, "tum", prev, cur, weight_mode=mode)
        assert_trace_equal(trace, otrace)
        assert np.array_equal(pose, opose)


def test_robust_weights_batch_and_rejections(oracle, pairs):
    import uw_slam_b200 as U
    import uw_slam_b200._lib as L
    calib = "small"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    B = 12
    t = make_tracker(calib, max_frames=2 * B, weight_mode=L.WEIGHT_TUKEY)
    prevs = np.stack([pairs(calib, s)[0] for s in range(B)])
    curs = np.stack([pairs(calib, s)[1] for s in range(B)])
    fp = t.AddFrames(list(range(B)), prevs)
    fc = t.AddFrames(list(range(B, 2 * B)), curs)
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    poses = t.EstimatePose(fp, fc)
    p = oracle.default_params(w, h, fx, fy, cx, cy, weight_mode=1)
    for s in range(B):
        rp = oracle.FrameData(prevs[s])
        rc = oracle.FrameData(curs[s], with_candidates=False)
        assert np.array_equal(poses[s], oracle.estimate_pose(p, rp, rc)[0]), s
    # the sharded single-frame mode is identity-weights only and says so
    rc = t._lib.uwt_shard_begin(t._h, 0, B, 0, 1, None)
    assert rc == L.E_INVALID
    t.close()
    with pytest.raises(U.UwtError):
        make_tracker(calib, weight_mode=7)
    with pytest.raises(U.UwtError):
        make_tracker(calib, weight_mode=L.WEIGHT_HUBER, huber_delta=0.0)
    with pytest.raises(U.UwtError):
        make_tracker(calib, weight_mode=L.WEIGHT_TUKEY, flags=L.FLAG_DMMA_ACCUM)


def test_dataflow_kernel_traces_match_oracle(oracle, pairs):
    """Batches of 24 .. 295 problems run on the persistent dataflow kernel (chunk tasks from a global
    ring): every per-sweep quantity of every problem must equal the oracle's, including problems
    without candidate points (ARITHMETIC.md U2) and identical frames."""
    import uw_slam_b200._lib as L
    calib = "small"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    B = 40
    prevs = [pairs(calib, s)[0] for s in range(B)]
    curs = [pairs(calib, s)[1] for s in range(B)]
    prevs[3] = np.full((h, w), 90, np.uint8)          # flat: no candidates on any level
    curs[7] = prevs[7].copy()                         # identical frames
    prevs[11] = prevs[11].copy()
    prevs[11][: h // 2] = 17                          # half flat: fewer points, empty coarse rows
    t = make_tracker(calib, max_frames=2 * B, flags=L.FLAG_TRACE)
    fp = t.AddFrames(list(range(B)), np.stack(prevs))
    fc = t.AddFrames(list(range(B, 2 * B)), np.stack(curs))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    init = np.tile(np.array([0, 0, 0, 1, 0, 0, 0], np.float32), (B, 1))
    init[5] = oracle.se3_exp(np.array([1e-3, -2e-3, 1e-3, 2e-3, 1e-3, -1e-3], np.float32))
    poses, stats = t.EstimatePose(fp, fc, init_poses=init, return_stats=True)
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    for s in range(B):
        rp = oracle.FrameData(prevs[s])
        rc = oracle.FrameData(curs[s], with_candidates=False)
        opose, ostats, otrace = oracle.estimate_pose(p, rp, rc, init_pose=init[s])
        assert_trace_equal(t.get_trace(s), otrace)
        assert np.array_equal(poses[s], opose), s
        assert list(stats[s].iterations)[:5] == list(ostats.iterations)[:5], s
        assert list(stats[s].evaluations)[:5] == list(ostats.evaluations)[:5], s
        assert list(stats[s].n_points)[:5] == list(ostats.n_points)[:5], s
    t.close()


@pytest.mark.parametrize("calib,B", [("small", 36), ("tum", 24)])
@pytest.mark.parametrize("mode", [1, 2])
def test_dataflow_kernel_robust_weights_match_oracle(oracle, pairs, mode, calib, B):
    """Robust weights on the dataflow kernel (>= 24 problems): Huber uses one table per CTA; a
    Tukey sweep is a histogram round and an accumulation round of chunk tasks.  Every per-sweep
    quantity must equal the oracle's, also for problems with no or few points and with an
    occluder, and the cluster kernel must give the same bits."""
    import uw_slam_b200._lib as L
    w, h, fx, fy, cx, cy = synth.CALIB[calib]   # "tum": several chunks per sweep on level 1
    prevs = [pairs(calib, s % 6)[0] for s in range(B)]
    curs = [pairs(calib, s % 6)[1].copy() for s in range(B)]
    prevs[2] = np.full((h, w), 120, np.uint8)         # no candidates at all
    curs[4] = prevs[4].copy()                         # identical frames: all residuals zero
    curs[6][h // 4: h // 2, w // 4: w // 2] = 255     # occluder: the outliers Tukey rejects
    prevs[9] = prevs[9].copy()
    prevs[9][: h // 2] = 33
    out = []
    for flags in (L.FLAG_TRACE, L.FLAG_TRACE | L.FLAG_CLUSTER_KERNEL):
        t = make_tracker(calib, max_frames=2 * B, flags=flags, weight_mode=mode, huber_delta=6.0)
        fp = t.AddFrames(list(range(B)), np.stack(prevs))
        fc = t.AddFrames(list(range(B, 2 * B)), np.stack(curs))
        t.ApplyGradient(fp)
        t.ObtainCandidatePoints(fp)
        for rep in range(2):   # the histograms are re-armed per launch
            poses, stats = t.EstimatePose(fp, fc, return_stats=True)
        out.append((poses, [t.get_trace(s) for s in range(B)]))
        t.close()
    assert np.array_equal(out[0][0], out[1][0])
    p = oracle.default_params(w, h, fx, fy, cx, cy, weight_mode=mode, huber_delta=6.0)
    for s in range(B):
        rp = oracle.FrameData(prevs[s])
        rc = oracle.FrameData(curs[s], with_candidates=False)
        opose, ostats, otrace = oracle.estimate_pose(p, rp, rc)
        assert_trace_equal(out[0][1][s], otrace)
        assert_trace_equal(out[1][1][s], otrace)
        assert np.array_equal(out[0][0][s], opose), s


def test_dataflow_and_cluster_kernels_agree_bitwise(pairs):
    import uw_slam_b200._lib as L
    calib = "tum_mono"
    B = 32
    prevs = np.stack([pairs(calib, s % 4)[0] for s in range(B)])
    curs = np.stack([pairs(calib, (s * 7) % 5)[1] for s in range(B)])
    out = []
    for flags in (0, L.FLAG_CLUSTER_KERNEL):
        t = make_tracker(calib, max_frames=2 * B, flags=flags)
        fp = t.AddFrames(list(range(B)), prevs)
        fc = t.AddFrames(list(range(B, 2 * B)), curs)
        t.ApplyGradient(fp)
        t.ObtainCandidatePoints(fp)
        for rep in range(3):   # the ring and the control block are re-armed per launch
            poses, stats = t.EstimatePose(fp, fc, return_stats=True)
        out.append((poses, [list(s.evaluations) + list(s.n_points) for s in stats]))
        t.close()
    assert np.array_equal(out[0][0], out[1][0])
    assert out[0][1] == out[1][1]


def test_dmma_accumulator_variant_matches_oracle(oracle, pairs):
    import uw_slam_b200._lib as L
    prev, cur = pairs("tum", 5)
    pose, _, trace, opose, _, otrace = run_both(oracle, "tum", prev, cur,
                                                extra_flags=L.FLAG_DMMA_ACCUM)
    assert_trace_equal(trace, otrace)
    assert np.array_equal(pose, opose)
    pose, _, trace, opose, _, otrace = run_both(oracle, "tum", prev, cur, cluster_size=1,
                                                extra_flags=L.FLAG_DMMA_ACCUM)
    assert_trace_equal(trace, otrace)


def test_batch_of_independent_pairs(oracle, pairs):
    calib = "small"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    B = 24
    t = make_tracker(calib, max_frames=2 * B)
    prevs = np.stack([pairs(calib, s)[0] for s in range(B)])
    curs = np.stack([pairs(calib, s)[1] for s in range(B)])
    ps, cs = list(range(B)), list(range(B, 2 * B))
    t.AddFrames(ps, prevs)
    t.AddFrames(cs, curs)
    t.ApplyGradient(ps)
    t.ObtainCandidatePoints(ps)
    poses = t.EstimatePose(ps, cs)
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    for s in range(B):
        rp, rc = oracle.FrameData(prevs[s]), oracle.FrameData(curs[s], with_candidates=False)
        opose, _, _ = oracle.estimate_pose(p, rp, rc)
        assert np.array_equal(poses[s], opose), s
    # uploads 2 (fused pyramid + gradient kernel: ApplyGradient has nothing left to launch),
    # candidates 4 (count, scan, scatter, bitmask), estimate 2 (ring init + dataflow kernel)
    assert t.launch_count() == 2 + 0 + 4 + 2
    t.close()


def test_device_resident_input_and_sequence(oracle):
    torch = pytest.importorskip("torch")
    calib = "small"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    frames, _, _ = synth.render_sequence(calib, 2, 4)
    dev = torch.from_numpy(np.stack(frames)).cuda()
    t = make_tracker(calib, max_frames=2)
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    t.AddFramesDevice([0], dev[0].data_ptr())
    t.ApplyGradient([0])
    t.ObtainCandidatePoints([0])
    for i in range(1, 4):  # frame-to-frame: System::Tracking direct order
        prev_slot, cur_slot = (i - 1) % 2, i % 2
        t.AddFramesDevice([cur_slot], dev[i].data_ptr())
        pose = t.EstimatePose([prev_slot], [cur_slot])[0]
        t.ApplyGradient([cur_slot])
        t.ObtainCandidatePoints([cur_slot])
        rp = oracle.FrameData(frames[i - 1])
        rc = oracle.FrameData(frames[i], with_candidates=False)
        opose, _, _ = oracle.estimate_pose(p, rp, rc)
        assert np.array_equal(pose, opose), i
    t.close()


def test_state_errors_are_reported():
    import uw_slam_b200 as U
    t = make_tracker("tiny")
    with pytest.raises(U.UwtError) as e:
        t.ApplyGradient([0])
    assert e.value.code == -3
    with pytest.raises(U.UwtError):
        t.AddFrames([5], np.zeros((48, 64), np.uint8))
    t.close()
    with pytest.raises(U.UwtError):
        U.Tracker(False).InitializePyramid(100, 100, np.eye(3, dtype=np.float32))


def test_sharded_mode_single_rank_equals_estimate_pose(oracle, pairs):
    from uw_slam_b200.sharded import TrackerShardBackend, estimate_pose_sharded
    pytest.importorskip("torch")
    calib = "tum"
    prev, cur = pairs(calib, 4)
    t = make_tracker(calib)
    fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    ref_pose, ref_stats = t.EstimatePose(fp, fc, return_stats=True)
    pose, stats, sweeps = estimate_pose_sharded(TrackerShardBackend(t, 0, 1))
    assert np.array_equal(pose, ref_pose[0])
    assert list(stats.iterations)[:5] == list(ref_stats[0].iterations)[:5]
    assert sweeps == sum(ref_stats[0].evaluations)
    t.close()


@pytest.mark.parametrize("nranks", [2, 3])
def test_sharded_mode_emulated_ranks_match_oracle(oracle, pairs, nranks):
    # nranks handles on ONE GPU play the ranks; the "all-reduce" is a torch sum of the partial
    # sums.  Checks the range partition inside the kernels and the redundant update.
    torch = pytest.importorskip("torch")
    calib = "euroc"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    prev, cur = pairs(calib, 6)
    ts = [make_tracker(calib) for _ in range(nranks)]
    sums = [torch.zeros(32, dtype=torch.float64, device="cuda") for _ in range(nranks)]
    for r, t in enumerate(ts):
        fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
        t.ApplyGradient(fp)
        t.ObtainCandidatePoints(fp)
        t.ShardBegin(0, 1, r, nranks)
    for _ in range(400):
        for r, t in enumerate(ts):
            t.ShardAccumulate(sums[r].data_ptr())
            t.synchronize()
        total = torch.stack(sums).sum(0)
        torch.cuda.synchronize()
        done = [t.ShardUpdate(total.data_ptr()) for t in ts]
        assert len(set(done)) == 1
        if done[0]:
            break
    poses = [t.ShardResult()[0] for t in ts]
    rp, rc = oracle.FrameData(prev), oracle.FrameData(cur, with_candidates=False)
    opose, _, _ = oracle.estimate_pose(oracle.default_params(w, h, fx, fy, cx, cy), rp, rc)
    for p in poses:
        assert np.array_equal(p, opose)
    for t in ts:
        t.close()


def test_fused_sharded_single_rank_equals_estimate_pose(pairs):
    from uw_slam_b200.sharded import connect_fused, estimate_pose_sharded_fused
    calib = "tum"
    prev, cur = pairs(calib, 4)
    t = make_tracker(calib)
    fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    ref_pose, ref_stats = t.EstimatePose(fp, fc, return_stats=True)
    connect_fused(t)
    for _ in range(3):  # sequence numbers keep running across calls
        pose, stats = estimate_pose_sharded_fused(t, 0, 1)
        assert np.array_equal(pose, ref_pose[0])
        assert list(stats.iterations)[:5] == list(ref_stats[0].iterations)[:5]
    t.close()


def test_fused_sharded_two_emulated_ranks_match_oracle(oracle, pairs):
    # two handles on ONE GPU play two ranks: their persistent kernels run concurrently (64 CTAs
    # each) and exchange the sums through each other's mailboxes, exactly as two GPUs would
    # over NVLink.  Every wait in the kernel is bounded, so a failure cannot hang the device.
    calib = "euroc"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    prev, cur = pairs(calib, 6)
    ts = [make_tracker(calib) for _ in range(2)]
    for t in ts:
        fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
        t.ApplyGradient(fp)
        t.ObtainCandidatePoints(fp)
    for r, t in enumerate(ts):
        t.ShardConnectLocal(r, ts)
    rp, rc = oracle.FrameData(prev), oracle.FrameData(cur, with_candidates=False)
    opose, _, _ = oracle.estimate_pose(oracle.default_params(w, h, fx, fy, cx, cy), rp, rc)
    for _ in range(2):
        for t in ts:
            t.ShardEstimateFusedAsync(0, 1, grid=64)
        poses = [t.ShardEstimateFusedWait()[0] for t in ts]
        assert np.array_equal(poses[0], poses[1])
        assert np.array_equal(poses[0], opose)
    for t in ts:
        t.close()


def test_sharded_mode_huber_weights_match_oracle(oracle, pairs):
    """Huber weights in the sharded mode (both forms): the weight is a fixed function of the
    integer residual, so every rank builds the same tables and the 32 exchanged sums carry the
    weighted error term.  Two emulated ranks per form against the oracle with the same weights;
    Tukey / MAD weights (a histogram all-reduce per sweep) stay rejected."""
    torch = pytest.importorskip("torch")
    import uw_slam_b200 as U
    import uw_slam_b200._lib as L
    calib = "euroc"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    prev, cur = pairs(calib, 6)
    cur = cur.copy()
    cur[100:220, 300:520] = 255          # an occluder: the weights must matter
    rp, rc = oracle.FrameData(prev), oracle.FrameData(cur, with_candidates=False)
    opose, _, _ = oracle.estimate_pose(
        oracle.default_params(w, h, fx, fy, cx, cy, weight_mode=2, huber_delta=6.0), rp, rc)
    ident, _, _ = oracle.estimate_pose(oracle.default_params(w, h, fx, fy, cx, cy), rp, rc)
    assert not np.array_equal(opose, ident)
    ts = [make_tracker(calib, weight_mode=L.WEIGHT_HUBER, huber_delta=6.0) for _ in range(2)]
    for t in ts:
        fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
        t.ApplyGradient(fp)
        t.ObtainCandidatePoints(fp)
    # NCCL-form kernels, the all-reduce played by a torch sum
    sums = [torch.zeros(32, dtype=torch.float64, device="cuda") for _ in ts]
    for r, t in enumerate(ts):
        t.ShardBegin(0, 1, r, 2)
    for _ in range(400):
        for r, t in enumerate(ts):
            t.ShardAccumulate(sums[r].data_ptr())
            t.synchronize()
        total = torch.stack(sums).sum(0)
        torch.cuda.synchronize()
        done = [t.ShardUpdate(total.data_ptr()) for t in ts]
        if done[0]:
            break
    for t in ts:
        assert np.array_equal(t.ShardResult()[0], opose)
    # fused form: peer mailboxes
    for r, t in enumerate(ts):
        t.ShardConnectLocal(r, ts)
    for t in ts:
        t.ShardEstimateFusedAsync(0, 1, grid=64)
    poses = [t.ShardEstimateFusedWait()[0] for t in ts]
    assert np.array_equal(poses[0], opose) and np.array_equal(poses[1], opose)
    for t in ts:
        t.close()
    tk = make_tracker(calib, weight_mode=L.WEIGHT_TUKEY)
    fp, fc = tk.AddFrames([0, 1], np.stack([prev, cur]))
    tk.ApplyGradient(fp)
    tk.ObtainCandidatePoints(fp)
    with pytest.raises(U.UwtError):
        tk.ShardBegin(0, 1, 0, 1)
    tk.close()


def test_two_devices_in_one_process(oracle, pairs):
    """Handles on different GPUs of one process (function attributes are per device): both must
    track, on every estimate kernel.  Needs >= 2 visible GPUs."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    calib = "tum"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    B = 32
    prevs = np.stack([pairs(calib, s % 3)[0] for s in range(B)])
    curs = np.stack([pairs(calib, s % 3)[1] for s in range(B)])
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    ref = [oracle.estimate_pose(p, oracle.FrameData(prevs[s]),
                                oracle.FrameData(curs[s], with_candidates=False))[0]
           for s in range(3)]
    for dev in (1, 0):
        t = make_tracker(calib, max_frames=2 * B, device=dev)
        fp = t.AddFrames(list(range(B)), prevs)
        fc = t.AddFrames(list(range(B, 2 * B)), curs)
        t.ApplyGradient(fp)
        t.ObtainCandidatePoints(fp)
        poses = t.EstimatePose(fp, fc)                       # dataflow kernel
        one = t.EstimatePose(fp[:1], fc[:1])                 # 16-CTA cluster kernel
        for s in range(B):
            assert np.array_equal(poses[s], ref[s % 3]), (dev, s)
        assert np.array_equal(one[0], ref[0]), dev
        t.close()


def test_lazy_levels_flag_is_observationally_identical(oracle, pairs):
    """UWT_FLAG_LAZY_LEVELS: gradient / candidates only on the optimised levels; the other levels
    appear on read-back, with the values the eager mode (and the oracle) has."""
    import uw_slam_b200._lib as L
    calib = "euroc"
    prev, cur = pairs(calib, 6)
    ref = oracle.FrameData(prev)
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    opose = oracle.estimate_pose(oracle.default_params(w, h, fx, fy, cx, cy), ref,
                                 oracle.FrameData(cur, with_candidates=False))[0]
    t = make_tracker(calib, flags=L.FLAG_LAZY_LEVELS)
    fp, fc = t.AddFrames([0, 1], np.stack([prev, cur]))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    n0 = t.launch_count()
    assert np.array_equal(t.EstimatePose(fp, fc)[0], opose)
    # optimised levels are there without extra work ...
    for lvl in (1, 4):
        assert np.array_equal(fp.candidatePoints(lvl), ref.cand[lvl])
        assert np.array_equal(fp.gradients(lvl)[2], ref.g[lvl])
    # ... level 0 is materialised on demand (gradient image, then the candidate list)
    n1 = t.launch_count()
    gx, gy, g = fp.gradients(0)
    assert np.array_equal(g, ref.g[0]) and np.array_equal(gx, ref.gx[0])
    assert np.array_equal(fp.candidatePoints(0), ref.cand[0])
    assert t.launch_count() > n1 and n1 > n0
    # nothing changed for the tracker
    assert np.array_equal(t.EstimatePose(fp, fc)[0], opose)
    for lvl in range(5):
        assert np.array_equal(fp.candidatePoints(lvl), ref.cand[lvl])
    t.close()


def test_dataflow_kernel_sparse_candidates(oracle):
    """A frame whose candidates are spread thinly over all columns: the warp-level dataflow kernel
    has to move its 32-column transform window inside a row of 32 records."""
    import uw_slam_b200._lib as L
    calib = "tum"
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    rng = np.random.default_rng(9)
    B = 24
    prevs, curs = [], []
    for s in range(B):
        img = np.full((h, w), 100, np.uint8)
        # isolated bright dots, about one per 3 columns and level-0 row block: strong gradients
        # on a flat background at every pyramid level
        ys = rng.integers(8, h - 8, 700)
        xs = rng.integers(8, w - 8, 700)
        for y, x in zip(ys, xs):
            img[y - 4:y + 4, x - 4:x + 4] = 255
        prevs.append(img)
        curs.append(np.roll(img, (1, 2), (0, 1)))
    t = make_tracker(calib, max_frames=2 * B, flags=L.FLAG_TRACE)
    fp = t.AddFrames(list(range(B)), np.stack(prevs))
    fc = t.AddFrames(list(range(B, 2 * B)), np.stack(curs))
    t.ApplyGradient(fp)
    t.ObtainCandidatePoints(fp)
    poses = t.EstimatePose(fp, fc)
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    for s in range(0, B, 5):
        rp = oracle.FrameData(prevs[s])
        rc = oracle.FrameData(curs[s], with_candidates=False)
        opose, _, otrace = oracle.estimate_pose(p, rp, rc)
        assert_trace_equal(t.get_trace(s), otrace)
        assert np.array_equal(poses[s], opose), s
    t.close()


def test_dataflow_traces_at_headline_size(oracle):
    """The headline configuration itself (1280x1024, 16 chunk tasks per level-1 sweep, identity
    weights, a batch on the dataflow kernel): every sweep of four problems against the oracle --
    N_valid, sum r^2, error, A, b, delta, pose -- bit for bit, and every final pose."""
    import uw_slam_b200._lib as L
    calib, B = "tum_mono", 24
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    pairs = [synth.render_pair(calib, 300 + i, rot=5e-3 * (0.5 + 0.05 * i),
                               trans=5e-3 * (1.4 - 0.04 * i))[:2] for i in range(4)]
    t = make_tracker(calib, max_frames=2 * B, flags=L.FLAG_TRACE)
    prevs = np.stack([pairs[i % 4][0] for i in range(B)])
    curs = np.stack([pairs[i % 4][1] for i in range(B)])
    ps, cs = list(range(B)), list(range(B, 2 * B))
    t.AddFrames(ps, prevs)
    t.AddFrames(cs, curs)
    t.ApplyGradient(ps)
    t.ObtainCandidatePoints(ps)
    poses = t.EstimatePose(ps, cs)
    p = oracle.default_params(w, h, fx, fy, cx, cy)
    for i in range(4):
        rp, rc = oracle.FrameData(pairs[i][0]), oracle.FrameData(pairs[i][1], with_candidates=False)
        opose, _, otr = oracle.estimate_pose(p, rp, rc)
        for j in range(i, B, 4):                      # the same pair in several queue positions
            assert np.array_equal(poses[j], opose), (i, j)
        tr = t.get_trace(i)
        assert len(tr) == len(otr) and len(tr) >= 8
        assert max(a.n_valid for a in tr) > 8192      # multi-chunk sweeps were exercised
        for a, b in zip(tr, otr):
            assert (a.level, a.k, a.n_valid, a.broke, a.sum_r2) == \
                (b.level, b.k, b.n_valid, b.broke, b.sum_r2), (i, b.level, b.k)
            for f in ("A", "b", "delta", "pose"):
                assert np.array_equal(np.array(getattr(a, f)[:]), np.array(getattr(b, f)[:])), \
                    (i, b.level, b.k, f)
            assert np.float32(a.error) == np.float32(b.error)
    t.close()
