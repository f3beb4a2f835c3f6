"""The fused frame kernel (pyramid + gradient images of all levels from one tensor-copy-staged
read) against the CPU oracle and against the separate K1 / K2 kernels, bit for bit; plus the
direct check of Tracker::InitializePyramid's per-level intrinsics."""
import numpy as np
import pytest

from uw_slam_b200 import synth

pytestmark = pytest.mark.gpu


def tracker(w, h, **cfg):
    import uw_slam_b200 as U
    t = U.Tracker(False)
    cfg.setdefault("max_frames", 4)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, 500.0, 510.0, w / 2 - 0.5,
                                                            h / 2 - 0.5).GetK(), **cfg)
    return t


# sizes chosen for the tile logic: one partial tile in x and y; image border inside the first
# tile on both sides; exact multiples of the 128-pixel tile; tall and wide
SIZES = [(64, 48), (160, 128), (144, 272), (400, 144), (256, 256), (752, 480), (1280, 1024)]


@pytest.mark.parametrize("w,h", SIZES)
def test_fused_equals_oracle_and_separate_kernels(oracle, w, h):
    import uw_slam_b200._lib as L
    rng = np.random.default_rng(w * 7 + h)
    # white noise (every saturation / tie / border path) and a smooth ramp (low gradients)
    yy, xx = np.mgrid[0:h, 0:w]
    imgs = [rng.integers(0, 256, (h, w), dtype=np.uint8),
            ((xx * 3 + yy * 5) % 256).astype(np.uint8)]
    fused, sep = tracker(w, h), tracker(w, h, flags=L.FLAG_SEPARATE_GRADIENT)
    for k, img in enumerate(imgs):
        ref = oracle.FrameData(img)
        lf0 = fused.launch_count()
        ff = fused.AddFrames([k], img)[0]
        assert fused.launch_count() - lf0 == 1          # ONE kernel: pyramid + gradients
        fused.ApplyGradient(ff)
        assert fused.launch_count() - lf0 == 1          # nothing left to launch
        fused.ObtainCandidatePoints(ff)
        fs = sep.AddFrames([k], img)[0]
        sep.ApplyGradient(fs)
        sep.ObtainCandidatePoints(fs)
        for lvl in range(5):
            assert np.array_equal(ff.image(lvl), ref.images[lvl]), ("image", lvl)
            assert np.array_equal(fs.image(lvl), ref.images[lvl]), ("image/separate", lvl)
            gx, gy, g = ff.gradients(lvl)
            assert np.array_equal(g, ref.g[lvl]), ("g", lvl)
            assert np.array_equal(gx, ref.gx[lvl]) and np.array_equal(gy, ref.gy[lvl])
            assert np.array_equal(fs.gradients(lvl)[2], ref.g[lvl]), ("g/separate", lvl)
            assert np.array_equal(ff.candidatePoints(lvl), ref.cand[lvl]), ("cand", lvl)
            assert np.array_equal(fs.candidatePoints(lvl), ref.cand[lvl]), ("cand/separate", lvl)
    fused.close()
    sep.close()


def test_fused_batch_strided_source_and_fallback(oracle):
    """Frames inside a larger pinned allocation (row stride > width) go through the tensor map;
    a source whose stride is not a multiple of 16 falls back to the separate kernels."""
    import ctypes as C
    import uw_slam_b200._lib as L
    w, h = 160, 128
    rng = np.random.default_rng(5)
    n = 3
    for stride in (w + 32, w + 8):
        big = rng.integers(0, 256, (n, h, stride), dtype=np.uint8)
        t = tracker(w, h)
        l0 = t.launch_count()
        t.AddFramesHostPtr(list(range(n)), big.ctypes.data, stride, stride * h)
        # staging is dense (the H2D copy packs rows), so both strides take the fused kernel
        assert t.launch_count() - l0 == 1
        t.ApplyGradient(list(range(n)))
        t.ObtainCandidatePoints(list(range(n)))
        for k in range(n):
            ref = oracle.FrameData(np.ascontiguousarray(big[k, :, :w]))
            for lvl in range(5):
                assert np.array_equal(t.get_image(k, lvl), ref.images[lvl])
                assert np.array_equal(t.get_gradients(k, lvl)[2], ref.g[lvl])
                assert np.array_equal(t.get_candidates(k, lvl), ref.cand[lvl])
        t.close()


def test_fused_device_source_unaligned_falls_back(oracle):
    import torch
    w, h = 160, 128
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(3)
    buf = torch.randint(0, 256, (h * w + 64,), dtype=torch.uint8, generator=g).to(dev)
    torch.cuda.synchronize()
    for off, fused in ((0, True), (16, True), (4, False)):
        t = tracker(w, h)
        l0 = t.launch_count()
        t.AddFramesDevice([0], buf.data_ptr() + off)
        assert t.launch_count() - l0 == 1
        t.ApplyGradient([0])
        launched_gradient = t.launch_count() - l0 == 2
        assert launched_gradient == (not fused)
        t.ObtainCandidatePoints([0])
        img = buf[off:off + h * w].cpu().numpy().reshape(h, w)
        ref = oracle.FrameData(img)
        for lvl in range(5):
            assert np.array_equal(t.get_image(0, lvl), ref.images[lvl])
            assert np.array_equal(t.get_gradients(0, lvl)[2], ref.g[lvl])
            assert np.array_equal(t.get_candidates(0, lvl), ref.cand[lvl])
        t.close()


@pytest.mark.parametrize("calib", ["tiny", "tum", "euroc", "tum_mono", "uhd"])
def test_level_info_equals_initialize_pyramid(oracle, calib):
    """Row a2: uwt_get_level_info returns Tracker::InitializePyramid's per-level values
    (Tracker.cpp:297-340) -- checked directly against the oracle's restatement, bit for bit."""
    import uw_slam_b200 as U
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK())
    ref = oracle.init_pyramid(w, h, fx, fy, cx, cy, 5)
    for lvl in range(5):
        i = t.level_info(lvl)
        got = np.array([i.fx, i.fy, i.cx, i.cy, i.invfx, i.invfy], np.float32)
        exp = np.array([ref[k][lvl] for k in ("fx", "fy", "cx", "cy", "invfx", "invfy")],
                       np.float32)
        assert (i.width, i.height) == (int(ref["w"][lvl]), int(ref["h"][lvl]))
        assert np.array_equal(got.view(np.int32), exp.view(np.int32)), (lvl, got, exp)
    t.close()


def test_trace_state_after_untraced_batch():
    """uwt_get_trace after a batch larger than the trace capacity reports UWT_E_STATE instead of
    handing out an older batch's rows."""
    import uw_slam_b200 as U
    import uw_slam_b200._lib as L
    w, h, fx, fy, cx, cy = synth.CALIB["tiny"]
    n = 65
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2 * n, flags=L.FLAG_TRACE)
    prev, cur, _, _ = synth.render_pair("tiny", 1)
    t.AddFrames(list(range(n)), np.stack([prev] * n))
    t.AddFrames(list(range(n, 2 * n)), np.stack([cur] * n))
    t.ApplyGradient(list(range(n)))
    t.ObtainCandidatePoints(list(range(n)))
    t.EstimatePose(list(range(8)), list(range(n, n + 8)))
    assert len(t.get_trace(0)) > 0
    t.EstimatePose(list(range(n)), list(range(n, 2 * n)))
    with pytest.raises(U.tracker.UwtError) as e:
        t.get_trace(0)
    assert e.value.code == L.E_STATE
    t.close()
