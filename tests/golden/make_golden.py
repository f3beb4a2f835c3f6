"""Generates the committed golden vectors of the hot path.

The reference ships no golden vectors, so they are produced here from the
CPU oracle AFTER checking, on the very same inputs, that every OpenCV primitive of the oracle
equals the real OpenCV (cv2): the integer image outputs stored below are therefore cv2's own.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

from oracle import uw_oracle as O  # noqa: E402
from uw_slam_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def check_against_cv2(img, fd):
    for l in range(5):
        if l:
            assert np.array_equal(fd.images[l], cv2.resize(fd.images[l - 1], None, fx=0.5, fy=0.5))
        gx = cv2.Scharr(fd.images[l], cv2.CV_16S, 1, 0)
        gy = cv2.Scharr(fd.images[l], cv2.CV_16S, 0, 1)
        assert np.array_equal(fd.gx[l], gx) and np.array_equal(fd.gy[l], gy)
        g = cv2.addWeighted(cv2.convertScaleAbs(gx), 0.5, cv2.convertScaleAbs(gy), 0.5, 0)
        assert np.array_equal(fd.g[l], g)
        m, _ = cv2.meanStdDev(g)
        _, filt = cv2.threshold(g, float(np.float32(m[0, 0] + 20.0)), 255, cv2.THRESH_BINARY)
        xs, ys = np.nonzero(filt.T)
        ref = np.stack([xs, ys, np.ones_like(xs), np.ones_like(xs)], 1).astype(np.float32)
        assert np.array_equal(fd.cand[l], ref)


def trace_arrays(trace):
    return {
        "lvl_k_nvalid_broke": np.array([[t.level, t.k, t.n_valid, t.broke] for t in trace], np.int32),
        "sum_r2": np.array([t.sum_r2 for t in trace], np.int64),
        "error": np.array([t.error for t in trace], np.float32),
        "A": np.array([t.A[:] for t in trace], np.float32),
        "b": np.array([t.b[:] for t in trace], np.float32),
        "delta": np.array([t.delta[:] for t in trace], np.float32),
        "pose": np.array([t.pose[:] for t in trace], np.float32),
    }


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    full = {}
    for calib, seeds in (("tiny", [0, 1]), ("small", [0, 1, 2])):
        w, h, fx, fy, cx, cy = synth.CALIB[calib]
        for s in seeds:
            prev, cur, _, _ = synth.render_pair(calib, s)
            fp, fc = O.FrameData(prev), O.FrameData(cur, with_candidates=False)
            check_against_cv2(prev, fp)
            for mode in (0, 1):
                p = O.default_params(w, h, fx, fy, cx, cy, solve_mode=mode)
                pose, st, tr = O.estimate_pose(p, fp, fc)
                key = "%s_%d_m%d" % (calib, s, mode)
                full[key + "_final"] = pose
                for k, v in trace_arrays(tr).items():
                    full[key + "_" + k] = v
            key = "%s_%d" % (calib, s)
            full[key + "_prev"], full[key + "_cur"] = prev, cur
            for l in range(5):
                full["%s_img%d" % (key, l)] = fp.images[l]
                full["%s_gx%d" % (key, l)] = fp.gx[l]
                full["%s_gy%d" % (key, l)] = fp.gy[l]
                full["%s_g%d" % (key, l)] = fp.g[l]
                full["%s_cand%d" % (key, l)] = fp.cand[l][:, :2].astype(np.uint16)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **full)

    # larger configs: inputs are regenerated from the seed, outputs pinned by hash + values
    big = {}
    for calib, seeds in (("tum", [0, 1, 2, 3]), ("euroc", [0, 1]), ("tum_mono", [0])):
        w, h, fx, fy, cx, cy = synth.CALIB[calib]
        for s in seeds:
            prev, cur, _, _ = synth.render_pair(calib, s)
            fp, fc = O.FrameData(prev), O.FrameData(cur, with_candidates=False)
            check_against_cv2(prev, fp)
            p = O.default_params(w, h, fx, fy, cx, cy)
            pose, st, tr = O.estimate_pose(p, fp, fc)
            key = "%s_%d" % (calib, s)
            big[key + "_input_sha"] = np.array([sha(prev), sha(cur)])
            big[key + "_level_sha"] = np.array(
                [[sha(fp.images[l]), sha(fp.gx[l]), sha(fp.gy[l]), sha(fp.g[l]), sha(fp.cand[l])]
                 for l in range(5)])
            big[key + "_ncand"] = np.array([c.shape[0] for c in fp.cand], np.int32)
            big[key + "_final"] = pose
            big[key + "_iterations"] = np.array(list(st.iterations)[:5], np.int32)
            for k, v in trace_arrays(tr).items():
                big[key + "_" + k] = v
    np.savez_compressed(os.path.join(HERE, "golden_big.npz"), **big)
    for f in ("golden_small.npz", "golden_big.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


def main_robust():
    """Golden traces of the robust-weight modes: Tukey/MAD as in
    Tracker.cpp:496,1571-1654 and the Huber option; inputs are regenerated from the seed."""
    rob = {}
    for calib, seeds in (("tiny", [0]), ("small", [1, 2]), ("tum", [0, 4])):
        w, h, fx, fy, cx, cy = synth.CALIB[calib]
        for s in seeds:
            prev, cur, _, _ = synth.render_pair(calib, s)
            fp, fc = O.FrameData(prev), O.FrameData(cur, with_candidates=False)
            for mode, name in ((O.WEIGHT_TUKEY, "tukey"), (O.WEIGHT_HUBER, "huber")):
                p = O.default_params(w, h, fx, fy, cx, cy, weight_mode=mode, huber_delta=7.5)
                pose, st, tr = O.estimate_pose(p, fp, fc)
                key = "%s_%d_%s" % (calib, s, name)
                rob[key + "_input_sha"] = np.array([sha(prev), sha(cur)])
                rob[key + "_final"] = pose
                rob[key + "_iterations"] = np.array(list(st.iterations)[:5], np.int32)
                for k, v in trace_arrays(tr).items():
                    rob[key + "_" + k] = v
    np.savez_compressed(os.path.join(HERE, "golden_robust.npz"), **rob)
    print("golden_robust.npz", os.path.getsize(os.path.join(HERE, "golden_robust.npz")) // 1024,
          "KiB")


if __name__ == "__main__":
    if "--robust-only" not in sys.argv:
        main()
    main_robust()
