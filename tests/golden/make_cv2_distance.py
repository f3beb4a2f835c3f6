"""Generates tests/golden/cv2_distance.json: the distance between the oracle (canonical rule U3:
fp64 sums rounded once) and Tracker::EstimatePose run through the real OpenCV calls
(oracle/cv2_transliteration.py) on the golden pairs.  Run from the repo root:
    python tests/golden/make_cv2_distance.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = [("small", s) for s in range(6)] + [("tum", s) for s in range(4)] + [("euroc", 0),
                                                                             ("tum_mono", 0)]


def measure(cv2, O, T, synth, calib, seed, folded=False):
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    prev, cur, _, _ = synth.render_pair(calib, seed)
    fp, fc = O.FrameData(prev), O.FrameData(cur, with_candidates=False)
    p = O.default_params(w, h, fx, fy, cx, cy)
    opose, ost, otr = O.estimate_pose(p, fp, fc)
    cpose, citers, ctr = T.estimate_pose(cv2, O, p, fp, fc, folded=folded)
    ang, rel, dt = T.pose_distance(cpose, opose)
    # first sweep of every level has identical inputs on both sides: A, b differ only by the
    # accumulation arithmetic of cv::gemm versus rule U3
    relA, relb = 0.0, 0.0
    for (lvl, k, n, err, A, b, d), o in zip(ctr, otr):
        if A is None or o.broke or (o.level, o.k) != (lvl, k):
            continue
        oA, ob = np.array(o.A[:], np.float64).reshape(6, 6), np.array(o.b[:], np.float64)
        relA = max(relA, float(np.max(np.abs(A - oA) / np.maximum(np.abs(oA), 1e-30))))
        relb = max(relb, float(np.max(np.abs(b - ob) / np.maximum(np.abs(ob), 1e-30))))
    return {"calib": calib, "seed": seed, "folded_matexpr": folded,
            "same_sweep_structure": [(r[0], r[1]) for r in ctr] == [(o.level, o.k) for o in otr],
            "rotation_diff_rad": ang, "translation_diff_rel": rel, "translation_diff_abs": dt,
            "max_rel_diff_A": relA, "max_rel_diff_b": relb,
            "oracle_pose": [float(v) for v in opose], "cv2_pose": [float(v) for v in cpose]}


def main():
    import cv2
    from oracle import cv2_transliteration as T
    from oracle import uw_oracle as O
    from uw_slam_b200 import synth
    out = {"cv2_version": cv2.__version__,
           "what": "oracle (rule U3) vs EstimatePose through real cv2.gemm / cv2.invert calls",
           "cases": [measure(cv2, O, T, synth, c, s) for c, s in CASES] +
                    [measure(cv2, O, T, synth, "small", 0, folded=True),
                     measure(cv2, O, T, synth, "tum", 0, folded=True)]}
    with open(os.path.join(HERE, "cv2_distance.json"), "w") as f:
        json.dump(out, f, indent=1)
    for c in out["cases"]:
        print(c["calib"], c["seed"], c["folded_matexpr"], c["same_sweep_structure"],
              "rot %.3e rad" % c["rotation_diff_rad"], "trans rel %.3e" % c["translation_diff_rel"],
              "A %.2e b %.2e" % (c["max_rel_diff_A"], c["max_rel_diff_b"]))


if __name__ == "__main__":
    main()
