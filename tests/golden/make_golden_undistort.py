"""Fixture of the undistortion front-end (tests/golden/golden_undistort.npz).

Made from the CPU oracle AFTER checking each stage against the real OpenCV (cv2) on the same
inputs, so the stored hashes are cv2's own outputs.  Run: python tests/golden/make_golden_undistort.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2  # noqa: E402

from oracle import uw_oracle as O  # noqa: E402
from test_undistort import CASES, K_of, sha, make_image  # noqa: E402

out = {}
for name, c in CASES.items():
    K, d = K_of(c), np.array(c["dist"], np.float32)
    nK = O.optimal_new_camera_matrix(K, d, c["in_size"], 1.0, c["out_size"])
    ref_K, _ = cv2.getOptimalNewCameraMatrix(K, d.reshape(4, 1), c["in_size"], 1.0, c["out_size"],
                                             False)
    assert np.array_equal(nK, ref_K)
    m1, m2 = O.init_undistort_rectify_map(K, d, nK, c["out_size"])
    r1, r2 = cv2.initUndistortRectifyMap(K, d.reshape(4, 1), None, ref_K, c["out_size"],
                                         cv2.CV_16SC2)
    assert np.array_equal(m1, r1) and np.array_equal(m2, r2)
    img = make_image(c)
    und = O.remap_bilinear(img, m1, m2)
    assert np.array_equal(und, cv2.remap(img, r1, r2, cv2.INTER_LINEAR))
    out[name + "_newK"] = nK
    out[name + "_sha"] = np.array([sha(m1), sha(m2), sha(und)])
    out[name + "_roi"] = np.array(O.calculate_roi(und), np.int32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_undistort.npz"), **out)
print("golden_undistort.npz written:", {k: v.tolist() for k, v in out.items() if k.endswith("roi")})
