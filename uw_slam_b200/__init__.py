"""uw_slam_b200 -- B200-native direct photometric tracker (hot path of uw-slam's Tracker).

The compute lives in libuwtrack.so (hand-written CUDA for sm_100a behind the C ABI of
include/uwtrack.h); this package is the thin host-side mirror of the reference's
Tracker / CameraModel / Frame surface plus the synthetic-input generator.
"""
from .tracker import (AlignROI, CalculateROI, CameraModel, Frame, Tracker,  # noqa: F401
                      UwtError)
from . import synth  # noqa: F401
