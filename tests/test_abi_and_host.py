"""CPU: the C-ABI library loads and exports every symbol include/uwtrack.h declares, the
header is valid C and agrees with the ctypes mirror, the C++ facade compiles and links, and
the host-side logic (calibration parsing, error reporting) behaves like the reference's."""
import ctypes as C
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "uwtrack.h")


@pytest.fixture(scope="module")
def built():
    from uw_slam_b200 import build
    return build.build()


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uwt_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from uw_slam_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libuwtrack.so does not export " + n
    # and the python binding mirrors exactly the header's function set
    assert sorted(_lib.SIGNATURES) == names
    out = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = set(re.findall(r"\bT (uwt_[a-z_0-9]+)", out))
    assert set(names) <= exported


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path, built):
    from uw_slam_b200 import _lib
    src = tmp_path / "abi.c"
    src.write_text(textwrap.dedent("""
        #include <stdio.h>
        #include <stddef.h>
        #include "uwtrack.h"
        int main(void) {
          printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(uwt_config), sizeof(uwt_level_info),
                 sizeof(uwt_track_stats), sizeof(uwt_iter_trace),
                 offsetof(uwt_config, gradient_threshold), offsetof(uwt_iter_trace, A));
          return 0;
        }"""))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic",
                           "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert got == [C.sizeof(_lib.Config), C.sizeof(_lib.LevelInfo), C.sizeof(_lib.TrackStats),
                   C.sizeof(_lib.IterTrace), _lib.Config.gradient_threshold.offset,
                   _lib.IterTrace.A.offset]


def test_default_config_is_the_reference(built):
    from uw_slam_b200 import _lib
    c = _lib.Config()
    assert _lib.load().uwt_default_config(C.byref(c)) == 0
    # Options.cpp:26-27, Tracker.cpp:364-369,559, calibrationTUM.xml:20
    assert (c.levels, c.first_level, c.last_level, c.max_iterations) == (5, 4, 1, 50)
    assert (c.width, c.height, c.fx, c.cx, c.cy) == (640, 480, 525.0, 319.5, 239.5)
    assert abs(c.epsilon - 0.001) < 1e-9 and c.residual_scale == 50.0
    assert c.gradient_threshold == 20.0 and c.solve_mode == _lib.SOLVE_LU


def test_cpp_facade_compiles_and_links(tmp_path, built):
    exe = tmp_path / "track_sequence"
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I",
                           os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "track_sequence.cpp"),
                           "-L", os.path.dirname(built), "-luwtrack",
                           "-Wl,-rpath," + os.path.dirname(built), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def test_no_gpu_fails_loudly_not_silently(built):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import uw_slam_b200 as U
    t = U.Tracker(False)
    with pytest.raises(U.UwtError) as e:
        t.InitializePyramid(640, 480, np.array([[525, 0, 319.5], [0, 525, 239.5], [0, 0, 1]],
                                               np.float32))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "uw_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "uw_oracle" not in txt and "oracle/" not in txt, f
    for f in ("uwtrack.h", os.path.join("uw", "uw_tracker.hpp")):
        assert "oracle" not in open(os.path.join(ROOT, "include", f)).read()


CALIB_XML = """<?xml version="1.0"?>
<opencv_storage>
<in_width type_id="integer"> {iw} </in_width>
<in_height type_id="integer"> {ih} </in_height>
<out_width type_id="integer"> {ow} </out_width>
<out_height type_id="integer"> {oh} </out_height>
<calibration_values type_id="opencv-matrix">
  <rows>1</rows><cols>4</cols><dt>f</dt>
  <data> {c} </data></calibration_values>
<rectification type_id="opencv-matrix">
  <rows>1</rows><cols>4</cols><dt>f</dt>
  <data> {d} </data></rectification>
</opencv_storage>
"""


def test_camera_model_reads_the_reference_xml_schema(tmp_path):
    import uw_slam_b200 as U
    # the TUM-RGBD file of the reference: pixel intrinsics, no distortion -> not "valid"
    p = tmp_path / "tum.xml"
    p.write_text(CALIB_XML.format(iw=640, ih=480, ow=640, oh=480, c="525 525 319.5 239.5",
                                  d="0 0 0 1"))
    m = U.CameraModel().GetCameraModel(str(p))
    K = m.GetK()
    assert (K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[2, 2]) == (525, 525, 319.5, 239.5, 1)
    assert (m.GetOutputWidth(), m.GetOutputHeight(), m.GetInputWidth()) == (640, 480, 640)
    assert not m.IsValid()
    # TUM-mono convention: normalised intrinsics are rescaled (CameraModel.cpp:61-68)
    p2 = tmp_path / "mono.xml"
    p2.write_text(CALIB_XML.format(iw=1280, ih=1024, ow=1280, oh=1024,
                                   c="0.5357 0.6696 0.4932 0.5004", d="-0.2 0.04 0 0"))
    m2 = U.CameraModel().GetCameraModel(str(p2))
    K2 = m2._original_K()
    assert K2[0, 0] == np.float32(0.5357) * np.float32(1280)
    assert K2[1, 2] == np.float32(0.5004) * np.float32(1024)
    # non-zero distortion: the rectify branch runs (CameraModel.cpp:84-98) and GetK() is the
    # new camera matrix; maps have the output size
    assert m2.IsValid()
    assert m2.GetK()[0, 0] != K2[0, 0] and m2.GetK()[2, 2] == 1
    assert m2.GetMap1().shape == (1024, 1280, 2) and m2.GetMap2().shape == (1024, 1280)


def test_synth_is_seeded_and_consistent():
    from uw_slam_b200 import synth
    a = synth.render_pair("tiny", 7)
    b = synth.render_pair("tiny", 7)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not np.array_equal(a[0], synth.render_pair("tiny", 8)[0])
    # zero motion renders the same frame twice
    cal = synth.CALIB["small"]
    f0 = synth.render_frame(cal, 1)
    f1 = synth.render_frame(cal, 1, synth.homography_inv(cal, np.zeros(3), np.zeros(3)))
    assert np.array_equal(f0, f1)
    assert 20 < f0.std() < 90
