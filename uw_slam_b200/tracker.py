"""Host-side mirror of the reference's Tracker / CameraModel / Frame surface, on top of the
C ABI (include/uwtrack.h).  Method names follow /root/reference/include/Tracker.h:
InitializePyramid, ApplyGradient, ObtainCandidatePoints, EstimatePose, WarpFunction.

The C++ facade (include/uw/uw_tracker.hpp) is what a C++ `System` links against; this file
is the same surface for Python callers (tests, bench).
"""
import ctypes as C
import xml.etree.ElementTree as ET

import numpy as np

from . import _lib as L


class UwtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("uwtrack error %d: %s" % (code, msg))
        self.code = code


class CameraModel:
    """uw::CameraModel (include/CameraModel.h:42-145) for rectified pinhole input.

    Reads the reference's calibration XML (calibration/*.xml): in/out size, fx fy cx cy and
    the four distortion coefficients; normalised intrinsics (cx, cy < 1) are rescaled by the
    input size as in src/CameraModel.cpp:61-68.  With a non-zero first distortion coefficient
    the rectify branch (CameraModel.cpp:84-98) runs: new camera matrix for alpha = 1 and the
    fixed-point undistortion maps, both computed by libuwtrack's host functions.
    """

    def __init__(self):
        self.in_width = self.in_height = self.out_width = self.out_height = 0
        self.calib = np.zeros(4, np.float32)
        self.dist = np.zeros(4, np.float32)
        self.valid = False
        self.output_K = None
        self.map1 = self.map2 = None

    def GetCameraModel(self, path):
        root = ET.parse(path).getroot()

        def num(tag):
            return int(root.find(tag).text.split()[0])

        def mat(tag):
            return np.array(root.find(tag).find("data").text.split(), np.float32)

        self.in_width, self.in_height = num("in_width"), num("in_height")
        self.out_width, self.out_height = num("out_width"), num("out_height")
        self.calib = mat("calibration_values")
        self.dist = mat("rectification")
        if self.calib[2] < 1 and self.calib[3] < 1:  # CameraModel.cpp:61-68
            self.calib = self.calib * np.array([self.in_width, self.in_height,
                                                self.in_width, self.in_height], np.float32)
        self.valid = bool(self.dist[0] != 0)  # CameraModel.cpp:78-83
        if self.valid:
            self._rectify()
        return self

    def _original_K(self):
        K = np.zeros((3, 3), np.float32)
        K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[2, 2] = (self.calib[0], self.calib[1],
                                                       self.calib[2], self.calib[3], 1)
        return K

    def _rectify(self):
        """CameraModel.cpp:84-98: getOptimalNewCameraMatrix(alpha = 1) + initUndistortRectifyMap
        (CV_16SC2 fixed-point maps)."""
        lib = L.load()
        K = np.ascontiguousarray(self._original_K())
        d = np.ascontiguousarray(self.dist, np.float32)
        nK = np.empty((3, 3), np.float32)
        rc = lib.uwt_camera_optimal_matrix(K.ctypes.data_as(L._fp), d.ctypes.data_as(L._fp),
                                           self.in_width, self.in_height, 1.0, self.out_width,
                                           self.out_height, nK.ctypes.data_as(L._fp))
        if rc != 0:
            raise UwtError(rc, "uwt_camera_optimal_matrix")
        self.map1 = np.empty((self.out_height, self.out_width, 2), np.int16)
        self.map2 = np.empty((self.out_height, self.out_width), np.uint16)
        rc = lib.uwt_camera_undistort_maps(K.ctypes.data_as(L._fp), d.ctypes.data_as(L._fp),
                                           nK.ctypes.data_as(L._fp), self.out_width,
                                           self.out_height, self.map1.ctypes.data_as(L._i16p),
                                           self.map2.ctypes.data_as(C.POINTER(C.c_uint16)))
        if rc != 0:
            raise UwtError(rc, "uwt_camera_undistort_maps")
        self.output_K = nK

    @classmethod
    def from_distorted(cls, in_size, out_size, fx, fy, cx, cy, dist):
        """The rectify branch without an XML file (tests, benches)."""
        m = cls()
        m.in_width, m.in_height = in_size
        m.out_width, m.out_height = out_size
        m.calib = np.array([fx, fy, cx, cy], np.float32)
        m.dist = np.asarray(dist, np.float32)
        m.valid = bool(m.dist[0] != 0)
        if m.valid:
            m._rectify()
        return m

    def Undistort(self, image, device=0):
        """CameraModel::Undistort (CameraModel.cpp:101-103): remap one host image on the GPU."""
        img = np.ascontiguousarray(image, np.uint8)
        out = np.empty((self.out_height, self.out_width), np.uint8)
        rc = L.load().uwt_undistort_image(device, img.ctypes.data_as(L._u8p), img.shape[1],
                                          img.shape[0], img.shape[1],
                                          self.map1.ctypes.data_as(L._i16p),
                                          self.map2.ctypes.data_as(C.POINTER(C.c_uint16)),
                                          self.out_width, self.out_height,
                                          out.ctypes.data_as(L._u8p))
        if rc != 0:
            raise UwtError(rc, "uwt_undistort_image")
        return out

    def GetMap1(self):
        return self.map1

    def GetMap2(self):
        return self.map2

    @classmethod
    def from_intrinsics(cls, width, height, fx, fy, cx, cy):
        m = cls()
        m.in_width = m.out_width = width
        m.in_height = m.out_height = height
        m.calib = np.array([fx, fy, cx, cy], np.float32)
        return m

    def GetK(self):
        # output_intrinsic_camera_: the new matrix when rectifying, else the original
        # (CameraModel.cpp:78-98)
        return self.output_K if self.valid else self._original_K()

    def GetOutputWidth(self):
        return self.out_width

    def GetOutputHeight(self):
        return self.out_height

    def GetInputWidth(self):
        return self.in_width

    def GetInputHeight(self):
        return self.in_height

    def IsValid(self):
        return self.valid


def CalculateROI(undistorted):
    """System::CalculateROI (System.cpp:148-191) on the first undistorted image.
    Returns (x, y, w, h): the crop corner and the new w_, h_."""
    u = np.ascontiguousarray(undistorted, np.uint8)
    roi = (C.c_int * 4)()
    rc = L.load().uwt_calculate_roi(u.ctypes.data_as(L._u8p), u.shape[1], u.shape[0], u.shape[1],
                                    roi)
    if rc != 0:
        raise UwtError(rc, "uwt_calculate_roi: the undistorted image is black on a mid line")
    return tuple(roi)


def AlignROI(roi, levels=5):
    """The reference lets w_, h_ be whatever CalculateROI finds, which breaks its own pyramid
    sizing; the tracker needs the width divisible by 16 and both sides by
    2^(levels-1).  Shrinks the ROI to the largest compliant size, keeping the corner."""
    x, y, w, h = roi
    div = max(16, 1 << (levels - 1))
    return x, y, w // div * div, h // (1 << (levels - 1)) * (1 << (levels - 1))


class Frame:
    """uw::Frame (include/System.h:63-103): a handle on one device-side frame slot."""

    def __init__(self, tracker, slot):
        self._t = tracker
        self.slot = slot
        self.obtained_gradients_ = False
        self.obtained_candidatePoints_ = False
        self.rigid_transformation_ = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)

    def image(self, lvl):
        return self._t.get_image(self.slot, lvl)

    def gradients(self, lvl):
        return self._t.get_gradients(self.slot, lvl)

    def candidatePoints(self, lvl):
        return self._t.get_candidates(self.slot, lvl)

    def depth(self, lvl):
        return self._t.get_depth(self.slot, lvl)


class Tracker:
    """uw::Tracker (include/Tracker.h:90-531), direct photometric path only."""

    def __init__(self, depth_available=False, **cfg):
        # Tracker::Tracker(bool _depth_available), Tracker.cpp:273-277: with depth the candidate
        # points need depth != 0 and carry Z = depth * 0.0002 (as shipped: the at<uchar> read)
        self._lib = L.load()
        self._h = None
        self._cfg_overrides = dict(cfg)
        if depth_available:
            self._cfg_overrides.setdefault("depth_mode", L.DEPTH_REFERENCE)
        self.cfg = None
        self._keep = []

    # -- Tracker::InitializePyramid(int w, int h, Mat K), Tracker.cpp:297 -------------------
    def InitializePyramid(self, width, height, K, **cfg):
        c = L.Config()
        self._lib.uwt_default_config(C.byref(c))
        c.width, c.height = width, height
        c.fx, c.fy, c.cx, c.cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
        for k, v in {**self._cfg_overrides, **cfg}.items():
            if not hasattr(c, k):
                raise AttributeError(k)
            setattr(c, k, v)
        h = L._H()
        rc = self._lib.uwt_create(C.byref(c), C.byref(h))
        if rc != 0:
            raise UwtError(rc, self._lib.uwt_last_error(None).decode())
        self._h, self.cfg = h, c
        self._src_w, self._src_h = c.width, c.height
        return self

    # -- System::AddFrame remap + ROI crop, System.cpp:232-235 ------------------------------
    def SetUndistortion(self, camera, roi_xy=(0, 0)):
        """After this, AddFrames takes DISTORTED in_width x in_height frames; remap and crop are
        fused into the pyramid kernel.  camera=None switches it off."""
        if camera is None:
            self._check(self._lib.uwt_set_undistortion(self._h, None, None, 0, 0, 0, 0, 0, 0))
            self._src_w, self._src_h = self.cfg.width, self.cfg.height
            return
        m1 = np.ascontiguousarray(camera.GetMap1(), np.int16)
        m2 = np.ascontiguousarray(camera.GetMap2(), np.uint16)
        self._check(self._lib.uwt_set_undistortion(
            self._h, m1.ctypes.data_as(L._i16p), m2.ctypes.data_as(C.POINTER(C.c_uint16)),
            camera.GetOutputWidth(), camera.GetOutputHeight(), camera.GetInputWidth(),
            camera.GetInputHeight(), int(roi_xy[0]), int(roi_xy[1])))
        self._src_w, self._src_h = camera.GetInputWidth(), camera.GetInputHeight()

    def close(self):
        if self._h is not None:
            self._lib.uwt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise UwtError(rc, self._lib.uwt_last_error(self._h).decode())

    @staticmethod
    def _slots(s):
        a = np.ascontiguousarray(np.atleast_1d(s), np.int32)
        return a, a.ctypes.data_as(L._ip), int(a.size)

    def level_info(self, lvl):
        info = L.LevelInfo()
        self._check(self._lib.uwt_get_level_info(self._h, lvl, C.byref(info)))
        return info

    # -- System::AddFrame pyramid loop, System.cpp:246-251 ----------------------------------
    def AddFrames(self, slots, frames):
        """frames: u8 array [n, H, W] (or [H, W]) in host memory."""
        a, p, n = self._slots(slots)
        f = np.ascontiguousarray(frames, np.uint8).reshape(n, self._src_h, self._src_w)
        self._check(self._lib.uwt_upload_frames(self._h, n, p, f.ctypes.data, self._src_w,
                                                self._src_w * self._src_h))
        return [Frame(self, int(s)) for s in a]

    def AddDepthFrames(self, slots, depth):
        """System::AddFrame depth part (System.cpp:241-250): u16 array [n, H, W] (or [H, W])."""
        a, p, n = self._slots(slots)
        d = np.ascontiguousarray(depth, np.uint16).reshape(n, self.cfg.height, self.cfg.width)
        self._check(self._lib.uwt_upload_depth_frames(self._h, n, p, d.ctypes.data,
                                                      self.cfg.width * 2,
                                                      self.cfg.width * self.cfg.height * 2))

    def get_depth(self, slot, lvl):
        i = self.level_info(lvl)
        out = np.empty((i.height, i.width), np.uint16)
        self._check(self._lib.uwt_get_depth(self._h, slot, lvl,
                                            out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    def AddFramesHostPtr(self, slots, ptr, row_stride, frame_stride):
        a, p, n = self._slots(slots)
        self._check(self._lib.uwt_upload_frames(self._h, n, p, ptr, row_stride, frame_stride))

    def AddFramesDevice(self, slots, dev_ptr, row_stride=None, frame_stride=None):
        a, p, n = self._slots(slots)
        rs = row_stride or self._src_w
        fs = frame_stride or rs * self._src_h
        self._check(self._lib.uwt_set_frames_device(self._h, n, p, dev_ptr, rs, fs))

    # -- Tracker::ApplyGradient(Frame*), Tracker.cpp:1127 ------------------------------------
    def ApplyGradient(self, frames):
        a, p, n = self._slots(self._as_slots(frames))
        self._check(self._lib.uwt_apply_gradient(self._h, n, p))
        for f in self._as_frames(frames):
            f.obtained_gradients_ = True

    # -- Tracker::ObtainCandidatePoints(Frame*), Tracker.cpp:1314 ----------------------------
    def ObtainCandidatePoints(self, frames):
        a, p, n = self._slots(self._as_slots(frames))
        self._check(self._lib.uwt_select_candidates(self._h, n, p))
        for f in self._as_frames(frames):
            f.obtained_candidatePoints_ = True

    # -- Tracker::ObtainAllPoints(Frame*), Tracker.cpp:1259: every pixel with depth > 0 ------
    def ObtainAllPoints(self, frames):
        if self.cfg is None or self.cfg.depth_mode != L.DEPTH_ALL_POINTS:
            raise UwtError(-1, "ObtainAllPoints needs a tracker created with "
                               "depth_mode=DEPTH_ALL_POINTS")
        self.ObtainCandidatePoints(frames)

    # -- Tracker::EstimatePose(Frame* prev, Frame* cur), Tracker.cpp:362 ---------------------
    def EstimatePose(self, prev, cur, init_poses=None, return_stats=False):
        pa, pp, n = self._slots(self._as_slots(prev))
        ca, cp, n2 = self._slots(self._as_slots(cur))
        assert n == n2
        out = np.empty((n, 7), np.float32)
        stats = (L.TrackStats * n)()
        ip = None
        if init_poses is not None:
            ipa = np.ascontiguousarray(init_poses, np.float32).reshape(n, 7)
            ip = ipa.ctypes.data_as(L._fp)
        self._check(self._lib.uwt_estimate_pose(self._h, n, pp, cp, ip,
                                                out.ctypes.data_as(L._fp), stats))
        for f, pose in zip(self._as_frames(prev), out):
            f.rigid_transformation_ = pose.copy()  # Tracker.cpp:595
        return (out, list(stats)) if return_stats else out

    def EstimatePoseAsync(self, prev, cur):
        pa, pp, n = self._slots(self._as_slots(prev))
        ca, cp, _ = self._slots(self._as_slots(cur))
        self._check(self._lib.uwt_estimate_pose_async(self._h, n, pp, cp, None))

    def FetchPoses(self, n):
        out = np.empty((n, 7), np.float32)
        self._check(self._lib.uwt_fetch_poses(self._h, n, out.ctypes.data_as(L._fp), None))
        return out

    # -- Tracker::WarpFunction(Mat, SE3, int), Tracker.cpp:1417 ------------------------------
    def WarpFunction(self, points, pose7, lvl):
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 4)
        pose = np.ascontiguousarray(pose7, np.float32)
        out = np.empty_like(pts)
        self._check(self._lib.uwt_warp_points(self._h, pts.ctypes.data_as(L._fp), pts.shape[0],
                                              pose.ctypes.data_as(L._fp), lvl,
                                              out.ctypes.data_as(L._fp)))
        return out

    # -- sharded single-frame mode (uwt_shard_*) -------------------------------------------
    def ShardBegin(self, prev_slot, cur_slot, rank, nranks, init_pose=None):
        ip = None
        if init_pose is not None:
            self._ip = np.ascontiguousarray(init_pose, np.float32)
            ip = self._ip.ctypes.data_as(L._fp)
        self._check(self._lib.uwt_shard_begin(self._h, prev_slot, cur_slot, rank, nranks, ip))

    def ShardAccumulate(self, dev_sums_ptr):
        self._check(self._lib.uwt_shard_accumulate(self._h, dev_sums_ptr))

    def ShardUpdate(self, dev_sums_ptr):
        done = C.c_int(0)
        self._check(self._lib.uwt_shard_update(self._h, dev_sums_ptr, C.byref(done)))
        return bool(done.value)

    def ShardResult(self):
        out = np.empty(7, np.float32)
        st = L.TrackStats()
        self._check(self._lib.uwt_shard_result(self._h, out.ctypes.data_as(L._fp), C.byref(st)))
        return out, st

    # -- fused (in-kernel all-reduce over peer memory) sharded mode -------------------------
    def ShardIpcExport(self):
        n = self._lib.uwt_shard_ipc_handle_size()
        buf = (C.c_uint8 * n)()
        self._check(self._lib.uwt_shard_ipc_export(self._h, buf))
        return bytes(buf)

    def ShardIpcConnect(self, rank, nranks, handles_bytes):
        buf = (C.c_uint8 * len(handles_bytes)).from_buffer_copy(handles_bytes)
        self._check(self._lib.uwt_shard_ipc_connect(self._h, rank, nranks, buf))

    def ShardConnectLocal(self, rank, peers):
        arr = (L._H * len(peers))(*[p._h for p in peers])
        self._check(self._lib.uwt_shard_connect_local(self._h, rank, len(peers), arr))

    def ShardEstimateFusedAsync(self, prev_slot, cur_slot, init_pose=None, grid=0):
        ip = None
        if init_pose is not None:
            self._ip = np.ascontiguousarray(init_pose, np.float32)
            ip = self._ip.ctypes.data_as(L._fp)
        self._check(self._lib.uwt_shard_estimate_fused_async(self._h, prev_slot, cur_slot, ip,
                                                             grid))

    def ShardEstimateFusedWait(self):
        out = np.empty(7, np.float32)
        st = L.TrackStats()
        self._check(self._lib.uwt_shard_estimate_fused_wait(self._h, out.ctypes.data_as(L._fp),
                                                            C.byref(st)))
        return out, st

    def synchronize(self):
        self._check(self._lib.uwt_synchronize(self._h))

    def launch_count(self, include_aux=False):
        """Compute kernels launched so far; with include_aux also the argument-staging kernels."""
        n = int(self._lib.uwt_launch_count(self._h))
        if include_aux:
            n += int(self._lib.uwt_aux_launch_count(self._h))
        return n

    def profile(self, on=True):
        self._check(self._lib.uwt_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{class: (milliseconds, launches)} accumulated since profile(True)."""
        ms = (C.c_double * len(L.KERNEL_CLASSES))()
        ln = (C.c_longlong * len(L.KERNEL_CLASSES))()
        self._check(self._lib.uwt_profile_read(self._h, ms, ln))
        return {k: (ms[i], int(ln[i])) for i, k in enumerate(L.KERNEL_CLASSES)}

    def stream_ptr(self):
        return int(self._lib.uwt_stream(self._h) or 0)

    # -- read-back accessors ------------------------------------------------------------------
    def get_image(self, slot, lvl):
        i = self.level_info(lvl)
        out = np.empty((i.height, i.width), np.uint8)
        self._check(self._lib.uwt_get_image(self._h, slot, lvl, out.ctypes.data_as(L._u8p)))
        return out

    def get_gradients(self, slot, lvl):
        i = self.level_info(lvl)
        gx = np.empty((i.height, i.width), np.int16)
        gy = np.empty((i.height, i.width), np.int16)
        g = np.empty((i.height, i.width), np.uint8)
        self._check(self._lib.uwt_get_gradients(self._h, slot, lvl, gx.ctypes.data_as(L._i16p),
                                                gy.ctypes.data_as(L._i16p),
                                                g.ctypes.data_as(L._u8p)))
        return gx, gy, g

    def get_gradient_image(self, slot, lvl):
        """gradient_[lvl] exactly as it sits in device memory (no int16 planes requested, so
        nothing is recomputed for the read-back)."""
        i = self.level_info(lvl)
        g = np.empty((i.height, i.width), np.uint8)
        self._check(self._lib.uwt_get_gradients(self._h, slot, lvl, None, None,
                                                g.ctypes.data_as(L._u8p)))
        return g

    def get_candidates(self, slot, lvl):
        n = C.c_int(0)
        self._check(self._lib.uwt_get_candidate_count(self._h, slot, lvl, C.byref(n)))
        pts = np.empty((max(n.value, 1), 4), np.float32)
        self._check(self._lib.uwt_get_candidates(self._h, slot, lvl, pts.ctypes.data_as(L._fp),
                                                 pts.shape[0], C.byref(n)))
        return pts[:n.value]

    def get_records(self, slot, lvl):
        """Decoded packed records of an optimised level: dict of x, y, i1, gx, gy arrays."""
        n = C.c_int(0)
        self._check(self._lib.uwt_get_records(self._h, slot, lvl, None, 0, C.byref(n)))
        buf = np.zeros(max(n.value, 1), np.uint64)
        self._check(self._lib.uwt_get_records(self._h, slot, lvl,
                                              buf.ctypes.data_as(C.POINTER(C.c_uint64)),
                                              buf.size, C.byref(n)))
        r = buf[:n.value]

        def s13(v):
            v = v.astype(np.int64) & 0x1FFF
            return np.where(v >= 4096, v - 8192, v).astype(np.int32)
        return {"x": (r & 0xFFF).astype(np.int32), "y": ((r >> 12) & 0xFFF).astype(np.int32),
                "i1": ((r >> 24) & 0xFF).astype(np.int32), "gx": s13(r >> 32),
                "gy": s13(r >> 45)}

    def get_trace(self, index=0, cap=512):
        buf = (L.IterTrace * cap)()
        n = C.c_int(0)
        self._check(self._lib.uwt_get_trace(self._h, index, buf, cap, C.byref(n)))
        return [buf[i] for i in range(n.value)]

    @staticmethod
    def _as_slots(frames):
        if isinstance(frames, Frame):
            return [frames.slot]
        if isinstance(frames, (list, tuple)) and frames and isinstance(frames[0], Frame):
            return [f.slot for f in frames]
        return frames

    @staticmethod
    def _as_frames(frames):
        if isinstance(frames, Frame):
            return [frames]
        if isinstance(frames, (list, tuple)) and frames and isinstance(frames[0], Frame):
            return list(frames)
        return []
