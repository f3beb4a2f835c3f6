"""ctypes binding of libuwtrack.so -- exactly the C ABI of include/uwtrack.h.

Fails loudly if the library has not been built: there is no CPU fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# UWT_LIBRARY: development knob to A/B a differently built libuwtrack (tools/ab_variants.sh)
SO_PATH = os.environ.get("UWT_LIBRARY") or os.path.join(HERE, "libuwtrack.so")
MAX_LEVELS = 7
OK, E_INVALID, E_CUDA, E_STATE, E_NOMEM = 0, -1, -2, -3, -4
SOLVE_LU, SOLVE_INVERSE, SOLVE_CHOLESKY_LM = 0, 1, 2
FLAG_TRACE = 1
FLAG_DMMA_ACCUM = 2
FLAG_CLUSTER_KERNEL = 4
FLAG_LAZY_LEVELS = 8
FLAG_SEPARATE_GRADIENT = 16
WEIGHT_IDENTITY, WEIGHT_TUKEY, WEIGHT_HUBER = 0, 1, 2
DEPTH_NONE, DEPTH_REFERENCE, DEPTH_U16, DEPTH_ALL_POINTS = 0, 1, 2, 3
GRADIENT_SCHARR, GRADIENT_SOBEL = 0, 1
SAMPLE_NEAREST, SAMPLE_BILINEAR = 0, 1
KERNEL_CLASSES = ("pyramid", "gradient", "candidates", "estimate")


class Config(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("levels", C.c_int), ("first_level", C.c_int), ("last_level", C.c_int),
        ("max_iterations", C.c_int), ("epsilon", C.c_float), ("residual_scale", C.c_float),
        ("gradient_threshold", C.c_double), ("solve_mode", C.c_int), ("device", C.c_int),
        ("max_frames", C.c_int), ("cluster_size", C.c_int), ("flags", C.c_uint),
        ("weight_mode", C.c_int), ("huber_delta", C.c_float), ("depth_mode", C.c_int),
        ("lm_lambda", C.c_float), ("gradient_op", C.c_int),
        ("sampling", C.c_int),
    ]


class LevelInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int)] + \
        [(n, C.c_float) for n in ("fx", "fy", "cx", "cy", "invfx", "invfy")]


class TrackStats(C.Structure):
    _fields_ = [
        ("iterations", C.c_int * MAX_LEVELS), ("evaluations", C.c_int * MAX_LEVELS),
        ("n_points", C.c_int * MAX_LEVELS), ("final_error", C.c_float * MAX_LEVELS),
    ]


class IterTrace(C.Structure):
    _fields_ = [
        ("level", C.c_int), ("k", C.c_int), ("n_valid", C.c_int), ("broke", C.c_int),
        ("sum_r2", C.c_longlong), ("error", C.c_float),
        ("A", C.c_float * 36), ("b", C.c_float * 6), ("delta", C.c_float * 6),
        ("pose", C.c_float * 7),
    ]


_H = C.c_void_p
_ip, _fp, _u8p, _i16p = (C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint8),
                         C.POINTER(C.c_int16))

# name -> (restype, argtypes); every symbol include/uwtrack.h declares
SIGNATURES = {
    "uwt_default_config": (C.c_int, [C.POINTER(Config)]),
    "uwt_create": (C.c_int, [C.POINTER(Config), C.POINTER(_H)]),
    "uwt_destroy": (C.c_int, [_H]),
    "uwt_last_error": (C.c_char_p, [_H]),
    "uwt_get_level_info": (C.c_int, [_H, C.c_int, C.POINTER(LevelInfo)]),
    "uwt_stream": (C.c_void_p, [_H]),
    "uwt_synchronize": (C.c_int, [_H]),
    "uwt_upload_frames": (C.c_int, [_H, C.c_int, _ip, C.c_void_p, C.c_size_t, C.c_size_t]),
    "uwt_set_frames_device": (C.c_int, [_H, C.c_int, _ip, C.c_void_p, C.c_size_t, C.c_size_t]),
    "uwt_camera_optimal_matrix": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_double, C.c_int,
                                            C.c_int, _fp]),
    "uwt_camera_undistort_maps": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, _i16p,
                                            C.POINTER(C.c_uint16)]),
    "uwt_undistort_image": (C.c_int, [C.c_int, _u8p, C.c_int, C.c_int, C.c_size_t, _i16p,
                                      C.POINTER(C.c_uint16), C.c_int, C.c_int, _u8p]),
    "uwt_calculate_roi": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_size_t, _ip]),
    "uwt_set_undistortion": (C.c_int, [_H, _i16p, C.POINTER(C.c_uint16), C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int]),
    "uwt_upload_depth_frames": (C.c_int, [_H, C.c_int, _ip, C.c_void_p, C.c_size_t, C.c_size_t]),
    "uwt_get_depth": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(C.c_uint16)]),
    "uwt_apply_gradient": (C.c_int, [_H, C.c_int, _ip]),
    "uwt_select_candidates": (C.c_int, [_H, C.c_int, _ip]),
    "uwt_estimate_pose": (C.c_int, [_H, C.c_int, _ip, _ip, _fp, _fp, C.POINTER(TrackStats)]),
    "uwt_estimate_pose_async": (C.c_int, [_H, C.c_int, _ip, _ip, _fp]),
    "uwt_fetch_poses": (C.c_int, [_H, C.c_int, _fp, C.POINTER(TrackStats)]),
    "uwt_shard_begin": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "uwt_shard_accumulate": (C.c_int, [_H, C.c_void_p]),
    "uwt_shard_update": (C.c_int, [_H, C.c_void_p, _ip]),
    "uwt_shard_result": (C.c_int, [_H, _fp, C.POINTER(TrackStats)]),
    "uwt_shard_ipc_handle_size": (C.c_int, []),
    "uwt_shard_ipc_export": (C.c_int, [_H, C.c_void_p]),
    "uwt_shard_ipc_connect": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p]),
    "uwt_shard_connect_local": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(_H)]),
    "uwt_shard_estimate_fused_async": (C.c_int, [_H, C.c_int, C.c_int, _fp, C.c_int]),
    "uwt_shard_estimate_fused_wait": (C.c_int, [_H, _fp, C.POINTER(TrackStats)]),
    "uwt_warp_points": (C.c_int, [_H, _fp, C.c_int, _fp, C.c_int, _fp]),
    "uwt_get_image": (C.c_int, [_H, C.c_int, C.c_int, _u8p]),
    "uwt_get_gradients": (C.c_int, [_H, C.c_int, C.c_int, _i16p, _i16p, _u8p]),
    "uwt_get_candidate_count": (C.c_int, [_H, C.c_int, C.c_int, _ip]),
    "uwt_get_candidates": (C.c_int, [_H, C.c_int, C.c_int, _fp, C.c_int, _ip]),
    "uwt_get_records": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_int, _ip]),
    "uwt_get_trace": (C.c_int, [_H, C.c_int, C.POINTER(IterTrace), C.c_int, _ip]),
    "uwt_launch_count": (C.c_longlong, [_H]),
    "uwt_aux_launch_count": (C.c_longlong, [_H]),
    "uwt_profile_enable": (C.c_int, [_H, C.c_int]),
    "uwt_profile_read": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
}

_lib = None


def load():
    """Loads libuwtrack.so; raises if it is missing (build with uw_slam_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "libuwtrack.so is not built (run `python uw_slam_b200/build.py` or "
                "__graft_entry__.build()); the tracker has no CPU fallback")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
