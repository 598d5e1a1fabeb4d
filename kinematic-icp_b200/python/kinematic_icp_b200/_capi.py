"""ctypes binding of libkicp_b200.so (include/kicp.h).  There is no CPU fallback: if the shared library is missing
the import fails loudly, and without a CUDA device every call raises KicpError."""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.normpath(os.path.join(_PKG, "..", ".."))  # kinematic-icp_b200/
LIB_PATH = os.environ.get("KICP_LIB", os.path.join(_ROOT, "lib", "libkicp_b200.so"))  # KICP_LIB: A/B builds of the same ABI
KICP_MAX_ITERATIONS = 64
KICP_UNIQUE_ID_BYTES = 128
KICP_IPC_HANDLE_BYTES = 64

KICP_OK = 0
KICP_ERR_CUDA, KICP_ERR_INVALID, KICP_ERR_UNSUPPORTED, KICP_ERR_NCCL, KICP_ERR_CAPACITY = 1, 2, 3, 4, 5
KICP_WARN_NO_CORRESPONDENCES = 16

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class RegParams(C.Structure):
    """kicp_reg_params — the public fields of kinematic_icp::KinematicRegistration (Registration.hpp:45-49)."""
    _fields_ = [("max_num_iterations", C.c_int32), ("use_adaptive_odometry_regularization", C.c_int32),
                ("convergence_criterion", C.c_double), ("fixed_regularization", C.c_double)]


class RegResult(C.Structure):
    """kicp_reg_result"""
    _fields_ = [("pose", C.c_double * 7), ("beta", C.c_double), ("last_dx_norm", C.c_double), ("iterations", C.c_int32),
                ("status", C.c_int32), ("sums", (C.c_double * 8) * KICP_MAX_ITERATIONS),
                ("dx", (C.c_double * 2) * KICP_MAX_ITERATIONS)]

    def pose_np(self):
        return np.array(self.pose[:], dtype=np.float64)

    def sums_np(self):
        return np.ctypeslib.as_array(self.sums)[: self.iterations].copy()

    def dx_np(self):
        return np.ctypeslib.as_array(self.dx)[: self.iterations].copy()


class Profile(C.Structure):
    """kicp_profile"""
    _fields_ = [("assoc_ms", C.c_double), ("assoc_launches", C.c_int64), ("idle_ms", C.c_double),
                ("idle_launches", C.c_int64), ("prep_ms", C.c_double), ("registrations", C.c_int64),
                ("assoc_iterations", C.c_int64)]


class FrameInput(C.Structure):
    """kicp_frame_input"""
    _fields_ = [("data", C.c_void_p), ("n", C.c_int64), ("dtype", C.c_int32), ("point_step", C.c_int32), ("offset_x", C.c_int32),
                ("offset_y", C.c_int32), ("offset_z", C.c_int32), ("stamps", c_dp), ("n_stamps", C.c_int64)]


class FrameParams(C.Structure):
    """kicp_frame_params"""
    _fields_ = [("max_range", C.c_double), ("min_range", C.c_double), ("deskew", C.c_int32), ("voxel_size", C.c_double),
                ("stage_clouds", C.c_int32), ("reg", RegParams)]


KICP_DTYPE_F64, KICP_DTYPE_F32 = 0, 1


class KicpError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib().kicp_last_error().decode() if status in (KICP_ERR_CUDA, KICP_ERR_NCCL, KICP_ERR_CAPACITY,
                                                              KICP_ERR_UNSUPPORTED, KICP_ERR_INVALID) else ""
        super().__init__("%s: %s%s" % (where, lib().kicp_status_string(status).decode(), (" — " + msg) if msg else ""))


_LIB = None

# every symbol include/kicp.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("kicp_status_string", C.c_char_p, [C.c_int]),
    ("kicp_last_error", C.c_char_p, []),
    ("kicp_ctx_create", C.c_int, [C.c_int, C.POINTER(_P)]),
    ("kicp_ctx_destroy", C.c_int, [_P]),
    ("kicp_ctx_synchronize", C.c_int, [_P]),
    ("kicp_ctx_stream", _P, [_P]),
    ("kicp_ctx_launch_count", C.c_int64, [_P]),
    ("kicp_ctx_set_option", C.c_int, [_P, C.c_char_p, C.c_int32]),
    ("kicp_ctx_profile_begin", C.c_int, [_P]),
    ("kicp_ctx_profile_end", C.c_int, [_P, C.POINTER(Profile)]),
    ("kicp_host_alloc", C.c_int, [C.c_uint64, C.POINTER(_P)]),
    ("kicp_host_free", C.c_int, [_P]),
    ("kicp_map_create", C.c_int, [_P, C.c_double, C.c_double, C.c_uint32, C.POINTER(_P)]),
    ("kicp_map_destroy", C.c_int, [_P]),
    ("kicp_map_reserve", C.c_int, [_P, C.c_int64]),
    ("kicp_map_clear", C.c_int, [_P]),
    ("kicp_map_empty", C.c_int, [_P, c_ip]),
    ("kicp_map_num_points", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("kicp_map_num_voxels", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("kicp_map_add_points", C.c_int, [_P, c_dp, C.c_int64]),
    ("kicp_map_remove_far", C.c_int, [_P, c_dp]),
    ("kicp_map_update", C.c_int, [_P, c_dp, C.c_int64, c_dp]),
    ("kicp_map_update_pose", C.c_int, [_P, c_dp, C.c_int64, c_dp]),
    ("kicp_map_pointcloud", C.c_int, [_P, c_dp, C.c_int64, C.POINTER(C.c_int64)]),
    ("kicp_map_export_voxels", C.c_int, [_P, c_ip, c_ip, c_dp, C.c_int64, C.c_int64, C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int64)]),
    ("kicp_map_load_voxels", C.c_int, [_P, c_ip, c_ip, c_dp, C.c_int64]),
    ("kicp_map_nearest", C.c_int, [_P, c_dp, C.c_int64, c_dp, c_dp]),
    ("kicp_register", C.c_int, [_P, c_dp, C.c_int64, c_dp, c_dp, C.c_double, C.POINTER(RegParams), c_dp,
                                C.POINTER(RegResult)]),
    ("kicp_scan_create", C.c_int, [_P, C.c_int64, C.POINTER(_P)]),
    ("kicp_scan_destroy", C.c_int, [_P]),
    ("kicp_scan_upload", C.c_int, [_P, c_dp, C.c_int64]),
    ("kicp_scan_upload_async", C.c_int, [_P, c_dp, C.c_int64]),
    ("kicp_scan_upload_points", C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("kicp_scan_upload_points_async", C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("kicp_register_points", C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_dp, c_dp,
                                       C.c_double, C.POINTER(RegParams), c_dp, C.POINTER(RegResult)]),
    ("kicp_register_scan_async", C.c_int, [_P, _P, c_dp, c_dp, C.c_double, C.POINTER(RegParams), C.POINTER(RegResult)]),
    ("kicp_voxel_downsample", C.c_int, [_P, c_dp, C.c_int64, C.c_double, c_dp, C.c_int64, C.POINTER(C.c_int64)]),
    ("kicp_preprocess", C.c_int, [_P, c_dp, C.c_int64, c_dp, C.c_int64, c_dp, c_dp, C.c_double, C.c_double, C.c_int32, c_dp,
                                  C.c_int64, C.POINTER(C.c_int64)]),
    ("kicp_register_frame", C.c_int, [_P, C.POINTER(FrameInput), c_dp, c_dp, c_dp, c_dp, C.c_double, C.POINTER(FrameParams), c_dp,
                                      c_dp, C.c_int64, C.POINTER(C.c_int64), c_dp, C.c_int64, C.POINTER(C.c_int64),
                                      C.POINTER(RegResult)]),
    ("kicp_frame_clouds", C.c_int, [_P, C.POINTER(c_dp), C.POINTER(C.c_int64), C.POINTER(c_dp), C.POINTER(C.c_int64)]),
    ("kicp_comm_unique_id", C.c_int, [C.POINTER(C.c_uint8)]),
    ("kicp_comm_init", C.c_int, [_P, C.POINTER(C.c_uint8), C.c_int32, C.c_int32]),
    ("kicp_comm_destroy", C.c_int, [_P]),
    ("kicp_comm_p2p_handle", C.c_int, [_P, C.POINTER(C.c_uint8)]),
    ("kicp_comm_p2p_init", C.c_int, [_P, C.POINTER(C.c_uint8), C.c_int32, C.c_int32]),
    ("kicp_register_sharded", C.c_int, [_P, c_dp, C.c_int64, c_dp, c_dp, C.c_double, C.POINTER(RegParams), c_dp,
                                        C.POINTER(RegResult)]),
    ("kicp_register_points_sharded", C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_dp, c_dp,
                                               C.c_double, C.POINTER(RegParams), c_dp, C.POINTER(RegResult)]),
    ("kicp_register_scan_sharded_async", C.c_int, [_P, _P, c_dp, c_dp, C.c_double, C.POINTER(RegParams),
                                                   C.POINTER(RegResult)]),
]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libkicp_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C kinematic-icp_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB


def check(status, where, allow=(KICP_OK,)):
    if status not in allow:
        raise KicpError(status, where)
    return status


def dp(a):
    return a.ctypes.data_as(c_dp)


def as_points(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("expected an (N, 3) array of points")
    return a


def as_cloud(a):
    """(array, dtype code) for an (N, 3) float32 or float64 cloud: float32 stays float32 (half the PCIe bytes)."""
    a = np.asarray(a)
    if a.dtype == np.float32:
        a = np.ascontiguousarray(a)
        code = KICP_DTYPE_F32
    else:
        a = np.ascontiguousarray(a, dtype=np.float64)
        code = KICP_DTYPE_F64
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("expected an (N, 3) array of points")
    return a, code


def as_pose(p):
    return np.ascontiguousarray(p, dtype=np.float64).reshape(7)
