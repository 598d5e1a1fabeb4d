"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes binding of the CPU oracle (oracle/libkicp_oracle.so) and, when built, of the reference's own
Registration.cpp compiled against header shims (oracle/_ref/libkicp_ref.so).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class Stats(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32),
        ("associations", C.c_int32),
        ("beta", C.c_double),
        ("last_dx_norm", C.c_double),
        ("sums", (C.c_double * 7) * 64),
        ("dx", (C.c_double * 2) * 64),
    ]

    def sums_np(self):
        return np.ctypeslib.as_array(self.sums)[: self.iterations].copy()

    def dx_np(self):
        return np.ctypeslib.as_array(self.dx)[: self.iterations].copy()


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists).  Building the checker is not using it."""
    so = os.path.join(_HERE, "libkicp_oracle.so")
    if force or not os.path.exists(so) or os.path.exists("/root/reference/cpp/kinematic_icp"):
        subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)
    return so


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libkicp_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.kor_map_create.restype = C.c_void_p
        L.kor_map_create.argtypes = [C.c_double, C.c_double, C.c_uint]
        for name in ("kor_map_destroy", "kor_map_clear"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.kor_map_empty.argtypes = [C.c_void_p]
        L.kor_map_num_points.argtypes = [C.c_void_p]
        L.kor_map_num_points.restype = C.c_int64
        L.kor_map_num_voxels.argtypes = [C.c_void_p]
        L.kor_map_num_voxels.restype = C.c_int64
        L.kor_map_add_points.argtypes = [C.c_void_p, c_dp, C.c_int64]
        L.kor_map_remove_far.argtypes = [C.c_void_p, c_dp]
        L.kor_map_update_origin.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp]
        L.kor_map_update_pose.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp]
        L.kor_map_pointcloud.argtypes = [C.c_void_p, c_dp, C.c_int64]
        L.kor_map_pointcloud.restype = C.c_int64
        L.kor_map_export_voxels.argtypes = [C.c_void_p, c_ip, c_ip, c_dp]
        L.kor_map_export_voxels.restype = C.c_int64
        L.kor_map_nearest.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp, c_dp]
        L.kor_map_neighbourhood_stats.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp, c_dp, c_dp]
        L.kor_register.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp, c_dp, C.c_double, C.c_int, C.c_double, C.c_int,
                                   C.c_double, C.c_int, c_dp, C.POINTER(Stats)]
        L.kor_voxel_downsample.argtypes = [c_dp, C.c_int64, C.c_double, c_dp]
        L.kor_voxel_downsample.restype = C.c_int64
        L.kor_preprocess.argtypes = [c_dp, C.c_int64, c_dp, C.c_int64, c_dp, C.c_double, C.c_double, C.c_int, c_dp]
        L.kor_preprocess.restype = C.c_int64
        L.kor_se3_exp.argtypes = [c_dp, c_dp]
        L.kor_se3_log.argtypes = [c_dp, c_dp]
        L.kor_se3_compose.argtypes = [c_dp, c_dp, c_dp]
        L.kor_se3_inverse.argtypes = [c_dp, c_dp]
        L.kor_se3_transform.argtypes = [c_dp, c_dp, C.c_int64, c_dp]
        L.kor_threshold_create.restype = C.c_void_p
        L.kor_threshold_create.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double]
        L.kor_threshold_destroy.argtypes = [C.c_void_p]
        L.kor_threshold_update.argtypes = [C.c_void_p, c_dp]
        L.kor_threshold_compute.argtypes = [C.c_void_p]
        L.kor_threshold_compute.restype = C.c_double
        L.kor_threshold_reset.argtypes = [C.c_void_p]
        L.kor_synth_scan.restype = C.c_int64
        L.kor_synth_scan.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                                     C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_double,
                                     C.c_int, c_dp, C.c_int64]
        _LIB = L
    return _LIB


def pose7(v):
    a = np.ascontiguousarray(v, dtype=np.float64).reshape(7)
    return a


IDENTITY = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


def se3_exp(tangent6):
    out = np.empty(7)
    lib().kor_se3_exp(_dp(np.ascontiguousarray(tangent6, dtype=np.float64)), _dp(out))
    return out


def se3_log(p7):
    out = np.empty(6)
    lib().kor_se3_log(_dp(pose7(p7)), _dp(out))
    return out


def se3_compose(a, b):
    out = np.empty(7)
    lib().kor_se3_compose(_dp(pose7(a)), _dp(pose7(b)), _dp(out))
    return out


def se3_inverse(a):
    out = np.empty(7)
    lib().kor_se3_inverse(_dp(pose7(a)), _dp(out))
    return out


def se3_transform(p7, pts):
    pts = _pts(pts)
    out = np.empty_like(pts)
    lib().kor_se3_transform(_dp(pose7(p7)), _dp(pts), len(pts), _dp(out))
    return out


def planar_pose(x, y, yaw):
    """SE3 pose of a planar robot as pose7."""
    return np.array([0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2), x, y, 0.0])


def pose_delta(a7, b7):
    """(translation L2, rotation angle of Ra^-1 Rb) between two pose7 — the parity metric of SURVEY.md §8(d)."""
    a7, b7 = pose7(a7), pose7(b7)
    dt = float(np.linalg.norm(a7[4:] - b7[4:]))
    qa, qb = a7[:4], b7[:4]
    # relative quaternion qa^-1 * qb, angle = 2 atan2(|vec|, |w|)
    ax, ay, az, aw = -qa[0], -qa[1], -qa[2], qa[3]
    bx, by, bz, bw = qb
    w = aw * bw - ax * bx - ay * by - az * bz
    x = aw * bx + ax * bw + ay * bz - az * by
    y = aw * by + ay * bw + az * bx - ax * bz
    z = aw * bz + az * bw + ax * by - ay * bx
    ang = 2.0 * float(np.arctan2(np.sqrt(x * x + y * y + z * z), abs(w)))
    return dt, ang


class OracleMap:
    """kiss_icp::VoxelHashMap restated on the CPU (KISS-ICP v1.2.0 semantics)."""

    def __init__(self, voxel_size=1.0, max_distance=100.0, max_points_per_voxel=20):
        self.voxel_size, self.max_distance, self.max_points_per_voxel = voxel_size, max_distance, max_points_per_voxel
        self.h = C.c_void_p(lib().kor_map_create(voxel_size, max_distance, max_points_per_voxel))

    def __del__(self):
        if getattr(self, "h", None) and _LIB is not None:
            _LIB.kor_map_destroy(self.h)
            self.h = None

    def clear(self):
        lib().kor_map_clear(self.h)

    def empty(self):
        return bool(lib().kor_map_empty(self.h))

    def num_points(self):
        return int(lib().kor_map_num_points(self.h))

    def num_voxels(self):
        return int(lib().kor_map_num_voxels(self.h))

    def add_points(self, pts):
        pts = _pts(pts)
        lib().kor_map_add_points(self.h, _dp(pts), len(pts))

    def remove_far(self, origin):
        o = np.ascontiguousarray(origin, dtype=np.float64)
        lib().kor_map_remove_far(self.h, _dp(o))

    def update_origin(self, pts, origin):
        pts = _pts(pts)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        lib().kor_map_update_origin(self.h, _dp(pts), len(pts), _dp(o))

    def update_pose(self, pts, p7):
        pts = _pts(pts)
        lib().kor_map_update_pose(self.h, _dp(pts), len(pts), _dp(pose7(p7)))

    def pointcloud(self):
        n = self.num_points()
        out = np.empty((max(n, 1), 3))
        m = lib().kor_map_pointcloud(self.h, _dp(out), n)
        return out[:m].copy()

    def export_voxels(self):
        """(keys[V,3] int32 sorted lexicographically, counts[V] int32, points[total,3])."""
        V, n = self.num_voxels(), self.num_points()
        keys = np.empty((max(V, 1), 3), dtype=np.int32)
        counts = np.empty(max(V, 1), dtype=np.int32)
        pts = np.empty((max(n, 1), 3))
        lib().kor_map_export_voxels(self.h, keys.ctypes.data_as(c_ip), counts.ctypes.data_as(c_ip), _dp(pts))
        return keys[:V].copy(), counts[:V].copy(), pts[:n].copy()

    def nearest(self, q):
        q = _pts(q)
        out = np.empty_like(q)
        d = np.empty(len(q))
        lib().kor_map_nearest(self.h, _dp(q), len(q), _dp(out), _dp(d))
        return out, d

    def neighbourhood_stats(self, pts, p7):
        pts = _pts(pts)
        a, b = C.c_double(), C.c_double()
        lib().kor_map_neighbourhood_stats(self.h, _dp(pts), len(pts), _dp(pose7(p7)), C.byref(a), C.byref(b))
        return a.value, b.value

    def register(self, frame, last_pose, rel_odom, tau, max_iter=10, conv=1e-3, adaptive=True, fixed_reg=0.0, threads=1):
        """KinematicRegistration::ComputeRobotMotion (Registration.cpp:151-190).  Returns (pose7, Stats)."""
        frame = _pts(frame)
        out = np.empty(7)
        st = Stats()
        lib().kor_register(self.h, _dp(frame), len(frame), _dp(pose7(last_pose)), _dp(pose7(rel_odom)), float(tau),
                           int(max_iter), float(conv), int(bool(adaptive)), float(fixed_reg), int(threads), _dp(out),
                           C.byref(st))
        return out, st


def set_downsample_order(mode):
    """Output order of VoxelDownsample in the oracle AND in the reference build (oracle/_ref holds its own copy of the restated
    KISS-ICP code): 0 = order of first occurrence (default; what the device emits), 1 = tsl::robin_map iteration order as recalled,
    2 = the same with the pre-v1.0 20-bit hash mask.  Modes 1 / 2 are unpinned (kicp_oracle.hpp)."""
    L = lib()
    L.kor_set_downsample_order.argtypes = [C.c_int]
    L.kor_set_downsample_order(int(mode))
    if ref_available():
        R = ref_lib()
        R.kref_set_downsample_order.argtypes = [C.c_int]
        R.kref_set_downsample_order(int(mode))


def voxel_downsample(pts, voxel_size):
    pts = _pts(pts)
    out = np.empty_like(pts)
    n = lib().kor_voxel_downsample(_dp(pts), len(pts), float(voxel_size), _dp(out))
    return out[:n].copy()


def preprocess(pts, stamps, rel_motion, max_range, min_range, deskew):
    pts = _pts(pts)
    stamps = np.ascontiguousarray(stamps, dtype=np.float64)
    out = np.empty_like(pts)
    n = lib().kor_preprocess(_dp(pts), len(pts), _dp(stamps), len(stamps), _dp(pose7(rel_motion)), float(max_range),
                             float(min_range), int(bool(deskew)), _dp(out))
    return out[:n].copy()


class OracleThreshold:
    def __init__(self, map_err, max_range, adaptive=True, fixed=1.0):
        self.h = C.c_void_p(lib().kor_threshold_create(map_err, max_range, int(adaptive), fixed))

    def update(self, err7):
        lib().kor_threshold_update(self.h, _dp(pose7(err7)))

    def compute(self):
        return float(lib().kor_threshold_compute(self.h))

    def reset(self):
        lib().kor_threshold_reset(self.h)

    def __del__(self):
        if getattr(self, "h", None) and _LIB is not None:
            _LIB.kor_threshold_destroy(self.h)
            self.h = None


def synth_scan(n_beams, elev_min, elev_max, n_az, x, y, yaw, seed, sigma=0.02, half_extent=95.0, pitch=6.0, wall_h=6.0,
               sensor_h=1.8, max_range=100.0, round_f32=True):
    cap = n_beams * n_az
    out = np.empty((cap, 3))
    n = lib().kor_synth_scan(half_extent, pitch, wall_h, n_beams, elev_min, elev_max, n_az, x, y, yaw, sensor_h, sigma,
                             seed, max_range, int(round_f32), _dp(out), cap)
    assert n >= 0
    return out[:n].copy()


# ----------------------------------------------------------------------------------------------
# The reference's own Registration.cpp compiled here against header shims (oracle/_ref).
# ----------------------------------------------------------------------------------------------
def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libkicp_ref.so"))


def ref_lib():
    global _REF
    if _REF is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libkicp_ref.so"))
        L.kref_map_create.restype = C.c_void_p
        L.kref_map_create.argtypes = [C.c_double, C.c_double, C.c_uint]
        L.kref_map_destroy.argtypes = [C.c_void_p]
        L.kref_map_add_points.argtypes = [C.c_void_p, c_dp, C.c_int64]
        L.kref_map_update_pose.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp]
        L.kref_map_num_points.argtypes = [C.c_void_p]
        L.kref_map_num_points.restype = C.c_int64
        L.kref_register.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp, c_dp, C.c_double, C.c_int, C.c_double, C.c_int,
                                    C.c_double, C.c_int, c_dp]
        L.kref_threshold_sequence.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, c_dp, C.c_int64, c_dp]
        L.kref_pipeline_create.restype = C.c_void_p
        L.kref_pipeline_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint, C.c_int, C.c_double, C.c_int,
                                           C.c_double, C.c_int, C.c_int, C.c_double, C.c_int]
        L.kref_pipeline_destroy.argtypes = [C.c_void_p]
        L.kref_pipeline_set_pose.argtypes = [C.c_void_p, c_dp]
        L.kref_pipeline_register_frame.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp, C.c_int64, c_dp, c_dp, c_dp]
        L.kref_pipeline_register_frame.restype = C.c_int64
        L.kref_pipeline_num_map_points.argtypes = [C.c_void_p]
        L.kref_pipeline_num_map_points.restype = C.c_int64
        _REF = L
    return _REF


class RefMap:
    """The oracle's KISS map exposed to the reference's own Registration.cpp (oracle/_ref)."""

    def __init__(self, voxel_size=1.0, max_distance=100.0, max_points_per_voxel=20):
        self.h = C.c_void_p(ref_lib().kref_map_create(voxel_size, max_distance, max_points_per_voxel))

    def __del__(self):
        if getattr(self, "h", None) and _REF is not None:
            _REF.kref_map_destroy(self.h)
            self.h = None

    def add_points(self, pts):
        pts = _pts(pts)
        ref_lib().kref_map_add_points(self.h, _dp(pts), len(pts))

    def update_pose(self, pts, p7):
        pts = _pts(pts)
        ref_lib().kref_map_update_pose(self.h, _dp(pts), len(pts), _dp(pose7(p7)))

    def num_points(self):
        return int(ref_lib().kref_map_num_points(self.h))

    def register(self, frame, last_pose, rel_odom, tau, max_iter=10, conv=1e-3, adaptive=True, fixed_reg=0.0, threads=1):
        frame = _pts(frame)
        out = np.empty(7)
        ref_lib().kref_register(self.h, _dp(frame), len(frame), _dp(pose7(last_pose)), _dp(pose7(rel_odom)), float(tau),
                                int(max_iter), float(conv), int(bool(adaptive)), float(fixed_reg), int(threads), _dp(out))
        return out


class _Pipeline:
    """kinematic_icp::pipeline::KinematicICP behind one of three builds with the same C entry points:
       prefix kref_  oracle/_ref/libkicp_ref.so      the reference's pipeline + the reference's Registration.cpp + CPU map
       prefix kgpu_  oracle/_ref/libkicp_ref_gpu.so  the reference's pipeline source over the product's GPU facade
       prefix kfac_  tests/hooks/_build/libkicp_facade_hooks.so        the product's own facade pipeline (GPU), through the test hooks."""

    def __init__(self, L, prefix, max_range=100.0, min_range=0.0, voxel_size=1.0, max_points_per_voxel=20, use_adaptive_threshold=True,
                 fixed_threshold=1.0, max_num_iterations=10, convergence_criterion=1e-3, max_num_threads=1, use_adaptive_reg=True,
                 fixed_reg=0.0, deskew=False):
        self.L, self.p = L, prefix
        f = getattr(L, prefix + "pipeline_create")
        f.restype = C.c_void_p
        f.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int,
                      C.c_double, C.c_int]
        getattr(L, prefix + "pipeline_destroy").argtypes = [C.c_void_p]
        getattr(L, prefix + "pipeline_set_pose").argtypes = [C.c_void_p, c_dp]
        g = getattr(L, prefix + "pipeline_register_frame")
        g.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp, C.c_int64, c_dp, c_dp, c_dp]
        g.restype = C.c_int64
        h = getattr(L, prefix + "pipeline_num_map_points")
        h.argtypes = [C.c_void_p]
        h.restype = C.c_int64
        self.h = C.c_void_p(f(max_range, min_range, voxel_size, max_points_per_voxel, int(use_adaptive_threshold), fixed_threshold,
                              max_num_iterations, convergence_criterion, max_num_threads, int(use_adaptive_reg), fixed_reg, int(deskew)))
        if not self.h:
            raise RuntimeError("pipeline_create failed (no CUDA device?)")

    def set_pose(self, p7):
        getattr(self.L, self.p + "pipeline_set_pose")(self.h, _dp(pose7(p7)))

    def register_frame(self, pts, stamps, lidar_to_base, rel_odom):
        pts = _pts(pts)
        stamps = np.ascontiguousarray(stamps, dtype=np.float64)
        out = np.empty(7)
        n_src = getattr(self.L, self.p + "pipeline_register_frame")(self.h, _dp(pts), len(pts), _dp(stamps), len(stamps),
                                                                     _dp(pose7(lidar_to_base)), _dp(pose7(rel_odom)), _dp(out))
        return out, int(n_src)

    def register_frame_raw(self, pts, stamps, lidar_to_base, rel_odom):
        """Facade only: a packed float64 or float32 [n,3] buffer straight to the device (kicp_frame_input), no std::vector copy."""
        pts = np.ascontiguousarray(pts).reshape(-1, 3)
        assert pts.dtype in (np.float32, np.float64)
        stamps = np.ascontiguousarray(stamps, dtype=np.float64)
        out = np.empty(7)
        g = getattr(self.L, self.p + "pipeline_register_frame_raw")
        g.restype = C.c_int64
        g.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_dp, C.c_int64, c_dp,
                      c_dp, c_dp]
        n_src = g(self.h, pts.ctypes.data, len(pts), 1 if pts.dtype == np.float32 else 0, 0, 0, 0, 0, _dp(stamps), len(stamps),
                  _dp(pose7(lidar_to_base)), _dp(pose7(rel_odom)), _dp(out))
        return out, int(n_src)

    def num_map_points(self):
        return int(getattr(self.L, self.p + "pipeline_num_map_points")(self.h))

    def close(self):
        if self.h:
            getattr(self.L, self.p + "pipeline_destroy")(self.h)
            self.h = None


def ref_pipeline(**kw):
    return _Pipeline(ref_lib(), "kref_", **kw)


def ref_gpu_pipeline(**kw):
    return _Pipeline(C.CDLL(os.path.join(_HERE, "_ref", "libkicp_ref_gpu.so")), "kgpu_", **kw)


def ref_gpu_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libkicp_ref_gpu.so"))


def facade_hooks_path():
    """extern "C" hooks over the product's C++ facade (tests/hooks/facade_hooks.cpp, built by kinematic-icp_b200/cpp/Makefile)."""
    return os.path.normpath(os.path.join(_HERE, "..", "tests", "hooks", "_build", "libkicp_facade_hooks.so"))


def facade_pipeline(**kw):
    return _Pipeline(C.CDLL(facade_hooks_path()), "kfac_", **kw)
