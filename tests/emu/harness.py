"""Python side of the emulation harnesses (test infrastructure): builds tests/emu/{kr,km,kf}_emu.cpp — the registration, map and
front-end kernels' unchanged sources compiled for the host against cuda_emu.hpp — and wraps their C entry points."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EMU = os.path.join(ROOT, "tests", "emu")
_LIBS = {}


def build(name):
    """g++ -ffp-contract=off (the product compiles the map / front-end kernels with -fmad=false) -> tests/emu/_build/lib<name>.so"""
    if name not in _LIBS:
        out = os.path.join(EMU, "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "lib%s.so" % name)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "kinematic-icp_b200", "csrc"), "-o", so, os.path.join(EMU, name + ".cpp"), "-lpthread"],
                       check=True)
        _LIBS[name] = C.CDLL(so)
    return _LIBS[name]


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------------ registration (kr_emu.cpp)
def kr_lib():
    return build("kr_emu")


def register(om_or_voxels, scan, last, odom, tau, grid=3, nranks=1, persistent=1, nn_cache=1, registrations=1, max_iter=10, conv=1e-3,
             adaptive=True, fixed_reg=0.0, late_upload=0, device_count=-1, voxel_size=None, cap=None):
    """One registration (or several in a row) of `scan` against a map given as an oracle map or as (keys, counts, points) + voxel_size
    + cap, on `nranks` emulated GPUs of `grid` CTAs each.  Returns ([RegResult per rank], [probes, candidate points, lines])."""
    from kinematic_icp_b200 import _capi
    lib = kr_lib()
    if isinstance(om_or_voxels, tuple):
        keys, counts, pts = om_or_voxels
    else:
        keys, counts, pts = om_or_voxels.export_voxels()
        voxel_size, cap = om_or_voxels.voxel_size, om_or_voxels.max_points_per_voxel
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    f32 = scan.dtype == np.float32
    scan = np.ascontiguousarray(scan)
    last, odom = np.ascontiguousarray(last, dtype=np.float64), np.ascontiguousarray(odom, dtype=np.float64)
    p = _capi.RegParams(max_iter, 1 if adaptive else 0, conv, fixed_reg)
    res = (_capi.RegResult * nranks)()
    stats = (C.c_uint64 * 3)()
    rc = lib.kr_emu_register(_vp(keys), _vp(counts), _vp(pts), C.c_int64(len(counts)), C.c_int32(cap), C.c_double(voxel_size), _vp(scan),
                             C.c_int64(len(scan)), C.c_int32(1 if f32 else 0), _vp(last), _vp(odom), C.c_double(tau), C.byref(p), C.c_int32(grid),
                             C.c_int32(nranks), C.c_int32(persistent), C.c_int32(nn_cache), C.c_int32(registrations), res, stats,
                             C.c_int32(late_upload), C.c_int32(device_count))
    assert rc == 0, "a launch must leave its counters zero for the next one: rc %d" % rc
    return list(res), list(stats)


# --------------------------------------------------------------------------------------------------------- map (km_emu.cpp)
def km_lib():
    L = build("km_emu")
    if not getattr(L, "_typed", False):
        L.km_emu_create.restype = C.c_void_p
        L.km_emu_create.argtypes = [C.c_double, C.c_double, C.c_int32, C.c_uint32]
        L.km_emu_destroy.argtypes = [C.c_void_p]
        for f in (L.km_emu_num_points, L.km_emu_num_voxels):
            f.restype, f.argtypes = C.c_int64, [C.c_void_p]
        L.km_emu_add_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.km_emu_remove_far.argtypes = [C.c_void_p, C.c_void_p]
        L.km_emu_exclusive_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.km_emu_exclusive_sum.restype = None
        L.km_emu_update_pose_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32]
        L.km_emu_load_voxels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.km_emu_export.restype = C.c_int64
        L.km_emu_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.km_emu_nearest.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L._typed = True
    return L


class EmuMap:
    """kiss_icp::VoxelHashMap on the emulated device (the kernels of kicp_map_kernels.cuh, host memory for the slab)."""

    def __init__(self, voxel_size, max_distance, cap, blocks_cap=60000):
        self.L = km_lib()
        self.voxel_size, self.max_distance, self.cap = voxel_size, max_distance, cap
        self.h = C.c_void_p(self.L.km_emu_create(voxel_size, max_distance, cap, blocks_cap))

    def add_points(self, pts, pose=None):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        p = None if pose is None else np.ascontiguousarray(pose, dtype=np.float64)
        assert self.L.km_emu_add_points(self.h, pts.ctypes.data, len(pts), None if p is None else p.ctypes.data) == 0

    def update_pose_async(self, pts, n_actual, pose, status=0):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        p = np.ascontiguousarray(pose, dtype=np.float64)
        assert self.L.km_emu_update_pose_async(self.h, pts.ctypes.data, len(pts), n_actual, p.ctypes.data, status) == 0

    def remove_far(self, origin):
        o = np.ascontiguousarray(origin, dtype=np.float64)
        assert self.L.km_emu_remove_far(self.h, o.ctypes.data) == 0

    def load_voxels(self, keys, counts, pts):
        keys = np.ascontiguousarray(keys, dtype=np.int32)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        assert self.L.km_emu_load_voxels(self.h, keys.ctypes.data, counts.ctypes.data, pts.ctypes.data, len(counts)) == 0

    def num_points(self):
        return int(self.L.km_emu_num_points(self.h))

    def num_voxels(self):
        return int(self.L.km_emu_num_voxels(self.h))

    def export_voxels(self):
        nv, npts = self.num_voxels(), self.num_points()
        keys, counts, pts = np.zeros((nv, 3), np.int32), np.zeros(nv, np.int32), np.zeros((npts, 3))
        assert self.L.km_emu_export(self.h, keys.ctypes.data, counts.ctypes.data, pts.ctypes.data) == nv
        assert int(counts.sum()) == npts
        return keys, counts, pts

    def nearest(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        out_p, out_d = np.zeros((len(q), 3)), np.zeros(len(q))
        assert self.L.km_emu_nearest(self.h, q.ctypes.data, len(q), out_p.ctypes.data, out_d.ctypes.data) == 0
        return out_p, out_d

    def close(self):
        if self.h:
            self.L.km_emu_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------------------------------- front end (kf_emu.cpp)
def kf_lib():
    L = build("kf_emu")
    if not getattr(L, "_typed", False):
        L.kf_emu_voxel_downsample.restype = C.c_int64
        L.kf_emu_voxel_downsample.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int32]
        L.kf_emu_preprocess.restype = C.c_int64
        L.kf_emu_preprocess.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int32,
                                        C.c_void_p]
        L.kf_emu_ingest.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L._typed = True
    return L


def downsample(pts, vs, n_actual=-1):
    """kiss_icp::VoxelDownsample; n_actual >= 0: the point count is read by the kernels from a device word (a stage of a frame)."""
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((max(len(pts), 1), 3))
    m = kf_lib().kf_emu_voxel_downsample(pts.ctypes.data, len(pts), vs, out.ctypes.data, n_actual)
    return out[:m]


def preprocess(ko, pts, stamps, motion, max_range, min_range, deskew, lidar_to_base=None):
    """kiss_icp::Preprocessor::Preprocess (+ the transform to the base frame, fused as in kicp_register_frame)."""
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    stamps = np.ascontiguousarray(stamps, dtype=np.float64)
    omega = np.ascontiguousarray(ko.se3_log(motion), dtype=np.float64)
    l2b = np.ascontiguousarray(ko.IDENTITY if lidar_to_base is None else lidar_to_base, dtype=np.float64)
    out = np.zeros((max(len(pts), 1), 3))
    m = kf_lib().kf_emu_preprocess(pts.ctypes.data, len(pts), stamps.ctypes.data if len(stamps) else None, len(stamps), omega.ctypes.data,
                                   l2b.ctypes.data, max_range, min_range, 1 if deskew else 0, out.ctypes.data)
    return out[:m]
