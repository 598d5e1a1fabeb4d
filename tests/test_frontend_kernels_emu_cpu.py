"""The front-end kernels' LOGIC, checked without a GPU: kinematic-icp_b200/csrc/kicp_frontend_kernels.cuh (VoxelDownsample's
min-index-per-voxel insert with slot locking, Preprocess with de-skew / range filter / base transform, the PointCloud2 ingest) compiled
unchanged by g++ against the SIMT emulator and driven by the launch sequences of kicp_frontend.cu restated on host memory
(tests/emu/kf_emu.cpp).  Same assertions as tests/test_gpu_frontend.py makes on the device.  Test infrastructure."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


from emu import harness as H


@pytest.fixture(scope="module")
def emu():
    return H.kf_lib()


def downsample(_lib, pts, vs, n_actual=-1):
    return H.downsample(pts, vs, n_actual)


def preprocess(_lib, ko, pts, stamps, motion, max_range, min_range, deskew, lidar_to_base=None):
    return H.preprocess(ko, pts, stamps, motion, max_range, min_range, deskew, lidar_to_base)


def test_voxel_downsample_bit_exact(emu, oracle, workload):
    """First point (input order) per voxel, survivors in input order — identical arrays; then the pipeline's double down-sample."""
    ko = oracle
    rng = np.random.default_rng(21)
    clouds = [rng.normal(size=(12000, 3)) * [8.0, 8.0, 1.0] - [3.0, 0.0, 0.5], workload(2).scan, np.zeros((0, 3)),
              np.array([[0.25, -0.25, 7.0]])]
    for pts in clouds:
        for vs in (0.5, 1.5, 0.37):
            out = downsample(emu, pts, vs)
            ref = ko.voxel_downsample(pts, vs) if len(pts) else np.zeros((0, 3))
            assert out.shape == ref.shape and np.array_equal(out, ref)
    w = workload(2)
    src = downsample(emu, downsample(emu, w.scan, 0.5), 1.5)
    assert np.array_equal(src, ko.voxel_downsample(ko.voxel_downsample(w.scan, 0.5), 1.5))
    # a stage of a frame: the buffer is sized for the worst case, the survivor count of the stage before sits in a device word,
    # the tail of the buffer (here: points that would open voxels of their own) must not be seen
    for n_actual in (0, 1, 5000, len(w.scan)):
        buf = np.concatenate([w.scan[:n_actual], w.scan[n_actual:] + 1000.0])
        out = downsample(emu, buf, 0.5, n_actual=n_actual)
        ref = ko.voxel_downsample(w.scan[:n_actual], 0.5) if n_actual else np.zeros((0, 3))
        assert out.shape == ref.shape and np.array_equal(out, ref)


def test_preprocess_bit_exact_and_deskew(emu, oracle, workload):
    ko = oracle
    w = workload(2)
    pts = np.concatenate([w.scan, w.scan * 3.0, w.scan * 0.01])  # ranges from centimetres to beyond max_range
    out = preprocess(emu, ko, pts, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False)
    ref = ko.preprocess(pts, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False)
    assert np.array_equal(out, ref) and 0 < len(ref) < len(pts)
    # deskew requested but no stamps: the frame is used as is (Preprocessing.cpp)
    assert np.array_equal(preprocess(emu, ko, pts, np.zeros(0), w.rel_odom, 100.0, 0.5, True), ref)
    # with the transform to the base frame fused in
    l2b = ko.se3_exp([0.2, -0.1, 0.3, 0.01, -0.02, 0.05])
    out = preprocess(emu, ko, pts, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False, lidar_to_base=l2b)
    assert np.array_equal(out, ko.se3_transform(l2b, ref))
    # de-skew: same kept set, points to 1e-12 m (the tolerance of the GPU test, where sin / cos differ from glibc's in the last bits)
    rng = np.random.default_rng(5)
    stamps = rng.uniform(10.0, 10.1, size=len(w.scan))
    motion = ko.se3_exp([0.6, 0.02, 0.0, 0.001, -0.002, 0.03])
    out = preprocess(emu, ko, w.scan, stamps, motion, 100.0, 0.0, True)
    ref = ko.preprocess(w.scan, stamps, motion, 100.0, 0.0, True)
    assert out.shape == ref.shape and np.abs(out - ref).max() < 1e-12


def test_ingest_widens_pointcloud2_fields(emu):
    """float32 x, y, z at unaligned offsets inside a 22-byte record, and float64 fields at a stride: widened exactly."""
    rng = np.random.default_rng(9)
    n = 3000
    xyz32 = rng.normal(size=(n, 3)).astype(np.float32) * 30
    rec = np.zeros(n, dtype=np.dtype({"names": ["pad", "x", "y", "z", "i"], "formats": ["u1", "<f4", "<f4", "<f4", "<f4"], "offsets": [0, 1, 5, 9, 13],
                                     "itemsize": 22}))
    rec["x"], rec["y"], rec["z"] = xyz32[:, 0], xyz32[:, 1], xyz32[:, 2]
    raw = np.frombuffer(rec.tobytes(), dtype=np.uint8).copy()
    out = np.zeros((n, 3))
    emu.kf_emu_ingest(raw.ctypes.data, n, 1, 22, 1, 5, 9, out.ctypes.data)
    assert np.array_equal(out, xyz32.astype(np.float64))
    xyz64 = rng.normal(size=(n, 3)) * 30
    rec = np.zeros(n, dtype=np.dtype({"names": ["x", "y", "z", "t"], "formats": ["<f8", "<f8", "<f8", "<f8"], "offsets": [0, 8, 16, 24], "itemsize": 40}))
    rec["x"], rec["y"], rec["z"] = xyz64[:, 0], xyz64[:, 1], xyz64[:, 2]
    raw = np.frombuffer(rec.tobytes(), dtype=np.uint8).copy()
    emu.kf_emu_ingest(raw.ctypes.data, n, 0, 40, 0, 8, 16, out.ctypes.data)
    assert np.array_equal(out, xyz64)


def test_fused_selects_over_many_tiles(emu, oracle):
    """The order-preserving selects fused into k_ds_select / k_preprocess_select (kicp_scan.cuh) on frames of a few hundred tiles:
    look-back over more than one window of 32 tiles, tiles with no survivor at all, everything kept, a device-resident count that ends
    in the middle of a tile — survivors in input order, bit for bit, and the count the last tile writes."""
    ko = oracle
    rng = np.random.default_rng(404)
    n = 150_000
    pts = rng.normal(size=(n, 3)) * [30.0, 30.0, 2.0]
    pts[40_000:90_000] = pts[:50_000] * 1e-3 + 5.0  # 49 consecutive tiles whose points all fall into a handful of voxels already taken
    for vs in (0.5, 1.5):
        out = downsample(emu, pts, vs)
        assert np.array_equal(out, ko.voxel_downsample(pts, vs))
    for n_actual in (1023, 1024, 1025, 40_000, 100_001):
        buf = np.concatenate([pts[:n_actual], pts[n_actual:] + 1000.0])
        assert np.array_equal(downsample(emu, buf, 0.5, n_actual=n_actual), ko.voxel_downsample(pts[:n_actual], 0.5))
    # Preprocess: nothing kept, everything kept, a long gap in the middle
    far = pts * 100.0
    assert len(preprocess(emu, ko, far, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False)) == len(ko.preprocess(far, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False))
    keep_all = preprocess(emu, ko, pts, np.zeros(0), ko.IDENTITY, 1e9, 0.0, False)
    assert np.array_equal(keep_all, ko.preprocess(pts, np.zeros(0), ko.IDENTITY, 1e9, 0.0, False)) and len(keep_all) == n
    gap = pts.copy()
    gap[10_000:120_000] *= 1e-4  # inside min_range
    out = preprocess(emu, ko, gap, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False)
    ref = ko.preprocess(gap, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False)
    assert np.array_equal(out, ref) and 0 < len(ref) < n
    # de-skew on a frame of many tiles: the stamps' min / max come from k_stamp_minmax (several CTAs + the last one's final pass)
    stamps = rng.uniform(3.0, 3.1, size=n)
    stamps[77_777], stamps[3] = 2.95, 3.15  # the extremes sit in different CTAs' shares
    motion = ko.se3_exp([0.6, 0.02, 0.0, 0.001, -0.002, 0.03])
    out = preprocess(emu, ko, pts, stamps, motion, 100.0, 0.5, True)
    ref = ko.preprocess(pts, stamps, motion, 100.0, 0.5, True)
    assert out.shape == ref.shape and np.abs(out - ref).max() < 1e-12
