"""Row (f)#2 of SURVEY.md §8: kiss_icp::VoxelDownsample and Preprocessor::Preprocess on the device, against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_voxel_downsample_bit_exact(oracle, gpu_ctx, workload):
    """First point (input order) per voxel, survivors in input order — identical arrays."""
    import kinematic_icp_b200 as kb
    ko = oracle
    rng = np.random.default_rng(21)
    clouds = [rng.normal(size=(20000, 3)) * [8.0, 8.0, 1.0] - [3.0, 0.0, 0.5],  # clustered, negative coordinates
              workload(2).scan, workload(3).scan[::3], np.zeros((0, 3)), np.array([[0.25, -0.25, 7.0]])]
    for pts in clouds:
        for vs in (0.5, 1.5, 0.37):
            out = kb.VoxelDownsample(gpu_ctx, pts, vs)
            ref = ko.voxel_downsample(pts, vs) if len(pts) else np.zeros((0, 3))
            assert out.shape == ref.shape and np.array_equal(out, ref)
    # the pipeline's double down-sample (KinematicICP.cpp:38-44)
    w = workload(2)
    fd = kb.VoxelDownsample(gpu_ctx, w.scan, 0.5)
    src = kb.VoxelDownsample(gpu_ctx, fd, 1.5)
    assert np.array_equal(src, ko.voxel_downsample(ko.voxel_downsample(w.scan, 0.5), 1.5))


def test_preprocess_range_filter_and_transform_bit_exact(oracle, gpu_ctx, workload):
    import kinematic_icp_b200 as kb
    ko = oracle
    w = workload(2)
    pts = np.concatenate([w.scan, w.scan * 3.0, w.scan * 0.01])  # ranges from centimetres to beyond max_range
    out = kb.Preprocess(gpu_ctx, pts, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False)
    ref = ko.preprocess(pts, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False)
    assert np.array_equal(out, ref) and 0 < len(ref) < len(pts)
    # deskew requested but no stamps: the frame is used as is (Preprocessing.cpp)
    out = kb.Preprocess(gpu_ctx, pts, np.zeros(0), w.rel_odom, 100.0, 0.5, True)
    assert np.array_equal(out, ref)
    # with the transform to the base frame fused in
    l2b = ko.se3_exp([0.2, -0.1, 0.3, 0.01, -0.02, 0.05])
    out = kb.Preprocess(gpu_ctx, pts, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False, lidar_to_base=l2b)
    assert np.array_equal(out, ko.se3_transform(l2b, ref))


def test_preprocess_deskew(oracle, gpu_ctx, workload):
    """De-skew uses sin/cos: device and glibc differ in the last bits, so points agree to 1e-12 m and the kept set is equal."""
    import kinematic_icp_b200 as kb
    ko = oracle
    w = workload(2)
    rng = np.random.default_rng(5)
    stamps = rng.uniform(10.0, 10.1, size=len(w.scan))
    motion = ko.se3_exp([0.6, 0.02, 0.0, 0.001, -0.002, 0.03])
    out = kb.Preprocess(gpu_ctx, w.scan, stamps, motion, 100.0, 0.0, True)
    ref = ko.preprocess(w.scan, stamps, motion, 100.0, 0.0, True)
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() < 1e-12
