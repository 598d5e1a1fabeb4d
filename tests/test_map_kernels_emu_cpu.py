"""The voxel-map kernels' LOGIC, checked without a GPU: kinematic-icp_b200/csrc/kicp_map_kernels.cuh (find-or-create with slot
locking, the per-voxel replay of the greedy insert rule in input order, eviction + order-preserving compaction + table rebuild, bulk
load, the batched nearest-neighbour query) compiled unchanged by g++ against the SIMT emulator (tests/emu/cuda_emu.hpp; CTAs handed
to a few OS threads, fibers inside) and driven by the launch sequences of kicp_map.cu restated on host memory (tests/emu/km_emu.cpp).
The map must equal the CPU oracle's bit for bit — the same assertions as tests/test_gpu_parity.py makes on the device.

Test infrastructure; the GPU suite remains the proof for the device build."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBL_MAX = np.finfo(np.float64).max


from emu import harness as H


@pytest.fixture(scope="module")
def emu():
    return H.km_lib()


def EmuMap(_lib, voxel_size, max_distance, cap):
    return H.EmuMap(voxel_size, max_distance, cap)


def sorted_voxels(keys, counts, pts):
    off = np.concatenate([[0], np.cumsum(counts)])
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return keys[order], counts[order], np.concatenate([pts[off[i]:off[i + 1]] for i in order]) if len(order) else pts


def same_map(em, om):
    assert em.num_points() == om.num_points() and em.num_voxels() == om.num_voxels()
    k1, c1, p1 = sorted_voxels(*em.export_voxels())
    k0, c0, p0 = sorted_voxels(*om.export_voxels())
    assert np.array_equal(k1, k0) and np.array_equal(c1, c0) and np.array_equal(p1, p0)


def test_addpoints_bit_exact(emu, oracle):
    """AddPoints reproduces the greedy, input-order dependent CPU result exactly (clustered points, negative coordinates, repeated
    inserts, a dense blob with hundreds of candidates per voxel: the long pending lists of k_add_commit)."""
    ko = oracle
    rng = np.random.default_rng(11)
    om = ko.OracleMap(1.0, 100.0, 20)
    em = EmuMap(emu, 1.0, 100.0, 20)
    for it in range(5):
        pts = rng.normal(size=(4000, 3)) * [6.0, 6.0, 1.5] + [-2.0, 1.0, 0.0]
        if it == 3:
            pts = np.concatenate([pts, rng.uniform(-1.0, 1.0, size=(2500, 3)) + [20.0, -7.0, 0.5]])
        om.add_points(pts)
        em.add_points(pts)
        assert em.num_points() == om.num_points() and em.num_voxels() == om.num_voxels()
    same_map(em, om)
    em.close()


def test_update_and_eviction_bit_exact(emu, oracle):
    """Update(points, origin) = AddPoints + RemovePointsFarFromLocation: first-point rule, >= max_distance, order-preserving compaction."""
    ko = oracle
    rng = np.random.default_rng(12)
    om = ko.OracleMap(0.5, 12.0, 8)
    em = EmuMap(emu, 0.5, 12.0, 8)
    for it in range(6):
        origin = np.array([3.0 * it, 0.5 * it, 0.0])
        pts = rng.uniform(-10, 10, size=(4000, 3)) * [1, 1, 0.2] + origin
        om.update_origin(pts, origin)
        em.add_points(pts)
        em.remove_far(origin)
        assert em.num_points() == om.num_points() and em.num_voxels() == om.num_voxels()
    same_map(em, om)
    em.close()


def test_update_pose_bit_exact(emu, oracle, workload):
    """Update(points, pose): the pose applied by the kernel (Sophus quaternion formula, no FMA contraction)."""
    ko = oracle
    w = workload(1)
    om = ko.OracleMap(1.0, 100.0, 20)
    em = EmuMap(emu, 1.0, 100.0, 20)
    pose = ko.planar_pose(50.0, 0.0, 1.2)
    for k in range(3):
        om.update_pose(w.scan, pose)
        em.add_points(w.scan, pose)
        em.remove_far(pose[4:])
        pose = ko.se3_compose(pose, ko.se3_exp([0.8, 0.05, 0, 0, 0, 0.03]))
    same_map(em, om)
    em.close()


def test_frame_path_update_with_count_and_pose_on_the_device(emu, oracle, workload):
    """The map update of kicp_register_frame: buffer sized for the worst case, the actual survivor count and the pose read by the
    kernels themselves, eviction over an upper bound of the block count; a registration that failed (NaN pose) leaves the map alone."""
    ko = oracle
    w = workload(1)
    om = ko.OracleMap(1.0, 40.0, 20)
    em = EmuMap(emu, 1.0, 40.0, 20)
    pose = ko.planar_pose(50.0, 0.0, 1.2)
    rng = np.random.default_rng(8)
    for k in range(4):
        n_actual = int(rng.integers(len(w.scan) // 2, len(w.scan)))
        buf = np.concatenate([w.scan[:n_actual], np.full((len(w.scan) - n_actual, 3), 1e9)])  # the tail must never be read
        om.update_pose(w.scan[:n_actual], pose)
        em.update_pose_async(buf, n_actual, pose)
        assert em.num_points() == om.num_points() and em.num_voxels() == om.num_voxels()
        pose = ko.se3_compose(pose, ko.se3_exp([6.0, 0.3, 0, 0, 0, 0.05]))  # far enough for voxels to die (max_distance 40)
    same_map(em, om)
    before = em.export_voxels()
    em.update_pose_async(w.scan, len(w.scan), np.full(7, np.nan), status=16)  # KICP_WARN_NO_CORRESPONDENCES
    after = em.export_voxels()
    assert all(np.array_equal(a, b) for a, b in zip(before, after))
    em.close()


def test_nearest_neighbour_bit_exact(emu, oracle, workload):
    """GetClosestNeighbor on a bulk-loaded map: same point and same distance, bit for bit, including the empty neighbourhood."""
    ko = oracle
    w = workload(2)
    em = EmuMap(emu, w.voxel_size, w.max_range, w.max_points_per_voxel)
    em.load_voxels(*w.map.export_voxels())
    same_map(em, w.map)
    q = ko.se3_transform(w.prior, w.scan)[:6000]
    rng = np.random.default_rng(3)
    q = np.concatenate([q, rng.uniform(-120, 120, size=(1500, 3)), np.floor(q[:800]) + 0.0, -np.abs(q[:300])])
    pg, dg = em.nearest(q)
    po, do = w.map.nearest(q)
    assert np.array_equal(dg, do) and np.array_equal(pg, po)
    assert (do == DBL_MAX).sum() > 0
    em.close()


def test_single_pass_exclusive_sum(emu):
    """kicp_scan.cuh alone: the chained scan with decoupled look-back that renumbers the surviving blocks (and, fused into the front
    end's kernels, compacts the frame).  Sizes around the warp, tile (1 024 items) and look-back window (32 tiles) boundaries, all
    launches on ONE state: the status words of earlier launches are still there and must read as 'not written yet'."""
    rng = np.random.default_rng(77)
    sizes = [1, 2, 31, 32, 33, 127, 128, 129, 1023, 1024, 1025, 2048, 3000, 32 * 1024, 33 * 1024 - 1, 33 * 1024 + 5, 70001, 300000, 7, 1024]
    for rep, n in enumerate(sizes * 2):
        v = rng.integers(0, 21, size=n, dtype=np.uint32) if rep % 3 else np.ones(n, dtype=np.uint32)
        out = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
        emu.km_emu_exclusive_sum(v.ctypes.data, out.ctypes.data, n)
        ref = np.concatenate([[0], np.cumsum(v[:-1], dtype=np.uint64)]).astype(np.uint32)
        assert np.array_equal(out, ref), n


@pytest.mark.parametrize("workers", [1, 2, 24])
def test_single_pass_scans_with_other_amounts_of_concurrency(workers):
    """The look-back protocol must not care how many tiles are in flight: the scan and fused-select tests again in a fresh process
    with the CTAs of every launch handed to 1 OS thread (every predecessor has finished: the nearest status word is always a PREFIX),
    2 (the predecessor is usually still running) and 24 (many AGGREGATE words in the look-back window)."""
    import subprocess
    import sys
    env = dict(os.environ, KICP_EMU_WORKERS=str(workers))
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_map_kernels_emu_cpu.py"),
                        os.path.join(here, "test_frontend_kernels_emu_cpu.py"), "-x", "-q", "-k", "single_pass_exclusive_sum or fused_selects",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
