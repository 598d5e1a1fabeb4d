"""GPU parity: the CUDA path through the C-ABI against the CPU oracle and the reference-derived golden vectors.

Tolerance (BASELINE.json north_star): final pose within 1e-6 m / 1e-7 rad of the CPU reference.  The map and the
nearest-neighbour lookup are integer/index work and must be bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_T, TOL_R = 1e-6, 1e-7
DBL_MAX = np.finfo(np.float64).max


def gpu_map_from_oracle(kb, ctx, om):
    keys, counts, pts = om.export_voxels()
    gm = kb.VoxelHashMap(ctx, om.voxel_size, om.max_distance, om.max_points_per_voxel)
    gm.load_voxels(keys, counts, pts)
    return gm


def sorted_voxels(keys, counts, pts):
    off = np.concatenate([[0], np.cumsum(counts)])
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return keys[order], counts[order], np.concatenate([pts[off[i]:off[i + 1]] for i in order]) if len(order) else pts


def test_map_addpoints_bit_exact(oracle, gpu_ctx):
    """AddPoints on the device reproduces the greedy, input-order dependent CPU result exactly."""
    import kinematic_icp_b200 as kb
    ko = oracle
    rng = np.random.default_rng(11)
    om = ko.OracleMap(1.0, 100.0, 20)
    gm = kb.VoxelHashMap(gpu_ctx, 1.0, 100.0, 20)
    assert gm.Empty()
    for it in range(6):
        # clustered points: many per voxel, negative coordinates, repeated inserts
        pts = rng.normal(size=(6000, 3)) * [6.0, 6.0, 1.5] + [-2.0, 1.0, 0.0]
        if it == 4:  # a dense blob: hundreds of candidates per voxel (the long pending lists of k_add_commit)
            pts = np.concatenate([pts, rng.uniform(-1.0, 1.0, size=(3000, 3)) + [20.0, -7.0, 0.5]])
        om.add_points(pts)
        gm.AddPoints(pts)
        assert gm.num_points() == om.num_points() and gm.num_voxels() == om.num_voxels()
    k1, c1, p1 = sorted_voxels(*gm.export_voxels())
    k0, c0, p0 = om.export_voxels()
    assert np.array_equal(k1, k0) and np.array_equal(c1, c0) and np.array_equal(p1, p0)
    # Pointcloud(): every stored point exactly once (packed on the device, block order)
    cloud = gm.Pointcloud()
    assert not gm.Empty() and cloud.shape == (om.num_points(), 3)
    rows = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(rows(cloud), rows(p0))
    gm.Clear()
    assert gm.Empty() and gm.num_points() == 0


def test_map_update_and_eviction_bit_exact(oracle, gpu_ctx):
    """Update(points, origin) / RemovePointsFarFromLocation: first-point rule, >= max_distance."""
    import kinematic_icp_b200 as kb
    ko = oracle
    rng = np.random.default_rng(12)
    om = ko.OracleMap(0.5, 12.0, 8)
    gm = kb.VoxelHashMap(gpu_ctx, 0.5, 12.0, 8)
    for it in range(8):
        origin = np.array([3.0 * it, 0.5 * it, 0.0])
        pts = rng.uniform(-10, 10, size=(5000, 3)) * [1, 1, 0.2] + origin
        om.update_origin(pts, origin)
        gm.Update(pts, origin)
        assert gm.num_points() == om.num_points() and gm.num_voxels() == om.num_voxels()
    k1, c1, p1 = sorted_voxels(*gm.export_voxels())
    k0, c0, p0 = om.export_voxels()
    assert np.array_equal(k1, k0) and np.array_equal(c1, c0) and np.array_equal(p1, p0)


def test_map_update_pose_bit_exact(oracle, gpu_ctx, workload):
    """Update(points, pose) with the pose applied on the device (Sophus quaternion formula, no FMA contraction)."""
    import kinematic_icp_b200 as kb
    ko = oracle
    w = workload(1)
    om = ko.OracleMap(1.0, 100.0, 20)
    gm = kb.VoxelHashMap(gpu_ctx, 1.0, 100.0, 20)
    pose = ko.planar_pose(50.0, 0.0, 1.2)
    for k in range(3):
        om.update_pose(w.scan, pose)
        gm.Update(w.scan, pose)
        pose = ko.se3_compose(pose, ko.se3_exp([0.8, 0.05, 0, 0, 0, 0.03]))
    k1, c1, p1 = sorted_voxels(*gm.export_voxels())
    k0, c0, p0 = om.export_voxels()
    assert np.array_equal(k1, k0) and np.array_equal(c1, c0) and np.array_equal(p1, p0)


def test_nearest_neighbour_bit_exact(oracle, gpu_ctx, workload):
    """GetClosestNeighbor: same point and same distance, bit for bit, including the empty-neighbourhood case."""
    import kinematic_icp_b200 as kb
    ko = oracle
    w = workload(2)
    gm = gpu_map_from_oracle(kb, gpu_ctx, w.map)
    q = ko.se3_transform(w.prior, w.scan)
    rng = np.random.default_rng(3)
    q = np.concatenate([q, rng.uniform(-120, 120, size=(4000, 3)), np.floor(q[:2000]) + 0.0, -np.abs(q[:500])])
    pg, dg = gm.GetClosestNeighbor(q)
    po, do = w.map.nearest(q)
    assert np.array_equal(dg, do) and np.array_equal(pg, po)
    assert (do == DBL_MAX).sum() > 0  # the case is covered


def check_registration(ko, kb, ctx, om, gm, scan, last, odom, tau, ref_pose=None, **kw):
    reg = kb.KinematicRegistration(kw.get("max_iter", 10), kw.get("conv", 1e-3), 1, kw.get("adaptive", True),
                                   kw.get("fixed_reg", 0.0))
    pose = reg.ComputeRobotMotion(scan, gm, last, odom, tau)
    po, st = om.register(scan, last, odom, tau, max_iter=kw.get("max_iter", 10), conv=kw.get("conv", 1e-3),
                         adaptive=kw.get("adaptive", True), fixed_reg=kw.get("fixed_reg", 0.0))
    res = reg.last_result
    dt, ang = ko.pose_delta(pose, po)
    assert dt <= TOL_T and ang <= TOL_R, (dt, ang)
    assert res.iterations == st.iterations
    # N per iteration is an integer count of accepted correspondences: any NN or gate flip would show here
    assert np.array_equal(res.sums_np()[:, 5], st.sums_np()[:, 5])
    assert np.allclose(res.sums_np()[:, :5], st.sums_np()[:, :5], rtol=1e-9, atol=1e-9)
    assert res.beta == pytest.approx(st.beta, rel=1e-10)
    if ref_pose is not None:
        dt, ang = ko.pose_delta(pose, ref_pose)
        assert dt <= TOL_T and ang <= TOL_R, (dt, ang)
    return pose, dt, ang


@pytest.mark.parametrize("name", ["reg_cfg1", "reg_cfg2_small"])
def test_registration_vs_reference_golden(oracle, gpu_ctx, name):
    """Final pose against the pose the reference's own Registration.cpp produced (tests/golden)."""
    import kinematic_icp_b200 as kb
    ko = oracle
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    om = ko.OracleMap(float(z["voxel_size"]), float(z["max_range"]), int(z["max_points_per_voxel"]))
    om.add_points(z["map_points"])
    gm = kb.VoxelHashMap(gpu_ctx, float(z["voxel_size"]), float(z["max_range"]), int(z["max_points_per_voxel"]))
    gm.load_voxels(z["map_keys"], z["map_counts"], z["map_points"])
    for case in z["cases"]:
        check_registration(ko, kb, gpu_ctx, om, gm, z["scan"], z["last_pose"], z["rel_odom"], case[4], ref_pose=case[5:],
                           max_iter=int(case[0]), conv=case[1], adaptive=bool(case[2]), fixed_reg=case[3])


@pytest.mark.parametrize("cfg", [1, 2, 3, 4])
def test_registration_configs_vs_oracle(oracle, gpu_ctx, workload, cfg):
    """BASELINE.json configs 1-4 at full size: pose, iteration count, per-iteration N and sums — for every launch shape
    of the registration kernel."""
    import kinematic_icp_b200 as kb
    ko = oracle
    w = workload(cfg)
    gm = gpu_map_from_oracle(kb, gpu_ctx, w.map)
    try:
        # launch shapes of the one registration kernel: persistent cooperative launch (default) / one launch per iteration,
        # grid capped at 1 CTA per SM / occupancy limit; the work counters on
        for persistent, ctas, stats, cache in ((1, 0, 0, 2), (1, 0, 0, 0), (0, 0, 0, 2), (1, 1, 1, 2), (0, 1, 0, 0), (1, 0, 0, 1)):
            gpu_ctx.set_option("persistent", persistent)
            gpu_ctx.set_option("ctas_per_sm", ctas)
            gpu_ctx.set_option("stats", stats)
            gpu_ctx.set_option("nn_cache", cache)  # neighbour certificates carried between passes: 2 = always, 0 = never, 1 = by scan size
            pose, dt, ang = check_registration(ko, kb, gpu_ctx, w.map, gm, w.scan, w.last_pose, w.rel_odom, w.tau)
            print("cfg%d persistent=%d ctas_per_sm=%d nn_cache=%d N=%d M=%d pose delta %.3e m %.3e rad" %
                  (cfg, persistent, ctas, cache, w.N, w.map.num_points(), dt, ang))
            if stats:
                probes, cands, lines = gpu_ctx.last_stats()[:3]
                assert probes >= w.N and cands > 0 and lines > 0
    finally:
        gpu_ctx.set_option("persistent", 1)
        gpu_ctx.set_option("ctas_per_sm", 0)
        gpu_ctx.set_option("stats", 0)
        gpu_ctx.set_option("nn_cache", 1)
    pose, dt, ang = check_registration(ko, kb, gpu_ctx, w.map, gm, w.scan, w.last_pose, w.rel_odom, w.tau)
    print("cfg%d defaults pose delta %.3e m %.3e rad" % (cfg, dt, ang))
    # float32 ingest (the reference's callers hold float32 PointCloud2 fields, RosUtils.cpp:30-39): the workload's scan is
    # float32-representable, so the float32 upload must give the very same result
    scan32 = w.scan.astype(np.float32)
    assert np.array_equal(scan32.astype(np.float64), w.scan)
    reg = kb.KinematicRegistration()
    pose32 = reg.ComputeRobotMotion(scan32, gm, w.last_pose, w.rel_odom, w.tau)
    dt32, ang32 = ko.pose_delta(pose32, pose)
    assert dt32 <= 1e-12 and ang32 <= 1e-12, (dt32, ang32)
    gm.close()


def test_registration_edge_cases(oracle, gpu_ctx, workload):
    import kinematic_icp_b200 as kb
    ko = oracle
    w = workload(1)
    # empty map: the prediction, no iterations (Registration.cpp:157)
    gm = kb.VoxelHashMap(gpu_ctx, 1.0, 100.0, 20)
    reg = kb.KinematicRegistration()
    pose = reg.ComputeRobotMotion(w.scan, gm, w.last_pose, w.rel_odom, w.tau)
    dt, ang = ko.pose_delta(pose, ko.se3_compose(w.last_pose, w.rel_odom))
    assert dt < 1e-15 and ang < 1e-15 and reg.last_result.iterations == 0
    gm.close()
    gm = gpu_map_from_oracle(kb, gpu_ctx, w.map)
    # ragged sizes around the 32-point window and the empty scan
    for n in (1, 31, 32, 33, 1000):
        check_registration(ko, kb, gpu_ctx, w.map, gm, w.scan[:n], w.last_pose, w.rel_odom, w.tau)
    # no correspondences: NaN pose as in the reference, plus a status
    reg = kb.KinematicRegistration()
    pose = reg.ComputeRobotMotion(w.scan + 500.0, gm, w.last_pose, w.rel_odom, w.tau)
    assert np.all(np.isnan(pose)) and reg.last_result.status == kb.KICP_WARN_NO_CORRESPONDENCES
    pose = reg.ComputeRobotMotion(np.zeros((0, 3)), gm, w.last_pose, w.rel_odom, w.tau)
    assert np.all(np.isnan(pose))
    # strict gate and max_iter = 1, tiny tau, fixed regularisation
    check_registration(ko, kb, gpu_ctx, w.map, gm, w.scan, w.last_pose, w.rel_odom, 0.3, max_iter=1)
    check_registration(ko, kb, gpu_ctx, w.map, gm, w.scan, w.last_pose, w.rel_odom, w.tau, adaptive=False, fixed_reg=2.0)
    check_registration(ko, kb, gpu_ctx, w.map, gm, w.scan, w.last_pose, w.rel_odom, w.tau, conv=1e-6, max_iter=40)
    # points not representable in float32, general 3-D poses
    rng = np.random.default_rng(2)
    scan = w.scan + rng.normal(size=w.scan.shape) * 1e-3
    last = ko.se3_compose(w.last_pose, ko.se3_exp([0, 0, 0, 0.01, -0.02, 0.0]))
    check_registration(ko, kb, gpu_ctx, w.map, gm, scan, last, w.rel_odom, w.tau)
    gm.close()


def test_device_resident_async_path(oracle, gpu_ctx, workload):
    """kicp_register_scan_async: scan already in HBM, several registrations enqueued, one synchronisation."""
    import kinematic_icp_b200 as kb
    ko = oracle
    w = workload(2)
    gm = gpu_map_from_oracle(kb, gpu_ctx, w.map)
    scan = kb.Scan(gpu_ctx, w.N)
    scan.upload(w.scan)
    reg = kb.KinematicRegistration()
    results = [kb.pinned_result() for _ in range(3)]
    for r in results:
        reg.enqueue(scan, gm, w.last_pose, w.rel_odom, w.tau, r)
    gpu_ctx.synchronize()
    po, st = w.map.register(w.scan, w.last_pose, w.rel_odom, w.tau)
    for r in results:
        dt, ang = ko.pose_delta(r.pose_np(), po)
        assert dt <= TOL_T and ang <= TOL_R and r.iterations == st.iterations
    assert gpu_ctx.launch_count > 0
    scan.close()
    gm.close()


def test_map_storage_growth_and_reserve_bit_exact(oracle, gpu_ctx, monkeypatch):
    """The map's single device slab doubles when a batch could exceed it (contents migrate, table rebuilt) and
    kicp_map_reserve pre-sizes it: contents stay identical to the CPU map through growth, eviction and further inserts."""
    import kinematic_icp_b200 as kb
    ko = oracle
    monkeypatch.setenv("KICP_MAP_VOXELS", "16384")  # smallest initial capacity
    rng = np.random.default_rng(77)
    om = ko.OracleMap(1.0, 60.0, 5)
    gm = kb.VoxelHashMap(gpu_ctx, 1.0, 60.0, 5)
    monkeypatch.delenv("KICP_MAP_VOXELS")
    for it in range(5):
        origin = np.array([15.0 * it, -4.0 * it, 0.0])
        pts = rng.uniform(-70, 70, size=(30000, 3)) * [1, 1, 0.1] + origin  # almost every point opens its own voxel
        om.update_origin(pts, origin)
        gm.Update(pts, origin)
        if it == 2:
            gm.reserve(400000)  # explicit re-size in the middle of a drive
        assert gm.num_points() == om.num_points() and gm.num_voxels() == om.num_voxels()
    assert om.num_voxels() > 16384 * 2
    k1, c1, p1 = sorted_voxels(*gm.export_voxels())
    k0, c0, p0 = sorted_voxels(*om.export_voxels())
    assert np.array_equal(k1, k0) and np.array_equal(c1, c0) and np.array_equal(p1, p0)
    q = rng.uniform(-60, 120, size=(4000, 3)) * [1, 1, 0.1]
    gp, gd = gm.GetClosestNeighbor(q)
    op, od = om.nearest(q)
    assert np.array_equal(gp, op) and np.array_equal(gd, od)
