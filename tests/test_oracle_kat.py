"""Known-answer tests of the CPU oracle (SURVEY.md §4, KAT-1 .. KAT-8).  The reference ships no tests; these pin the
restatement against closed forms of the maths at the cited lines of cpp/kinematic_icp/registration/Registration.cpp."""
import math

import numpy as np
import pytest

DBL_MIN = np.finfo(np.float64).tiny
DBL_MAX = np.finfo(np.float64).max


def unicycle(ko, d, th):
    return ko.se3_exp([d * math.sin(th) / (th + DBL_MIN), d * (1 - math.cos(th)) / (th + DBL_MIN), 0, 0, 0, th])


def test_kat1_motion_model(oracle):
    """Registration.cpp:159-167: exp([d sin th/th, d(1-cos th)/th, 0,0,0, th]) is a planar arc."""
    ko = oracle
    d, th = 0.7, 0.3
    T = unicycle(ko, d, th)
    # closed form of SE3::exp with V(theta): t = V [ux, uy]
    ux, uy = d * math.sin(th) / th, d * (1 - math.cos(th)) / th
    s, c = math.sin(th), math.cos(th)
    tx = (s / th) * ux - ((1 - c) / th) * uy
    ty = ((1 - c) / th) * ux + (s / th) * uy
    assert np.allclose(T[4:], [tx, ty, 0.0], atol=1e-15)
    assert np.allclose(T[:4], [0, 0, math.sin(th / 2), math.cos(th / 2)], atol=1e-15)
    # theta -> 0: a translation of d along x
    T = unicycle(ko, d, 1e-9)
    assert np.allclose(T[4:], [d, 0, 0], atol=1e-8)
    # theta == 0 exactly: sin(0)/(0 + DBL_MIN) = 0 -> identity translation (the epsilon quirk of :46)
    T = unicycle(ko, d, 0.0)
    assert np.array_equal(T, ko.IDENTITY)


def test_se3_exp_log_roundtrip_and_inverse(oracle):
    ko = oracle
    rng = np.random.default_rng(0)
    for _ in range(20):
        xi = rng.normal(size=6) * np.array([2, 2, 2, 0.8, 0.8, 0.8])
        T = ko.se3_exp(xi)
        assert np.allclose(ko.se3_log(T), xi, atol=1e-12)
        I = ko.se3_compose(T, ko.se3_inverse(T))
        dt, ang = ko.pose_delta(I, ko.IDENTITY)
        assert dt < 1e-14 and ang < 1e-14
    assert np.allclose(ko.se3_log(ko.se3_exp(np.zeros(6))), 0.0)


def test_kat2_jacobian_finite_difference(oracle):
    """Registration.cpp:89-91: d/d(d,theta) of T exp(motion(dx)) p at 0 equals [R e_x | R (-p_y, p_x, 0)]."""
    ko = oracle
    T = ko.se3_compose(ko.se3_exp([1.0, -2.0, 0.3, 0.1, -0.2, 0.7]), ko.IDENTITY)
    p = np.array([[3.0, -1.5, 0.8]])
    h = 1e-6

    def f(d, th):
        return ko.se3_transform(ko.se3_compose(T, unicycle(ko, d, th)), p)[0]

    Jd = (f(h, 1e-300) - f(-h, 1e-300)) / (2 * h)
    Jt = (f(0.0, h) - f(0.0, -h)) / (2 * h)
    R_ex = ko.se3_transform(T, np.array([[1.0, 0, 0]]))[0] - T[4:]
    R_c1 = ko.se3_transform(T, np.array([[-p[0, 1], p[0, 0], 0.0]]))[0] - T[4:]
    assert np.allclose(Jd, R_ex, atol=1e-8)
    assert np.allclose(Jt, R_c1, atol=1e-8)


def _tiny_problem(ko):
    """4 map points, 4 scan points, each scan point within reach of exactly one map point."""
    m = ko.OracleMap(1.0, 100.0, 20)
    target = np.array([[2.3, 0.4, 0.5], [-1.6, 3.2, 0.4], [0.5, -2.7, 1.5], [4.4, 4.6, 0.6]])
    m.add_points(target)
    return m, target


def test_kat3_kat4_solve_and_beta(oracle):
    """One solve by hand: the /N and +diag(beta,0) terms (Registration.cpp:119-125) and beta (:48-60)."""
    ko = oracle
    m, target = _tiny_problem(ko)
    T0 = ko.planar_pose(0.05, -0.03, 0.02)
    src = ko.se3_transform(ko.se3_inverse(T0), target) + np.array([[0.02, -0.01, 0.0], [0.0, 0.03, 0.01],
                                                                   [-0.02, 0.0, 0.0], [0.01, 0.01, -0.02]])
    pose, st = m.register(src, T0, ko.IDENTITY, tau=0.5, max_iter=1)
    assert st.iterations == 1
    q = ko.se3_transform(T0, src)
    r = q - target
    R = np.array([ko.se3_transform(T0, np.eye(3))[i] - T0[4:] for i in range(3)]).T  # columns R e_i
    JTJ = np.zeros((2, 2))
    JTr = np.zeros(2)
    for i in range(4):
        J = np.stack([R @ np.array([1.0, 0, 0]), R @ np.array([-src[i, 1], src[i, 0], 0.0])], axis=1)
        JTJ += J.T @ J
        JTr += J.T @ r[i]
    beta = 1.0 / ((r ** 2).sum() / 4 + DBL_MIN)
    assert st.beta == pytest.approx(beta, rel=1e-13)
    A = JTJ / 4 + np.diag([beta, 0.0])
    dx = -np.linalg.inv(A) @ (JTr / 4)
    assert np.allclose(st.dx_np()[0], dx, rtol=1e-10, atol=1e-14)
    s = st.sums_np()[0]
    assert np.allclose(s[:5], [JTJ[0, 0], JTJ[0, 1], JTJ[1, 1], JTr[0], JTr[1]], rtol=1e-12, atol=1e-14)
    assert s[5] == 4 and s[6] == pytest.approx((r ** 2).sum(), rel=1e-13)
    assert np.allclose(pose, ko.se3_compose(T0, unicycle(ko, dx[0], dx[1])), atol=1e-13)
    # fixed regularisation (:171-177)
    _, st2 = m.register(src, T0, ko.IDENTITY, tau=0.5, max_iter=1, adaptive=False, fixed_reg=0.25)
    assert st2.beta == 0.25


def test_kat5_nearest_neighbour_semantics(oracle):
    """GetClosestNeighbor (KISS-ICP v1.2.0): 27-neighbourhood brute force, floor voxelisation, (0, DBL_MAX) when empty."""
    ko = oracle
    rng = np.random.default_rng(5)
    m = ko.OracleMap(1.0, 100.0, 20)
    pts = rng.uniform(-6, 6, size=(4000, 3))
    m.add_points(pts)
    stored = m.pointcloud()
    q = rng.uniform(-5, 5, size=(500, 3))
    nn, d = m.nearest(q)
    for i in range(len(q)):
        # brute force restricted to the 27 voxels around floor(q)
        v = np.floor(q[i])
        sel = np.all(np.abs(np.floor(stored) - v) <= 1, axis=1)
        dist = np.linalg.norm(stored[sel] - q[i], axis=1)
        assert d[i] == pytest.approx(dist.min(), rel=1e-15)
        # whenever the global NN is within one voxel ring it is the one returned
        gd = np.linalg.norm(stored - q[i], axis=1)
        if gd.min() < 1.0 - np.max(np.abs(q[i] - v - 0.5)) + 0.5:
            assert d[i] == pytest.approx(gd.min(), rel=1e-15)
    # negative coordinates use floor, not truncation: a point at x = -0.25 lives in voxel -1
    m2 = ko.OracleMap(1.0, 100.0, 20)
    m2.add_points(np.array([[-0.25, 0.5, 0.5]]))
    keys, counts, _ = m2.export_voxels()
    assert keys.tolist() == [[-1, 0, 0]] and counts.tolist() == [1]
    # farther than one ring: nothing found
    nn, d = m2.nearest(np.array([[2.5, 0.5, 0.5]]))
    assert d[0] == DBL_MAX and np.all(nn[0] == 0.0)
    # a farther point in the ring is returned even if a closer one exists outside it
    m2.add_points(np.array([[4.01, 0.5, 0.5]]))
    nn, d = m2.nearest(np.array([[2.99, 0.5, 0.5]]))  # voxel 2: ring covers voxels 1..3, so (4.01,..) is invisible
    assert d[0] == DBL_MAX


def test_kat6_gate_is_strict(oracle):
    """Registration.cpp:75: distance < max_correspondance_distance (strict)."""
    ko = oracle
    m = ko.OracleMap(1.0, 100.0, 20)
    m.add_points(np.array([[0.5, 0.5, 0.5], [5.5, 0.5, 0.5]]))
    src = np.array([[0.75, 0.5, 0.5], [5.5, 0.75, 0.5]])  # both exactly 0.25 away
    _, st = m.register(src, ko.IDENTITY, ko.IDENTITY, tau=0.25, max_iter=1)
    assert st.sums_np()[0][5] == 0  # d == tau is rejected -> N = 0
    _, st = m.register(src, ko.IDENTITY, ko.IDENTITY, tau=np.nextafter(0.25, 1.0), max_iter=1)
    assert st.sums_np()[0][5] == 2


def test_kat7_recovers_known_motion(oracle, workload):
    """End to end on the planar config: the estimate moves from the corrupted prior towards the truth."""
    ko = oracle
    w = workload(1)
    pose, st = w.map.register(w.scan, w.last_pose, w.rel_odom, w.tau)
    prior = ko.se3_compose(w.last_pose, w.rel_odom)
    assert ko.pose_delta(pose, w.true_pose)[1] < 0.2 * ko.pose_delta(prior, w.true_pose)[1]
    assert 1 <= st.iterations <= 10
    # without regularisation the translation is free to move too
    pose2, _ = w.map.register(w.scan, w.last_pose, w.rel_odom, w.tau, adaptive=False, fixed_reg=0.0, max_iter=30)
    assert ko.pose_delta(pose2, w.true_pose)[0] < ko.pose_delta(prior, w.true_pose)[0]


def test_kat8_empty_map_and_early_break(oracle):
    """Registration.cpp:157 (empty map -> prediction) and :184 (break before re-association)."""
    ko = oracle
    m = ko.OracleMap(1.0, 100.0, 20)
    last, odom = ko.planar_pose(1, 2, 0.3), ko.planar_pose(0.5, 0.0, 0.05)
    pose, st = m.register(np.zeros((5, 3)), last, odom, tau=1.0)
    assert np.array_equal(pose, ko.se3_compose(last, odom)) and st.iterations == 0 and st.associations == 0
    # perfectly aligned input: |dx| < conv at j = 0 -> exactly one association, one solve
    m, target = _tiny_problem(ko)
    pose, st = m.register(target, ko.IDENTITY, ko.IDENTITY, tau=0.5)
    assert st.iterations == 1 and st.associations == 1
    # N == 0 -> NaN pose, undefended (Registration.cpp:119-125)
    pose, st = m.register(target + 50.0, ko.IDENTITY, ko.IDENTITY, tau=0.5)
    assert np.all(np.isnan(pose))


def test_map_insert_rules(oracle):
    """AddPoints: per-voxel cap and min spacing map_resolution = sqrt(vs^2 / cap); RemoveFar uses the FIRST point."""
    ko = oracle
    m = ko.OracleMap(1.0, 10.0, 4)  # map_resolution = 0.5
    m.add_points(np.array([[0.1, 0.1, 0.1], [0.2, 0.1, 0.1], [0.7, 0.1, 0.1], [0.1, 0.7, 0.1], [0.7, 0.7, 0.1],
                           [0.7, 0.7, 0.7], [0.1, 0.1, 0.7]]))
    keys, counts, pts = m.export_voxels()
    assert counts.tolist() == [4]  # second point too close, 6th/7th rejected because the voxel is full
    assert pts.tolist() == [[0.1, 0.1, 0.1], [0.7, 0.1, 0.1], [0.1, 0.7, 0.1], [0.7, 0.7, 0.1]]
    # the spacing test never looks into the neighbouring voxel
    m.add_points(np.array([[1.05, 0.1, 0.1]]))
    assert m.num_points() == 5
    # eviction: first point of the voxel decides (>= max_distance)
    m.remove_far(np.array([10.1, 0.1, 0.1]))
    # voxel 0's first point is exactly 10.0 away (>= evicts all 4 of its points); voxel 1's single point stays
    assert m.num_points() == 1 and m.num_voxels() == 1
    # VoxelDownsample keeps the first point per voxel
    ds = ko.voxel_downsample(np.array([[0.1, 0.1, 0.1], [0.2, 0.2, 0.2], [1.5, 0.1, 0.1], [0.3, 0.3, 0.3]]), 1.0)
    assert sorted(ds.tolist()) == [[0.1, 0.1, 0.1], [1.5, 0.1, 0.1]]


def test_threshold(oracle):
    """CorrespondenceThreshold.cpp:29-64."""
    ko = oracle
    th = ko.OracleThreshold(0.2236, 100.0, True, 1.0)
    assert th.compute() == pytest.approx(3 * 0.2236)  # sqrt(0 / 1e-8) = 0
    err = ko.se3_exp([0.03, 0.0, 0, 0, 0, 0.001])
    th.update(err)
    theta = 0.001
    e = np.linalg.norm(err[4:]) + 2 * 100.0 * math.sin(theta / 2)
    assert th.compute() == pytest.approx(3 * (0.2236 + math.sqrt(e * e / (1 + 1e-8))), rel=1e-12)
    assert ko.OracleThreshold(0.2, 100.0, False, 1.5).compute() == 1.5


def test_threaded_oracle_matches_sequential(oracle, workload):
    ko = oracle
    w = workload(2)
    p1, s1 = w.map.register(w.scan, w.last_pose, w.rel_odom, w.tau, threads=1)
    p4, s4 = w.map.register(w.scan, w.last_pose, w.rel_odom, w.tau, threads=4)
    dt, ang = ko.pose_delta(p1, p4)
    assert dt < 1e-12 and ang < 1e-12 and s1.iterations == s4.iterations
    assert np.array_equal(s1.sums_np()[:, 5], s4.sums_np()[:, 5])


def test_front_end_voxel_downsample_semantics(oracle):
    """kiss_icp::VoxelDownsample (KISS-ICP v1.2.0, call sites pipeline/KinematicICP.cpp:38-44): floor voxelisation, the FIRST
    point of the input order wins its voxel, survivors keep their input order."""
    ko = oracle
    pts = np.array([[0.10, 0.10, 0.10],    # voxel (0,0,0)   kept
                    [0.90, 0.90, 0.90],    # voxel (0,0,0)   dropped (later)
                    [-0.10, 0.10, 0.10],   # voxel (-1,0,0)  kept: floor, not truncation
                    [1.00, 0.00, 0.00],    # voxel (1,0,0)   kept: the upper face belongs to the next voxel
                    [-1.00, 0.0, 0.0],     # voxel (-1,0,0)  dropped
                    [0.999999, 0.5, 0.5]])  # voxel (0,0,0)  dropped
    out = ko.voxel_downsample(pts, 1.0)
    assert np.array_equal(out, pts[[0, 2, 3]])
    # a permutation changes the winners, not the voxel set
    out2 = ko.voxel_downsample(pts[::-1].copy(), 1.0)
    assert np.array_equal(out2, pts[::-1][[0, 1, 2]])
    # the pipeline's double down-sample is idempotent per level and never grows
    rng = np.random.default_rng(8)
    cloud = rng.normal(size=(5000, 3)) * [10, 10, 2]
    a = ko.voxel_downsample(cloud, 0.5)
    b = ko.voxel_downsample(a, 1.5)
    assert np.array_equal(ko.voxel_downsample(a, 0.5), a) and len(b) <= len(a) <= len(cloud)
    keys = np.floor(b / 1.5).astype(np.int64)
    assert len(np.unique(keys, axis=0)) == len(b)  # one survivor per voxel
    assert ko.voxel_downsample(np.zeros((0, 3)), 1.0).shape == (0, 3)


def test_front_end_preprocess_semantics(oracle):
    """kiss_icp::Preprocessor::Preprocess (call site pipeline/KinematicICP.cpp:54-57): strict range gate on both sides;
    de-skew p <- exp((s - 1) log T) p with stamps normalised to [0, 1] (any affine time scale gives the same result)."""
    ko = oracle
    pts = np.array([[0.5, 0, 0], [0.5 + 1e-12, 0, 0], [100.0, 0, 0], [100.0 - 1e-9, 0, 0], [3.0, 4.0, 0.0], [0, 0, 0]])
    out = ko.preprocess(pts, np.zeros(0), ko.IDENTITY, 100.0, 0.5, False)
    assert np.array_equal(out, pts[[1, 3, 4]])  # r == min and r == max are both rejected
    # deskew requested without stamps, or stamps without deskew: the frame is only range-filtered
    T = ko.se3_exp([0.8, 0.05, 0.0, 0.0, 0.0, 0.04])
    assert np.array_equal(ko.preprocess(pts, np.zeros(0), T, 100.0, 0.5, True), out)
    assert np.array_equal(ko.preprocess(pts, np.linspace(0, 1, len(pts)), T, 100.0, 0.5, False), out)
    # closed form at the ends of the sweep: stamp 1 -> identity, stamp 0 -> T^-1
    cloud = np.random.default_rng(9).uniform(-20, 20, size=(64, 3))
    stamps = np.linspace(0.0, 1.0, len(cloud))
    d = ko.preprocess(cloud, stamps, T, 1e9, 0.0, True)
    assert np.allclose(d[-1], cloud[-1], atol=1e-13)
    assert np.allclose(d[0], ko.se3_transform(ko.se3_inverse(T), cloud[:1])[0], atol=1e-12)
    # mid-sweep: exp(-0.5 log T) = the inverse of the half motion
    half = ko.se3_exp(0.5 * np.asarray(ko.se3_log(T)))
    k = len(cloud) // 2
    s = stamps[k]
    mid = ko.se3_exp((s - 1.0) * np.asarray(ko.se3_log(T)))
    assert np.allclose(d[k], ko.se3_transform(mid, cloud[k:k + 1])[0], atol=1e-12)
    assert np.allclose(ko.se3_compose(half, half), T, atol=1e-12)
    # time scale invariance (TimeStampHandler.cpp:129-135 normalises as well; doing it twice is harmless)
    d2 = ko.preprocess(cloud, 1.7e9 + 0.1 * stamps, T, 1e9, 0.0, True)
    assert np.allclose(d2, d, atol=1e-6)  # 0.1 s on 1.7e9 s keeps ~7 digits of the stamp
    d3 = ko.preprocess(cloud, 5.0 + 0.25 * stamps, T, 1e9, 0.0, True)
    assert np.allclose(d3, d, atol=1e-12)


def test_map_invariants_after_random_updates(oracle):
    """Size-independent properties of the restated VoxelHashMap (KISS-ICP v1.2.0 semantics) that the bit-exact
    GPU comparison inherits: per-voxel cap, floor keys, min spacing inside a voxel, first-point eviction rule."""
    ko = oracle
    rng = np.random.default_rng(31)
    vs, max_d, cap = 0.8, 25.0, 6
    m = ko.OracleMap(vs, max_d, cap)
    res = math.sqrt(vs * vs / cap)
    origin = np.zeros(3)
    for it in range(12):
        origin = origin + [2.5, 0.7, 0.0]
        pts = rng.normal(size=(4000, 3)) * [12.0, 12.0, 1.0] + origin
        m.update_origin(pts, origin)
        keys, counts, stored = m.export_voxels()
        assert counts.min() >= 1 and counts.max() <= cap and counts.sum() == len(stored) == m.num_points()
        off = np.concatenate([[0], np.cumsum(counts)])
        vox = np.repeat(keys, counts, axis=0)
        assert np.array_equal(np.floor(stored / vs).astype(np.int64), vox.astype(np.int64))  # every point lies in its voxel
        first = stored[off[:-1]]
        assert (np.linalg.norm(first - origin, axis=1) < max_d).all()  # survivors: FIRST point strictly inside max_distance
        for v in rng.integers(0, len(counts), 200):  # spacing rule inside a voxel: later points keep >= map_resolution
            p = stored[off[v]:off[v + 1]]
            if len(p) > 1:
                d = np.linalg.norm(p[:, None, :] - p[None, :, :], axis=2)[np.triu_indices(len(p), 1)]
                assert (d >= res * (1 - 1e-12)).all()
    assert len(np.unique(keys, axis=0)) == len(keys)  # one block per voxel
