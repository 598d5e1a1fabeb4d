"""The known gap of the third-party half, measured instead of guessed: kiss_icp::VoxelDownsample returns its points in the iteration
order of a tsl::robin_map, this repo (oracle, device code) in the order of first occurrence.  The oracle can emit the library's order AS
RECALLED (kicp_oracle.hpp, SetDownsampleOrder; unpinned — neither library is available offline).  Two things are checked here:
the recalled robin-hood table against an independent restatement, and how far that order moves a whole trajectory — which is NOT
"to rounding": the second down-sample keeps the first point of every 1.5-voxel in the order the first one emitted, so another order
means other source points and, through the greedy map insert, another map."""
import ctypes as C
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def robin_order_py(keys, reserve_count, mask20):
    """tsl::robin_map as recalled: power-of-two buckets >= 2 * reserve, robin-hood insertion (richer stays on equal distance)."""
    want = max(2 * reserve_count, 2 * len(keys))
    nb = 1
    while nb < want:
        nb <<= 1
    dist = [-1] * nb
    idx = [0] * nb
    for i, (x, y, z) in enumerate(keys):
        h = ((int(x) & 0xFFFFFFFF) * 73856093 & 0xFFFFFFFF) ^ ((int(y) & 0xFFFFFFFF) * 19349669 & 0xFFFFFFFF) ^ ((int(z) & 0xFFFFFFFF) * 83492791 & 0xFFFFFFFF)
        if mask20:
            h &= (1 << 20) - 1
        ib, d = h & (nb - 1), 0
        while d <= dist[ib]:
            ib, d = (ib + 1) & (nb - 1), d + 1
        cd, ci = d, i
        while dist[ib] >= 0:
            if cd > dist[ib]:
                (cd, ci), (dist[ib], idx[ib]) = (dist[ib], idx[ib]), (cd, ci)
            ib, cd = (ib + 1) & (nb - 1), cd + 1
        dist[ib], idx[ib] = cd, ci
    return [idx[b] for b in range(nb) if dist[b] >= 0]


def test_recalled_robin_map_order_against_independent_restatement(oracle):
    ko = oracle
    L = ko.lib()
    L.kor_robin_order.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(1, 400))
        span = int(rng.choice([3, 8, 40, 1000]))  # small spans: many colliding home buckets
        keys = np.unique(rng.integers(-span, span, size=(n, 3)).astype(np.int32), axis=0)
        keys = np.ascontiguousarray(keys[rng.permutation(len(keys))])
        reserve = int(rng.choice([0, len(keys), 3 * len(keys)]))
        for mask20 in (0, 1):
            out = np.zeros(len(keys), dtype=np.int64)
            L.kor_robin_order(keys.ctypes.data, len(keys), reserve, mask20, out.ctypes.data)
            assert sorted(out.tolist()) == list(range(len(keys)))  # a permutation
            assert out.tolist() == robin_order_py(keys.tolist(), reserve, mask20)


def test_downsample_order_modes_keep_the_same_points(oracle, workload):
    ko = oracle
    w = workload(2)
    rows = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    try:
        base = ko.voxel_downsample(w.scan, 0.5)
        for mode in (1, 2):
            ko.set_downsample_order(mode)
            other = ko.voxel_downsample(w.scan, 0.5)
            assert other.shape == base.shape and np.array_equal(rows(other), rows(base)) and not np.array_equal(other, base)
    finally:
        ko.set_downsample_order(0)
    assert np.array_equal(ko.voxel_downsample(w.scan, 0.5), base)


@pytest.mark.skipif(not os.path.isdir("/root/reference/cpp/kinematic_icp"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("deskew", [False, True])
def test_trajectory_sensitivity_to_the_downsample_order(oracle, deskew):
    """The reference's own pipeline sources over the restated KISS-ICP, golden drive, with the down-sample emitting the recalled
    library order instead of the order of first occurrence: same algorithm, same frames — and a trajectory that differs at the
    centimetre level (measured here: 6 mm / 22 mm, 4 mrad).  This is the size of the third-party gap a real KISS-ICP build would show
    against this repo's pipeline (and against any build of the library with another hash-map iteration order); the registration hot
    path itself — same source cloud, same map in, same pose out — does not depend on it."""
    from oracle import sequences as S
    ko = oracle
    z = np.load(os.path.join(GOLDEN, "pipeline_seq.npz"))
    seq = S.unpack_sequence(z, deskew)
    poses = {}
    try:
        for mode in (0, 1):
            ko.set_downsample_order(mode)
            pipe = ko.ref_pipeline(max_num_threads=1, deskew=deskew)
            poses[mode], n_src, _ = S.run_pipeline(pipe, seq)
            pipe.close()
    finally:
        ko.set_downsample_order(0)
    assert np.array_equal(poses[0], z["deskew_poses" if deskew else "plain_poses"])  # mode 0 is the committed golden
    d = [ko.pose_delta(a, b) for a, b in zip(poses[1], poses[0])]
    dt, da = max(x[0] for x in d), max(x[1] for x in d)
    print("down-sample order gap on the golden drive (deskew=%s): %.3e m, %.3e rad" % (deskew, dt, da))
    assert 1e-6 < dt < 0.1 and da < 0.02  # far above rounding, far below a different scene
