"""The voxel-sorted registration kernel's LOGIC, checked without a GPU: kinematic-icp_b200/csrc/kicp_register_sorted.cu is compiled
unchanged by g++ against a small SIMT emulator (tests/emu/cuda_emu.hpp: one fiber per CUDA thread, warp collectives and
__syncthreads() as rendezvous points, one OS thread per CTA, real atomics between CTAs) and run as a multi-CTA grid on a map laid
out exactly as the device holds it.  Its result must equal the CPU oracle's: pose within the north-star tolerance, the same
number of iterations, the same integer correspondence count in every pass (any neighbour or gate flip would show there).

This is test infrastructure: nothing here is linked into the product, and the GPU parity tests (tests/test_gpu_parity.py, option
"engine") remain the proof for the device build — the emulator cannot see the GPU memory model, launch limits or register hazards."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
GOLDEN = os.path.join(ROOT, "tests", "golden")
TOL_T, TOL_R = 1e-6, 1e-7


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(EMU, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libks_emu.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "kinematic-icp_b200", "csrc"), "-o", so, os.path.join(EMU, "ks_emu.cpp"), "-lpthread"], check=True)
    return C.CDLL(so)


def run_emu(lib, om, scan, last, odom, tau, grid, max_iter=10, conv=1e-3, adaptive=True, fixed_reg=0.0):
    from kinematic_icp_b200 import _capi
    keys, counts, pts = om.export_voxels()
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    f32 = scan.dtype == np.float32
    scan = np.ascontiguousarray(scan)
    n = len(scan)
    last, odom = np.ascontiguousarray(last, dtype=np.float64), np.ascontiguousarray(odom, dtype=np.float64)
    p = _capi.RegParams(max_iter, 1 if adaptive else 0, conv, fixed_reg)
    r = _capi.RegResult()
    so = np.zeros((max(n, 1), 4))
    stats = (C.c_uint64 * 11)()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.ks_emu_register(vp(keys), vp(counts), vp(pts), C.c_int64(len(counts)), C.c_int32(om.max_points_per_voxel), C.c_double(om.voxel_size),
                             vp(scan), C.c_int64(n), C.c_int32(1 if f32 else 0), vp(last), vp(odom), C.c_double(tau), C.byref(p), C.c_int32(grid),
                             C.byref(r), vp(so), stats)
    assert rc == 0, "the launch must leave its scratch clean (table empty, counts and counters zero): rc %d" % rc
    return r, so[:n], list(stats)


def check(lib, ko, om, scan, last, odom, tau, grid, **kw):
    r, so, stats = run_emu(lib, om, scan, last, odom, tau, grid, **kw)
    po, st = om.register(np.asarray(scan, dtype=np.float64), last, odom, tau, max_iter=kw.get("max_iter", 10), conv=kw.get("conv", 1e-3),
                         adaptive=kw.get("adaptive", True), fixed_reg=kw.get("fixed_reg", 0.0))
    dt, ang = ko.pose_delta(r.pose_np(), po)
    assert dt <= TOL_T and ang <= TOL_R, (dt, ang)
    assert r.iterations == st.iterations
    assert np.array_equal(r.sums_np()[:, 5], st.sums_np()[:, 5])
    assert np.allclose(r.sums_np()[:, :5], st.sums_np()[:, :5], rtol=1e-9, atol=1e-9)
    assert r.beta == pytest.approx(st.beta, rel=1e-10)
    # the sort is a permutation of the frame
    rows = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(rows(np.ascontiguousarray(so[:, :3])), rows(np.asarray(scan, dtype=np.float64))) and np.all(so[:, 3] == 0)
    return r, stats


@pytest.mark.parametrize("cfg,grid", [(1, 1), (1, 3), (2, 4)])
def test_sorted_engine_matches_oracle(emu, oracle, workload, cfg, grid):
    ko = oracle
    w = workload(cfg)
    r, stats = check(emu, ko, w.map, w.scan, w.last_pose, w.rel_odom, w.tau, grid)
    assert stats[0] >= w.N * r.iterations and stats[1] > 0
    # float32 ingest: the workload's coordinates are float32-representable, so the very same result
    r32, _ = check(emu, ko, w.map, w.scan.astype(np.float32), w.last_pose, w.rel_odom, w.tau, grid)
    assert np.array_equal(r32.pose_np(), r.pose_np())


def test_sorted_engine_edge_cases(emu, oracle, workload):
    ko = oracle
    w = workload(1)
    for n in (1, 31, 32, 33, 1000):  # ragged sizes around the 32-point chunk
        check(emu, ko, w.map, w.scan[:n], w.last_pose, w.rel_odom, w.tau, 2)
    # no correspondences: NaN pose as in the reference, plus a status; the empty frame
    from kinematic_icp_b200 import _capi
    r, _, _ = run_emu(emu, w.map, w.scan + 500.0, w.last_pose, w.rel_odom, w.tau, 2)
    assert np.all(np.isnan(r.pose_np())) and r.status == _capi.KICP_WARN_NO_CORRESPONDENCES
    r, _, _ = run_emu(emu, w.map, np.zeros((0, 3)), w.last_pose, w.rel_odom, w.tau, 2)
    assert np.all(np.isnan(r.pose_np()))
    # strict gate and max_iter = 1, fixed regularisation, many iterations (the seeded passes)
    check(emu, ko, w.map, w.scan, w.last_pose, w.rel_odom, 0.3, 2, max_iter=1)
    check(emu, ko, w.map, w.scan, w.last_pose, w.rel_odom, w.tau, 2, adaptive=False, fixed_reg=2.0)
    check(emu, ko, w.map, w.scan, w.last_pose, w.rel_odom, w.tau, 3, conv=1e-6, max_iter=40)
    # points not representable in float32, a general 3-D pose
    rng = np.random.default_rng(2)
    scan = w.scan + rng.normal(size=w.scan.shape) * 1e-3
    last = ko.se3_compose(w.last_pose, ko.se3_exp([0, 0, 0, 0.01, -0.02, 0.0]))
    check(emu, ko, w.map, scan, last, w.rel_odom, w.tau, 2)


@pytest.mark.parametrize("name", ["reg_cfg1", "reg_cfg2_small"])
def test_sorted_engine_vs_reference_golden(emu, oracle, name):
    """Final pose against the pose the reference's own Registration.cpp produced (tests/golden)."""
    ko = oracle
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    om = ko.OracleMap(float(z["voxel_size"]), float(z["max_range"]), int(z["max_points_per_voxel"]))
    om.add_points(z["map_points"])
    for case in z["cases"]:
        r, _ = check(emu, ko, om, z["scan"], z["last_pose"], z["rel_odom"], case[4], 3, max_iter=int(case[0]), conv=case[1],
                     adaptive=bool(case[2]), fixed_reg=case[3])
        dt, ang = ko.pose_delta(r.pose_np(), case[5:])
        assert dt <= TOL_T and ang <= TOL_R, (dt, ang)
