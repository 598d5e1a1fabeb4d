"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Synthetic workloads of BASELINE.json `configs` / SURVEY.md §8(d): a deterministic analytic world, a spinning-lidar
scan of the named beam/azimuth counts, and a voxel map built by the CPU oracle's VoxelHashMap::Update along a
trajectory until it holds >= M points, then frozen.  Used by tests/ and bench.py.
"""
import os
import math

import numpy as np

from . import kicp_oracle_py as ko

MAX_POINTS_PER_VOXEL = 20   # pipeline::Config default, cpp/kinematic_icp/pipeline/KinematicICP.hpp:44
MAX_RANGE = 100.0           # KinematicICP.hpp:40


def tau_for(voxel_size):
    """3 (sigma_map + sigma_odom) with sigma_map = Config::map_resolution() (KinematicICP.hpp:46) and
    sigma_odom = 0.1 m, mirroring CorrespondenceThreshold.cpp:54: 0.9708 m at voxel 1.0, 0.6354 m at 0.5."""
    return 3.0 * (voxel_size / math.sqrt(MAX_POINTS_PER_VOXEL) + 0.1)


SCENE = dict(half_extent=95.0, pitch=5.0, wall_h=9.0)  # warehouse: ground, ceiling at 11.7 m, pillar/wall lattice

CONFIGS = {
    # id: scan pattern (beams, elev_min, elev_max, n_az), map points M, voxel size.
    # cfg4 uses voxel_size 0.5 m: at 1.0 m / 20 points per voxel / 100 m range the world saturates near 0.7 M
    # points, so the 1 M-point map BASELINE.json names is reached the way SURVEY.md §8(d) prescribes.
    1: dict(name="cfg1-planar-2k/10k", beams=1, elev=(0.0, 0.0), n_az=2000, M=10_000, voxel_size=1.0,
            mapper=dict(beams=16, elev=(-15.0, 15.0), n_az=900)),
    2: dict(name="cfg2-vlp16-29k/100k", beams=16, elev=(-15.0, 15.0), n_az=1800, M=100_000, voxel_size=1.0),
    3: dict(name="cfg3-64beam-131k/500k", beams=64, elev=(-24.8, 2.0), n_az=2048, M=500_000, voxel_size=1.0,
            mapper=dict(beams=128, elev=(-22.5, 22.5), n_az=2048)),
    4: dict(name="cfg4-os1-128-262k/1M", beams=128, elev=(-22.5, 22.5), n_az=2048, M=1_000_000, voxel_size=0.5),
}
CACHE_DIR = os.environ.get("KICP_WORKLOAD_CACHE", "/tmp/kicp_workloads")


def unicycle(d, theta):
    """SE3 of the unicycle arc (d, theta): the reference's motion_model, Registration.cpp:159-167."""
    eps = np.finfo(np.float64).tiny
    return ko.se3_exp([d * math.sin(theta) / (theta + eps), d * (1.0 - math.cos(theta)) / (theta + eps), 0, 0, 0, theta])


def _yaw_xy(p7):
    yaw = 2.0 * math.atan2(p7[2], p7[3])
    return p7[4], p7[5], yaw


class Workload:
    """scan (N,3), oracle map, last_pose, rel_odom (corrupted), true_pose, tau — one registration problem."""

    def __init__(self, cfg_id, M=None, n_az=None, cache=True):
        cfg = dict(CONFIGS[cfg_id])
        self.cfg_id = cfg_id
        self.name = cfg["name"]
        M = cfg["M"] if M is None else M
        n_az = cfg["n_az"] if n_az is None else n_az
        vs = cfg["voxel_size"]
        self.voxel_size, self.max_range, self.max_points_per_voxel = vs, MAX_RANGE, MAX_POINTS_PER_VOXEL
        self.tau = tau_for(vs)
        # wheel odometry = truth (0.5 m, 0.05 rad) corrupted by (+5 % distance, +0.01 rad)   
        self.rel_odom = unicycle(0.5 * 1.05, 0.05 + 0.01)
        self.map = ko.OracleMap(vs, MAX_RANGE, MAX_POINTS_PER_VOXEL)
        path = os.path.join(CACHE_DIR, "cfg%d_M%d_az%d_v2.npz" % (cfg_id, M, n_az))
        if cache and os.path.exists(path):
            z = np.load(path)
            self.map.add_points(z["map_points"])  # voxel-grouped, insertion order kept => identical map
            self.scan, self.last_pose, self.true_pose = z["scan"], z["last_pose"], z["true_pose"]
            self.mapping_scans = int(z["mapping_scans"])
            assert self.map.num_points() == len(z["map_points"])
            return
        mapper = cfg.get("mapper", dict(beams=cfg["beams"], elev=cfg["elev"], n_az=cfg["n_az"]))
        # trajectory: counter-clockwise circle of radius 50 m around the world origin (a ring corridor the scene
        # keeps free), 1 m per mapping scan
        pose = ko.planar_pose(50.0, 0.0, math.pi / 2)
        step = unicycle(1.0, 1.0 / 50.0)
        k = 0
        last_pose = pose
        while self.map.num_points() < M:
            x, y, yaw = _yaw_xy(pose)
            scan = ko.synth_scan(mapper["beams"], mapper["elev"][0], mapper["elev"][1], mapper["n_az"], x, y, yaw,
                                 seed=1000 + cfg_id + 7919 * (k + 1), **SCENE)
            # the pipeline inserts the 0.5*voxel_size downsample of the frame (KinematicICP.cpp:38-44,79)
            ds = ko.voxel_downsample(scan, 0.5 * vs)
            self.map.update_pose(ds, pose)
            last_pose = pose
            pose = ko.se3_compose(pose, step)
            k += 1
            if k > 320:
                raise RuntimeError("map did not reach %d points (%d)" % (M, self.map.num_points()))
        self.mapping_scans = k
        self.last_pose = last_pose
        self.true_pose = ko.se3_compose(last_pose, unicycle(0.5, 0.05))
        x, y, yaw = _yaw_xy(self.true_pose)
        self.scan = ko.synth_scan(cfg["beams"], cfg["elev"][0], cfg["elev"][1], n_az, x, y, yaw, seed=1000 + cfg_id,
                                  **SCENE)
        if cache:
            os.makedirs(CACHE_DIR, exist_ok=True)
            _, _, pts = self.map.export_voxels()
            tmp = path + ".tmp%d.npz" % os.getpid()
            np.savez(tmp, map_points=pts, scan=self.scan, last_pose=self.last_pose, true_pose=self.true_pose,
                     mapping_scans=self.mapping_scans)
            os.replace(tmp, path)

    @property
    def N(self):
        return len(self.scan)

    @property
    def prior(self):
        return ko.se3_compose(self.last_pose, self.rel_odom)

    def describe(self):
        return dict(name=self.name, N=int(self.N), M=int(self.map.num_points()), voxels=int(self.map.num_voxels()),
                    mapping_scans=int(self.mapping_scans), tau=float(self.tau), voxel_size=self.voxel_size,
                    max_points_per_voxel=MAX_POINTS_PER_VOXEL)
