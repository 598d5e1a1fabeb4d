"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.  A short synthetic drive (frames + wheel odometry) for whole-pipeline parity:
kinematic_icp::pipeline::KinematicICP::RegisterFrame called frame after frame, as offline_node does
(ros/src/kinematic_icp_ros/nodes/offline_node.cpp:99-149)."""
import math

import numpy as np

from . import kicp_oracle_py as ko
from . import workloads as W


def make_sequence(n_frames=8, beams=16, n_az=600, seed=4242, deskew=False):
    """Frames in the LIDAR frame, per-point stamps in [0,1], the lidar->base extrinsic, and corrupted odometry deltas."""
    rng = np.random.default_rng(seed)
    lidar_to_base = ko.se3_exp([0.2, 0.0, 0.3, 0.0, 0.0, 0.05])  # small mounting offset and yaw
    base_to_lidar = ko.se3_inverse(lidar_to_base)
    pose = ko.planar_pose(50.0, 0.0, math.pi / 2)
    frames, stamps, odoms, truth = [], [], [], [pose]
    for k in range(n_frames):
        # the robot moves by delta_k, then frame k is taken; RegisterFrame gets the (noisy) wheel odometry of that motion
        d, th = 0.6 + 0.05 * math.sin(k), 0.012 + 0.004 * math.cos(k)
        pose = ko.se3_compose(pose, W.unicycle(d, th))
        truth.append(pose)
        odoms.append(W.unicycle(d * (1.0 + 0.03 * rng.standard_normal()), th + 0.004 * rng.standard_normal()))
        x, y, yaw = W._yaw_xy(pose)
        scan_base = ko.synth_scan(beams, -15.0, 15.0, n_az, x, y, yaw, seed=seed + k, **W.SCENE)
        scan_lidar = ko.se3_transform(base_to_lidar, scan_base)
        scan_lidar = scan_lidar.astype(np.float32).astype(np.float64)  # what a PointCloud2 float32 field carries
        frames.append(scan_lidar)
        stamps.append(np.linspace(0.0, 1.0, len(scan_lidar)) if deskew else np.zeros(0))
    return dict(frames=frames, stamps=stamps, odoms=odoms, lidar_to_base=lidar_to_base, start=truth[0], truth=truth)


def run_pipeline(pipe, seq):
    pipe.set_pose(seq["start"])
    poses, n_src, n_map = [], [], []
    for f, s, o in zip(seq["frames"], seq["stamps"], seq["odoms"]):
        p, n = pipe.register_frame(f, s, seq["lidar_to_base"], o)
        poses.append(p), n_src.append(n), n_map.append(pipe.num_map_points())
    return np.array(poses), np.array(n_src), np.array(n_map)


def pack_sequence(seq):
    """Inputs of a sequence as arrays for an .npz fixture (frames are float32-representable, stored as float32)."""
    lens = np.array([len(f) for f in seq["frames"]])
    return dict(seq_frames=np.concatenate(seq["frames"]).astype(np.float32), seq_lens=lens, seq_odoms=np.array(seq["odoms"]),
                seq_lidar_to_base=seq["lidar_to_base"], seq_start=seq["start"])


def unpack_sequence(z, deskew):
    off = np.concatenate([[0], np.cumsum(z["seq_lens"])])
    frames = [z["seq_frames"][off[i]:off[i + 1]].astype(np.float64) for i in range(len(z["seq_lens"]))]
    stamps = [np.linspace(0.0, 1.0, len(f)) if deskew else np.zeros(0) for f in frames]
    return dict(frames=frames, stamps=stamps, odoms=list(z["seq_odoms"]), lidar_to_base=z["seq_lidar_to_base"], start=z["seq_start"])


def write_kseq(seq, path, header_stamps=None):
    """The sequence as a .kseq file, the input of the product's native replay harness (kinematic-icp_b200/cpp/tools/kicp_replay.cpp
    documents the layout): float32 points as a PointCloud2 message carries them, per-point stamps, the wheel odometry of every frame."""
    import struct
    with open(path, "wb") as f:
        f.write(b"KSEQ1\0\0\0")
        f.write(struct.pack("<ii", len(seq["frames"]), 0))
        f.write(np.asarray(seq["lidar_to_base"], dtype="<f8").tobytes())
        f.write(np.asarray(seq["start"], dtype="<f8").tobytes())
        for k, (fr, st, od) in enumerate(zip(seq["frames"], seq["stamps"], seq["odoms"])):
            has = 1 if len(st) else 0
            f.write(struct.pack("<iid", len(fr), has, float(header_stamps[k]) if header_stamps is not None else 0.1 * k))
            f.write(np.asarray(od, dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(fr, dtype="<f4").tobytes())
            if has:
                f.write(np.ascontiguousarray(st, dtype="<f8").tobytes())
