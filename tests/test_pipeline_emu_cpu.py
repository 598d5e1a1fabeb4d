"""The WHOLE frame pipeline on emulated kernels, without a GPU: every frame of the golden drive goes through the device kernels of the
product — Preprocess (de-skew, range filter, base transform), the two VoxelDownsamples, the persistent registration kernel, the map
update with eviction — compiled for the host against the SIMT emulator (tests/emu), chained the way KinematicICP::RegisterFrame chains
them (pipeline/KinematicICP.cpp:48-85; the two CorrespondenceThreshold scalars come from the oracle's restatement, as they come from
the facade's host code in the product).  The trajectory must be the one the reference's own pipeline sources produced
(tests/golden/pipeline_seq.npz), the map must have the same size after every frame.  Test infrastructure."""
import os

import numpy as np
import pytest

from emu import harness as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_T, TOL_R = 1e-6, 1e-7


@pytest.mark.parametrize("deskew", [False, True])
def test_golden_drive_through_emulated_kernels(oracle, deskew):
    from oracle import sequences as S
    ko = oracle
    z = np.load(os.path.join(GOLDEN, "pipeline_seq.npz"))
    seq = S.unpack_sequence(z, deskew)
    tag = "deskew" if deskew else "plain"
    voxel_size, cap, max_range, min_range = 1.0, 20, 100.0, 0.0   # pipeline::Config defaults (KinematicICP.hpp:38-60)
    local_map = H.EmuMap(voxel_size, max_range, cap)
    threshold = ko.OracleThreshold(voxel_size / np.sqrt(cap), max_range, True, 1.0)
    l2b = seq["lidar_to_base"]
    last = np.array(seq["start"], dtype=np.float64)
    for k, (frame, stamps, odom) in enumerate(zip(seq["frames"], seq["stamps"], seq["odoms"])):
        # de-skew in the lidar frame, then the frame in the base frame (:54-59)
        odom_in_lidar = ko.se3_compose(ko.se3_compose(ko.se3_inverse(l2b), odom), l2b)
        in_base = H.preprocess(ko, frame, stamps, odom_in_lidar, max_range, min_range, deskew, lidar_to_base=l2b)
        # Voxelize (:38-44)
        frame_downsample = H.downsample(in_base, voxel_size * 0.5)
        source = H.downsample(frame_downsample, voxel_size * 1.5)
        assert len(source) == z[tag + "_n_src"][k]
        tau = threshold.compute()
        if local_map.num_voxels() == 0:  # an empty map returns the prediction (Registration.cpp:157)
            new_pose = ko.se3_compose(last, odom)
        else:
            res, _ = H.register(local_map.export_voxels(), source, last, odom, tau, grid=3, voxel_size=voxel_size, cap=cap)
            assert res[0].status == 0
            new_pose = res[0].pose_np()
        # threshold model, map, pose (:75-80); the map update as the frame path issues it: count and pose read by the kernels
        threshold.update(ko.se3_compose(ko.se3_inverse(ko.se3_compose(last, odom)), new_pose))
        local_map.update_pose_async(frame_downsample, len(frame_downsample), new_pose)
        last = new_pose
        dt, ang = ko.pose_delta(new_pose, z[tag + "_poses"][k])
        assert dt <= TOL_T and ang <= TOL_R, (k, dt, ang)
        assert local_map.num_points() == z[tag + "_n_map"][k]
    local_map.close()
