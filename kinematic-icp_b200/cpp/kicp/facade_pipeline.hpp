// kinematic_icp::pipeline::{Config, KinematicICP} of this repo, reached through a forwarding header at the reference's
// include path kinematic_icp/pipeline/KinematicICP.hpp.  The public surface (field names and defaults of Config, the
// constructor, RegisterFrame, SetPose, LocalMap, VoxelMap, pose) is the one ros/src/kinematic_icp_ros/server/
// LidarOdometryServer.cpp:72-105,205-206 uses (reference: pipeline/KinematicICP.hpp:38-111); RegisterFrame itself is one
// device-resident call, kicp_register_frame (include/kicp.h).  Definitions: facade.cpp.
#pragma once
#include <Eigen/Core>
#include <cmath>
#include <sophus/se3.hpp>
#include <tuple>
#include <vector>

#include "kicp/facade_core.hpp"
#include "kiss_icp/core/Preprocessing.hpp"
#include "kiss_icp/core/VoxelHashMap.hpp"

namespace kinematic_icp::pipeline {

struct Config {
    double max_range = 100.0, min_range = 0.0;                     // sensor gate [m]
    double voxel_size = 1.0;                                       // local map
    unsigned int max_points_per_voxel = 20;
    bool use_adaptive_threshold = true;                            // correspondence threshold
    double fixed_threshold = 1.0;
    int max_num_iterations = 10, max_num_threads = 1;              // solver (the thread count is ignored on the device)
    double convergence_criterion = 0.001, fixed_regularization = 0.0;
    bool use_adaptive_odometry_regularization = true;
    bool deskew = false;                                           // motion compensation

    // expected spacing of the points a voxel keeps: the map's discretisation error fed to the threshold model
    constexpr double map_resolution() const { return voxel_size / std::sqrt(max_points_per_voxel); }
};

class KinematicICP {
public:
    using Vector3dVector = std::vector<Eigen::Vector3d>;
    using Vector3dVectorTuple = std::tuple<Vector3dVector, Vector3dVector>;

    explicit KinematicICP(const Config &config);

    // {preprocessed frame in the base frame, registration source}; advances pose() and the local map
    Vector3dVectorTuple RegisterFrame(const std::vector<Eigen::Vector3d> &frame, const std::vector<double> &timestamps,
                                      const Sophus::SE3d &lidar_to_base, const Sophus::SE3d &relative_odometry);
    // Extension (not in the reference): the same frame straight from a PointCloud2-shaped buffer (float32 or float64 fields
    // at a byte stride, include/kicp.h kicp_frame_input), skipping the host-side widening of RosUtils.cpp:30-39.
    Vector3dVectorTuple RegisterFrame(const kicp_frame_input &input, const Sophus::SE3d &lidar_to_base,
                                      const Sophus::SE3d &relative_odometry);

    void SetPose(const Sophus::SE3d &pose);  // also clears the map and the threshold statistics

    std::vector<Eigen::Vector3d> LocalMap() const { return map_.Pointcloud(); }
    kiss_icp::VoxelHashMap &VoxelMap() { return map_; }
    const kiss_icp::VoxelHashMap &VoxelMap() const { return map_; }
    Sophus::SE3d &pose() { return last_pose_; }
    const Sophus::SE3d &pose() const { return last_pose_; }

protected:
    Sophus::SE3d last_pose_;
    KinematicRegistration solver_;
    CorrespondenceThreshold threshold_;
    Config settings_;
    kiss_icp::Preprocessor front_end_;
    kiss_icp::VoxelHashMap map_;  // HBM-resident
};

}  // namespace kinematic_icp::pipeline
