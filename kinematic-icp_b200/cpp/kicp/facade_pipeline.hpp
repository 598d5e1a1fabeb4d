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

// Field order, names and defaults exactly as pipeline/KinematicICP.hpp:38-60 — callers aggregate-initialise this struct
// and assign its fields by name (ros/src/kinematic_icp_ros/server/LidarOdometryServer.cpp:72-97); the layout is part of
// the interface (tests/test_facade_layout_cpu.py compiles one TU against both headers).
struct Config {
    // Preprocessing
    double max_range = 100.0;
    double min_range = 0.0;
    // Mapping parameters
    double voxel_size = 1.0;
    unsigned int max_points_per_voxel = 20;
    // Derived parameter, will be computed from other parts of the configuration
    constexpr double map_resolution() const { return voxel_size / std::sqrt(max_points_per_voxel); }
    // Correspondence threshold parameters
    bool use_adaptive_threshold = true;
    double fixed_threshold = 1.0;  // <-- Ignored if use_adaptive_threshold = true

    // Registration Parameters
    int max_num_iterations = 10;
    double convergence_criterion = 0.001;
    int max_num_threads = 1;  // TBB width in the reference; accepted and ignored on the device
    bool use_adaptive_odometry_regularization = true;
    double fixed_regularization = 0.0;  // <-- Ignored if use_adaptive_odometry_regularization = true

    // Motion compensation
    bool deskew = false;
};

class KinematicICP {
public:
    using Vector3dVector = std::vector<Eigen::Vector3d>;
    using Vector3dVectorTuple = std::tuple<Vector3dVector, Vector3dVector>;

    explicit KinematicICP(const Config &config);

    // {preprocessed frame in the base frame, registration source}; advances pose() and the local map
    Vector3dVectorTuple RegisterFrame(const std::vector<Eigen::Vector3d> &frame,
                                      const std::vector<double> &timestamps,
                                      const Sophus::SE3d &lidar_to_base,
                                      const Sophus::SE3d &relative_odometry);
    // Extension (not in the reference): the same frame straight from a PointCloud2-shaped buffer (float32 or float64 fields
    // at a byte stride, include/kicp.h kicp_frame_input), skipping the host-side widening of RosUtils.cpp:30-39.
    Vector3dVectorTuple RegisterFrame(const kicp_frame_input &input, const Sophus::SE3d &lidar_to_base,
                                      const Sophus::SE3d &relative_odometry);

    void SetPose(const Sophus::SE3d &pose);  // also clears the map and the threshold statistics (KinematicICP.hpp:85-89)

    std::vector<Eigen::Vector3d> LocalMap() const { return local_map_.Pointcloud(); };

    const kiss_icp::VoxelHashMap &VoxelMap() const { return local_map_; };
    kiss_icp::VoxelHashMap &VoxelMap() { return local_map_; };

    const Sophus::SE3d &pose() const { return last_pose_; }
    Sophus::SE3d &pose() { return last_pose_; }

protected:  // same members, names and order as pipeline/KinematicICP.hpp:101-108 (subclasses reach them by name)
    Sophus::SE3d last_pose_;
    // Kinematic module
    KinematicRegistration registration_;
    CorrespondenceThreshold correspondence_threshold_;
    Config config_;
    // KISS-ICP pipeline modules
    kiss_icp::Preprocessor preprocessor_;
    kiss_icp::VoxelHashMap local_map_;  // HBM-resident
};

}  // namespace kinematic_icp::pipeline
