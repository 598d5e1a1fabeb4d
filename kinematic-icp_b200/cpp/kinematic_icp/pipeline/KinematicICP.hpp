// kinematic_icp::pipeline::{Config, KinematicICP} — the reference's pipeline/KinematicICP.hpp:38-111, member for member,
// so that ros/src/kinematic_icp_ros/server/LidarOdometryServer.cpp:72-105,205-206 compiles unchanged.
#pragma once
#include <Eigen/Core>
#include <cmath>
#include <kiss_icp/core/Preprocessing.hpp>
#include <kiss_icp/core/VoxelHashMap.hpp>
#include <sophus/se3.hpp>
#include <tuple>
#include <vector>

#include "kinematic_icp/correspondence_threshold/CorrespondenceThreshold.hpp"
#include "kinematic_icp/registration/Registration.hpp"

namespace kinematic_icp::pipeline {

struct Config {
    double max_range = 100.0;
    double min_range = 0.0;
    double voxel_size = 1.0;
    unsigned int max_points_per_voxel = 20;
    constexpr double map_resolution() const { return voxel_size / std::sqrt(max_points_per_voxel); }
    bool use_adaptive_threshold = true;
    double fixed_threshold = 1.0;
    int max_num_iterations = 10;
    double convergence_criterion = 0.001;
    int max_num_threads = 1;
    bool use_adaptive_odometry_regularization = true;
    double fixed_regularization = 0.0;
    bool deskew = false;
};

class KinematicICP {
public:
    using Vector3dVector = std::vector<Eigen::Vector3d>;
    using Vector3dVectorTuple = std::tuple<Vector3dVector, Vector3dVector>;

    explicit KinematicICP(const Config &config)
        : registration_(config.max_num_iterations, config.convergence_criterion, config.max_num_threads,
                        config.use_adaptive_odometry_regularization, config.fixed_regularization),
          correspondence_threshold_(config.map_resolution(), config.max_range, config.use_adaptive_threshold, config.fixed_threshold),
          config_(config),
          preprocessor_(config.max_range, config.min_range, config.deskew, config.max_num_threads),
          local_map_(config.voxel_size, config.max_range, config.max_points_per_voxel) {}

    Vector3dVectorTuple RegisterFrame(const std::vector<Eigen::Vector3d> &frame, const std::vector<double> &timestamps,
                                      const Sophus::SE3d &lidar_to_base, const Sophus::SE3d &relative_odometry);
    // Extension (not in the reference): the same frame straight from a PointCloud2-shaped buffer (float32 or float64 fields
    // at a byte stride, include/kicp.h kicp_frame_input), skipping the host-side widening of RosUtils.cpp:30-39.
    Vector3dVectorTuple RegisterFrame(const kicp_frame_input &input, const Sophus::SE3d &lidar_to_base,
                                      const Sophus::SE3d &relative_odometry);

    inline void SetPose(const Sophus::SE3d &pose) {
        last_pose_ = pose;
        local_map_.Clear();
        correspondence_threshold_.Reset();
    };

    std::vector<Eigen::Vector3d> LocalMap() const { return local_map_.Pointcloud(); };
    const kiss_icp::VoxelHashMap &VoxelMap() const { return local_map_; };
    kiss_icp::VoxelHashMap &VoxelMap() { return local_map_; };
    const Sophus::SE3d &pose() const { return last_pose_; }
    Sophus::SE3d &pose() { return last_pose_; }

protected:
    Sophus::SE3d last_pose_;
    KinematicRegistration registration_;
    CorrespondenceThreshold correspondence_threshold_;
    Config config_;
    kiss_icp::Preprocessor preprocessor_;
    kiss_icp::VoxelHashMap local_map_;
};

}  // namespace kinematic_icp::pipeline
