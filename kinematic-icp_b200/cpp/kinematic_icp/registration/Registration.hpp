// kinematic_icp::KinematicRegistration — the reference's registration/Registration.hpp:32-50, member for member.
// ComputeRobotMotion runs on the GPU (kicp_register, include/kicp.h); max_num_threads_ is kept for source
// compatibility and ignored.
#pragma once
#include <Eigen/Core>
#include <kiss_icp/core/VoxelHashMap.hpp>
#include <sophus/se3.hpp>
#include <vector>

namespace kinematic_icp {

struct KinematicRegistration {
    explicit KinematicRegistration(const int max_num_iteration, const double convergence_criterion, const int max_num_threads,
                                   const bool use_adaptive_odometry_regularization, const double fixed_regularization);

    Sophus::SE3d ComputeRobotMotion(const std::vector<Eigen::Vector3d> &frame, const kiss_icp::VoxelHashMap &voxel_map,
                                    const Sophus::SE3d &last_robot_pose, const Sophus::SE3d &relative_wheel_odometry,
                                    const double max_correspondence_distance);

    int max_num_iterations_;
    double convergence_criterion_;
    int max_num_threads_;
    bool use_adaptive_odometry_regularization_;
    double fixed_regularization_;
};
}  // namespace kinematic_icp
