// Device-backed stand-ins for the two solver-side classes of the reference's core, declared once here and reached
// through forwarding headers at the reference's include paths:
//   kinematic_icp/registration/Registration.hpp                          -> kinematic_icp::KinematicRegistration
//   kinematic_icp/correspondence_threshold/CorrespondenceThreshold.hpp   -> kinematic_icp::CorrespondenceThreshold
// Names, argument order and public data members follow the reference (registration/Registration.hpp:32-50,
// correspondence_threshold/CorrespondenceThreshold.hpp:30-55) because its pipeline and ROS layer address them by name;
// everything behind them is this repo's: ComputeRobotMotion is one kicp_register call (include/kicp.h), the threshold
// model is two host scalars.  Definitions: facade.cpp.
#pragma once
#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <vector>

#include "kiss_icp/core/VoxelHashMap.hpp"

namespace kinematic_icp {

struct KinematicRegistration {
    // ---- knobs (read on every call, so they may be changed between frames)
    int max_num_iterations_;
    double convergence_criterion_;
    int max_num_threads_;  // TBB width in the reference; kept for source compatibility, the device path ignores it
    bool use_adaptive_odometry_regularization_;
    double fixed_regularization_;

    explicit KinematicRegistration(const int max_num_iteration, const double convergence_criterion, const int max_num_threads,
                                   const bool use_adaptive_odometry_regularization, const double fixed_regularization);

    // The hot path: `frame` (robot base frame) against `voxel_map`, starting from last_robot_pose * relative_wheel_odometry,
    // correspondences gated at max_correspondence_distance.  Runs entirely on the GPU.
    Sophus::SE3d ComputeRobotMotion(const std::vector<Eigen::Vector3d> &frame, const kiss_icp::VoxelHashMap &voxel_map,
                                    const Sophus::SE3d &last_robot_pose, const Sophus::SE3d &relative_wheel_odometry,
                                    const double max_correspondence_distance);
};

struct CorrespondenceThreshold {
    // ---- configuration
    double map_discretization_error_;
    double max_range_;
    bool use_adaptive_threshold_;
    double fixed_threshold_;
    // ---- running second moment of the odometry error, in point space
    double odom_sse_;
    double num_samples_;

    explicit CorrespondenceThreshold(const double map_discretization_error, const double max_range, const bool use_adaptive_threshold,
                                     const double fixed_threshold);
    double ComputeThreshold() const;                                // tau handed to ComputeRobotMotion
    void UpdateOdometryError(const Sophus::SE3d &odometry_error);   // after each registration
    void Reset();                                                   // SetPose()
};

}  // namespace kinematic_icp
