// Implementation of the C++ facade over the C ABI (libkicp_b200.so).  Builds libkinematic_icp_b200.so, the library a
// ROS 2 workspace links instead of kinematic_icp_registration / _threshold / _pipeline.
#include <cstdlib>
#include <mutex>

#include "kicp/runtime.hpp"
#include "kinematic_icp/correspondence_threshold/CorrespondenceThreshold.hpp"
#include "kinematic_icp/pipeline/KinematicICP.hpp"
#include "kinematic_icp/registration/Registration.hpp"

namespace kicp {
kicp_ctx *default_context() {
    static kicp_ctx *ctx = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *dev = std::getenv("KICP_DEVICE");
        check(kicp_ctx_create(dev ? std::atoi(dev) : 0, &ctx), "kicp_ctx_create");
    });
    return ctx;
}
}  // namespace kicp

namespace kinematic_icp {

// registration/Registration.cpp:132-149 (the TBB thread cap has no GPU meaning)
KinematicRegistration::KinematicRegistration(const int max_num_iteration, const double convergence_criterion, const int max_num_threads,
                                             const bool use_adaptive_odometry_regularization, const double fixed_regularization)
    : max_num_iterations_(max_num_iteration),
      convergence_criterion_(convergence_criterion),
      max_num_threads_(max_num_threads),
      use_adaptive_odometry_regularization_(use_adaptive_odometry_regularization),
      fixed_regularization_(fixed_regularization) {}

// registration/Registration.cpp:151-190, on the device
Sophus::SE3d KinematicRegistration::ComputeRobotMotion(const std::vector<Eigen::Vector3d> &frame, const kiss_icp::VoxelHashMap &voxel_map,
                                                       const Sophus::SE3d &last_robot_pose, const Sophus::SE3d &relative_wheel_odometry,
                                                       const double max_correspondence_distance) {
    double last[7], odom[7], out[7];
    kicp::to_pose7(last_robot_pose, last);
    kicp::to_pose7(relative_wheel_odometry, odom);
    kicp_reg_params p;
    p.max_num_iterations = max_num_iterations_;
    p.use_adaptive_odometry_regularization = use_adaptive_odometry_regularization_ ? 1 : 0;
    p.convergence_criterion = convergence_criterion_;
    p.fixed_regularization = fixed_regularization_;
    kicp::check(kicp_register(voxel_map.handle_, kicp::xyz(frame), (int64_t)frame.size(), last, odom, max_correspondence_distance, &p, out,
                              nullptr),
                "kicp_register");
    return kicp::from_pose7(out);
}

// correspondence_threshold/CorrespondenceThreshold.cpp:36-64
CorrespondenceThreshold::CorrespondenceThreshold(const double map_discretization_error, const double max_range,
                                                 const bool use_adaptive_threshold, const double fixed_threshold)
    : map_discretization_error_(map_discretization_error),
      max_range_(max_range),
      use_adaptive_threshold_(use_adaptive_threshold),
      fixed_threshold_(fixed_threshold),
      odom_sse_(0.0),
      num_samples_(1e-8) {}

double CorrespondenceThreshold::ComputeThreshold() const {
    if (!use_adaptive_threshold_) return fixed_threshold_;
    const double sigma_odom = std::sqrt(odom_sse_ / num_samples_);
    return 3.0 * (map_discretization_error_ + sigma_odom);
}

void CorrespondenceThreshold::UpdateOdometryError(const Sophus::SE3d &odometry_error) {
    if (!use_adaptive_threshold_) return;
    // odometry error expressed in point space (:29-34): translation plus the chord a max_range lever arm sweeps
    const double theta = odometry_error.so3().logAndTheta().theta;
    const double e = odometry_error.translation().norm() + 2.0 * max_range_ * std::sin(theta / 2.0);
    odom_sse_ += e * e;
    num_samples_ += 1.0;
}

#ifndef KICP_FACADE_NO_PIPELINE  // the reference's own pipeline/KinematicICP.cpp can be compiled in its place
namespace pipeline {
// pipeline/KinematicICP.cpp:48-85
KinematicICP::Vector3dVectorTuple KinematicICP::RegisterFrame(const std::vector<Eigen::Vector3d> &frame, const std::vector<double> &timestamps,
                                                              const Sophus::SE3d &lidar_to_base, const Sophus::SE3d &relative_odometry) {
    // deskew in the lidar frame, range filter
    const Sophus::SE3d relative_odometry_in_lidar = lidar_to_base.inverse() * relative_odometry * lidar_to_base;
    const Vector3dVector preprocessed = preprocessor_.Preprocess(frame, timestamps, relative_odometry_in_lidar);
    Vector3dVector in_base(preprocessed.size());
    for (size_t i = 0; i < preprocessed.size(); ++i) in_base[i] = lidar_to_base * preprocessed[i];
    // two voxel down-samples: 0.5 vs for the map update, then 1.5 vs for the registration source
    const Vector3dVector frame_downsample = kiss_icp::VoxelDownsample(in_base, config_.voxel_size * 0.5);
    const Vector3dVector source = kiss_icp::VoxelDownsample(frame_downsample, config_.voxel_size * 1.5);
    const double tau = correspondence_threshold_.ComputeThreshold();
    const Sophus::SE3d new_pose = registration_.ComputeRobotMotion(source, local_map_, last_pose_, relative_odometry, tau);
    const Sophus::SE3d odometry_error = (last_pose_ * relative_odometry).inverse() * new_pose;
    correspondence_threshold_.UpdateOdometryError(odometry_error);
    local_map_.Update(frame_downsample, new_pose);
    last_pose_ = new_pose;
    return {in_base, source};
}
}  // namespace pipeline
#endif
}  // namespace kinematic_icp
