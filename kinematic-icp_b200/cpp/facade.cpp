// Implementation of the C++ facade over the C ABI (libkicp_b200.so).  Builds libkinematic_icp_b200.so, the library a
// ROS 2 workspace links instead of kinematic_icp_registration / _threshold / _pipeline.
#include <cstdlib>
#include <mutex>
#include <utility>

#include "kicp/facade_core.hpp"
#include "kicp/runtime.hpp"
#ifndef KICP_FACADE_NO_PIPELINE
#include "kicp/facade_pipeline.hpp"
#endif

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

void CorrespondenceThreshold::Reset() {
    odom_sse_ = 0.0;
    num_samples_ = 1e-8;
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
// pipeline/KinematicICP.hpp:73-79
KinematicICP::KinematicICP(const Config &config)
    : registration_(config.max_num_iterations, config.convergence_criterion, config.max_num_threads,
                    config.use_adaptive_odometry_regularization, config.fixed_regularization),
      correspondence_threshold_(config.map_resolution(), config.max_range, config.use_adaptive_threshold, config.fixed_threshold),
      config_(config),
      preprocessor_(config.max_range, config.min_range, config.deskew, config.max_num_threads),
      local_map_(config.voxel_size, config.max_range, config.max_points_per_voxel) {}

// pipeline/KinematicICP.hpp:85-89
void KinematicICP::SetPose(const Sophus::SE3d &pose) {
    last_pose_ = pose;
    local_map_.Clear();
    correspondence_threshold_.Reset();
}

// pipeline/KinematicICP.cpp:48-85: one kicp_register_frame call does the per-point work (ingest, de-skew, range filter,
// base transform, both voxel down-samples, registration, map update) with the frame resident in HBM throughout; the
// scalar threshold model and the pose bookkeeping stay here.
KinematicICP::Vector3dVectorTuple KinematicICP::RegisterFrame(const kicp_frame_input &input, const Sophus::SE3d &lidar_to_base,
                                                              const Sophus::SE3d &relative_odometry) {
    // deskew in the lidar frame
    const Sophus::SE3d relative_odometry_in_lidar = lidar_to_base.inverse() * relative_odometry * lidar_to_base;
    double motion[7], l2b[7], last[7], odom[7], out[7];
    kicp::to_pose7(relative_odometry_in_lidar, motion);
    kicp::to_pose7(lidar_to_base, l2b);
    kicp::to_pose7(last_pose_, last);
    kicp::to_pose7(relative_odometry, odom);
    kicp_frame_params fp;
    fp.max_range = preprocessor_.max_range_, fp.min_range = preprocessor_.min_range_, fp.deskew = preprocessor_.deskew_ ? 1 : 0;
    fp.voxel_size = config_.voxel_size;
    fp.reg.max_num_iterations = registration_.max_num_iterations_;
    fp.reg.use_adaptive_odometry_regularization = registration_.use_adaptive_odometry_regularization_ ? 1 : 0;
    fp.reg.convergence_criterion = registration_.convergence_criterion_;
    fp.reg.fixed_regularization = registration_.fixed_regularization_;
    const double tau = correspondence_threshold_.ComputeThreshold();
    fp.stage_clouds = 1;  // the two returned clouds are built below, in one pass each, from the library's pinned staging
    kicp::check(kicp_register_frame(local_map_.handle_, &input, motion, l2b, last, odom, tau, &fp, out, nullptr, 0, nullptr, nullptr, 0,
                                    nullptr, nullptr),
                "kicp_register_frame");
    const double *frame_xyz = nullptr, *source_xyz = nullptr;
    int64_t n_frame = 0, n_source = 0;
    kicp::check(kicp_frame_clouds(kicp::default_context(), &frame_xyz, &n_frame, &source_xyz, &n_source), "kicp_frame_clouds");
    const auto *fb = reinterpret_cast<const Eigen::Vector3d *>(frame_xyz);
    const auto *sb = reinterpret_cast<const Eigen::Vector3d *>(source_xyz);
    Vector3dVector in_base(fb, fb + n_frame), source(sb, sb + n_source);
    const Sophus::SE3d new_pose = kicp::from_pose7(out);
    const Sophus::SE3d odometry_error = (last_pose_ * relative_odometry).inverse() * new_pose;
    correspondence_threshold_.UpdateOdometryError(odometry_error);
    last_pose_ = new_pose;
    return {std::move(in_base), std::move(source)};
}

KinematicICP::Vector3dVectorTuple KinematicICP::RegisterFrame(const std::vector<Eigen::Vector3d> &frame, const std::vector<double> &timestamps,
                                                              const Sophus::SE3d &lidar_to_base, const Sophus::SE3d &relative_odometry) {
    kicp_frame_input input{};
    input.data = kicp::xyz(frame), input.n = (int64_t)frame.size(), input.dtype = KICP_DTYPE_F64, input.point_step = 0;
    input.stamps = timestamps.empty() ? nullptr : timestamps.data(), input.n_stamps = (int64_t)timestamps.size();
    return RegisterFrame(input, lidar_to_base, relative_odometry);
}
}  // namespace pipeline
#endif
}  // namespace kinematic_icp
