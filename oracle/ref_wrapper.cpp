// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// extern "C" wrapper around the REFERENCE'S OWN first-party sources, compiled in place from /root/reference by
// oracle/Makefile into oracle/_ref/libkicp_ref.so:
//   cpp/kinematic_icp/registration/Registration.cpp              (KinematicRegistration::ComputeRobotMotion)
//   cpp/kinematic_icp/correspondence_threshold/CorrespondenceThreshold.cpp
//   cpp/kinematic_icp/pipeline/KinematicICP.cpp                  (KinematicICP::RegisterFrame)
// Their Eigen/Sophus/TBB/kiss_icp includes resolve to header shims (those libraries are absent offline); the
// kiss_icp::VoxelHashMap behind them is the restated CPU map.  This is the "reference" CPU arm of bench.py and the
// generator of tests/golden/*.npz.
#include <tbb/global_control.h>
#include <tbb/task_arena.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "kinematic_icp/correspondence_threshold/CorrespondenceThreshold.hpp"
#include "kinematic_icp/pipeline/KinematicICP.hpp"
#include "kinematic_icp/registration/Registration.hpp"

namespace {
std::vector<Eigen::Vector3d> to_eigen(const double *xyz, int64_t n) {
    std::vector<Eigen::Vector3d> v(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) v[i] = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    return v;
}
Sophus::SE3d to_se3(const double *p) {
    // pose7 = {qx,qy,qz,qw,tx,ty,tz}; set the quaternion without renormalising through a rotation matrix
    return Sophus::SE3d(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]));
}
void from_se3(const Sophus::SE3d &T, double *p) {
    const auto &q = T.unit_quaternion();
    p[0] = q.x(), p[1] = q.y(), p[2] = q.z(), p[3] = q.w();
    p[4] = T.translation().x(), p[5] = T.translation().y(), p[6] = T.translation().z();
}
}  // namespace

extern "C" {

// (this library holds its own copy of the restated KISS-ICP code: the down-sample order mode is set per library)
void kref_set_downsample_order(int mode) { kicp_oracle::SetDownsampleOrder(mode); }

void *kref_map_create(double voxel_size, double max_distance, unsigned max_points_per_voxel) {
    return new kiss_icp::VoxelHashMap(voxel_size, max_distance, max_points_per_voxel);
}
void kref_map_destroy(void *h) { delete static_cast<kiss_icp::VoxelHashMap *>(h); }
void kref_map_add_points(void *h, const double *xyz, int64_t n) {
    static_cast<kiss_icp::VoxelHashMap *>(h)->AddPoints(to_eigen(xyz, n));
}
void kref_map_update_pose(void *h, const double *xyz, int64_t n, const double *pose7) {
    static_cast<kiss_icp::VoxelHashMap *>(h)->Update(to_eigen(xyz, n), to_se3(pose7));
}
int64_t kref_map_num_points(void *h) {
    return static_cast<int64_t>(static_cast<kiss_icp::VoxelHashMap *>(h)->impl_.NumPoints());
}

// kinematic_icp::KinematicRegistration::ComputeRobotMotion, the reference's own code path.
// threads <= 0 means all cores (Registration.cpp:139-143).  NOTE the reference installs its TBB cap in a
// function-local static (:147-148) so only the FIRST construction in a process would take effect with real TBB;
// the shim's global_control simply stores the latest value.
void kref_register(void *h, const double *frame, int64_t n, const double *last_pose7, const double *rel_odom7, double tau,
                   int max_iter, double conv, int adaptive, double fixed_reg, int threads, double *out_pose7) {
    tbb::global_control cap(tbb::global_control::max_allowed_parallelism,
                            static_cast<size_t>(threads > 0 ? threads : tbb::this_task_arena::max_concurrency()));
    kinematic_icp::KinematicRegistration reg(max_iter, conv, threads, adaptive != 0, fixed_reg);
    const Sophus::SE3d T = reg.ComputeRobotMotion(to_eigen(frame, n), *static_cast<kiss_icp::VoxelHashMap *>(h),
                                                  to_se3(last_pose7), to_se3(rel_odom7), tau);
    from_se3(T, out_pose7);
}

// CorrespondenceThreshold: feed a sequence of odometry errors, get tau after each update.
void kref_threshold_sequence(double map_err, double max_range, int adaptive, double fixed, const double *err7s, int64_t n,
                             double *taus) {
    kinematic_icp::CorrespondenceThreshold th(map_err, max_range, adaptive != 0, fixed);
    for (int64_t i = 0; i < n; ++i) {
        th.UpdateOdometryError(to_se3(err7s + 7 * i));
        taus[i] = th.ComputeThreshold();
    }
}

// kinematic_icp::pipeline::KinematicICP — the whole per-frame pipeline of the reference.
void *kref_pipeline_create(double max_range, double min_range, double voxel_size, unsigned max_points_per_voxel,
                           int use_adaptive_threshold, double fixed_threshold, int max_num_iterations,
                           double convergence_criterion, int max_num_threads, int use_adaptive_reg, double fixed_reg,
                           int deskew) {
    kinematic_icp::pipeline::Config c;
    c.max_range = max_range;
    c.min_range = min_range;
    c.voxel_size = voxel_size;
    c.max_points_per_voxel = max_points_per_voxel;
    c.use_adaptive_threshold = use_adaptive_threshold != 0;
    c.fixed_threshold = fixed_threshold;
    c.max_num_iterations = max_num_iterations;
    c.convergence_criterion = convergence_criterion;
    c.max_num_threads = max_num_threads;
    c.use_adaptive_odometry_regularization = use_adaptive_reg != 0;
    c.fixed_regularization = fixed_reg;
    c.deskew = deskew != 0;
    tbb::global_control cap(tbb::global_control::max_allowed_parallelism,
                            static_cast<size_t>(max_num_threads > 0 ? max_num_threads : tbb::this_task_arena::max_concurrency()));
    return new kinematic_icp::pipeline::KinematicICP(c);
}
void kref_pipeline_destroy(void *h) { delete static_cast<kinematic_icp::pipeline::KinematicICP *>(h); }
void kref_pipeline_set_pose(void *h, const double *pose7) {
    static_cast<kinematic_icp::pipeline::KinematicICP *>(h)->SetPose(to_se3(pose7));
}
// returns the number of `source` points used for registration; out_pose7 = pose() after the frame
int64_t kref_pipeline_register_frame(void *h, const double *xyz, int64_t n, const double *stamps, int64_t n_stamps,
                                     const double *lidar_to_base7, const double *rel_odom7, double *out_pose7) {
    auto *p = static_cast<kinematic_icp::pipeline::KinematicICP *>(h);
    std::vector<double> ts(stamps, stamps + n_stamps);
    const auto [frame, source] = p->RegisterFrame(to_eigen(xyz, n), ts, to_se3(lidar_to_base7), to_se3(rel_odom7));
    from_se3(p->pose(), out_pose7);
    return static_cast<int64_t>(source.size());
}
int64_t kref_pipeline_num_map_points(void *h) {
    return static_cast<int64_t>(static_cast<kinematic_icp::pipeline::KinematicICP *>(h)->LocalMap().size());
}

}  // extern "C"
