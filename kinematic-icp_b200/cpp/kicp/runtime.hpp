// Process-wide glue between the C++ facade classes and the C ABI (include/kicp.h): one default kicp_ctx, pose
// conversion, and error translation.  The reference's core throws nothing and returns no status;
// the facade throws std::runtime_error only for conditions the reference cannot have (no CUDA device, out of memory).
#pragma once
#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <stdexcept>
#include <string>

#include "kicp.h"

namespace kicp {

kicp_ctx *default_context();  // created on first use on device $KICP_DEVICE (default 0); defined in facade.cpp

inline void check(int status, const char *what) {
    if (status != KICP_OK && status != KICP_WARN_NO_CORRESPONDENCES)
        throw std::runtime_error(std::string(what) + ": " + kicp_status_string(status) + " — " + kicp_last_error());
}

inline void to_pose7(const Sophus::SE3d &T, double p[7]) {
    const auto &q = T.unit_quaternion();
    p[0] = q.x(), p[1] = q.y(), p[2] = q.z(), p[3] = q.w();
    p[4] = T.translation().x(), p[5] = T.translation().y(), p[6] = T.translation().z();
}
inline Sophus::SE3d from_pose7(const double p[7]) {
    return Sophus::SE3d(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]));
}
// std::vector<Eigen::Vector3d> is 3 contiguous doubles per point: the C ABI's row-major xyz layout
inline const double *xyz(const std::vector<Eigen::Vector3d> &v) { return v.empty() ? nullptr : v.front().data(); }
static_assert(sizeof(Eigen::Vector3d) == 3 * sizeof(double), "Eigen::Vector3d must be 3 packed doubles");

}  // namespace kicp
