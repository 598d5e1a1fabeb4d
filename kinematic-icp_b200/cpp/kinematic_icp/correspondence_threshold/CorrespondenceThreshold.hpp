// kinematic_icp::CorrespondenceThreshold — the reference's correspondence_threshold/CorrespondenceThreshold.hpp:30-55.
// Scalar, once per frame: stays on the host (it produces the `tau` argument of the hot path).
#pragma once
#include <cmath>
#include <sophus/se3.hpp>

namespace kinematic_icp {
struct CorrespondenceThreshold {
    explicit CorrespondenceThreshold(const double map_discretization_error, const double max_range, const bool use_adaptive_threshold,
                                     const double fixed_threshold);
    void UpdateOdometryError(const Sophus::SE3d &odometry_error);
    double ComputeThreshold() const;
    inline void Reset() {
        odom_sse_ = 0.0;
        num_samples_ = 1e-8;
    }
    double map_discretization_error_;
    double max_range_;
    bool use_adaptive_threshold_;
    double fixed_threshold_;
    double odom_sse_;
    double num_samples_;
};
}  // namespace kinematic_icp
