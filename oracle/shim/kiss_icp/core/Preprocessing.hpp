// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// kiss_icp/core/Preprocessing.hpp (KISS-ICP v1.2.0) surface for oracle/_ref, backed by the restated CPU oracle.
#pragma once
#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <vector>

#include "VoxelUtils.hpp"
#include "kicp_oracle.hpp"

namespace kiss_icp {
struct Preprocessor {
    Preprocessor(const double max_range, const double min_range, const bool deskew, const int max_num_threads)
        : max_range_(max_range), min_range_(min_range), deskew_(deskew), max_num_threads_(max_num_threads) {}
    std::vector<Eigen::Vector3d> Preprocess(const std::vector<Eigen::Vector3d> &frame, const std::vector<double> &timestamps,
                                            const Sophus::SE3d &relative_motion) const {
        kicp_oracle::SE3 T;
        const auto &q = relative_motion.unit_quaternion();
        T.q = {q.x(), q.y(), q.z(), q.w()};
        T.t = {relative_motion.translation().x(), relative_motion.translation().y(), relative_motion.translation().z()};
        return shim::from_oracle(kicp_oracle::Preprocess(shim::to_oracle(frame), timestamps, T, max_range_, min_range_, deskew_));
    }
    double max_range_;
    double min_range_;
    bool deskew_;
    int max_num_threads_;
};
}  // namespace kiss_icp
