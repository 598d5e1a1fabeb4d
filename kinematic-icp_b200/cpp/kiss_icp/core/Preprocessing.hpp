// kiss_icp/core/Preprocessing.hpp surface (KISS-ICP v1.2.0): Preprocessor(max_range, min_range, deskew, threads).
// Preprocess runs on the device (kicp_preprocess, include/kicp.h): de-skew by exp((s-1) log(relative_motion)) with the
// stamps normalised to [0,1], then the range filter min_range < |p| < max_range; survivors keep their input order.
#pragma once
#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <vector>

#include "kicp/runtime.hpp"

namespace kiss_icp {
struct Preprocessor {
    Preprocessor(const double max_range, const double min_range, const bool deskew, const int max_num_threads)
        : max_range_(max_range), min_range_(min_range), deskew_(deskew), max_num_threads_(max_num_threads) {}

    std::vector<Eigen::Vector3d> Preprocess(const std::vector<Eigen::Vector3d> &frame, const std::vector<double> &timestamps,
                                            const Sophus::SE3d &relative_motion) const {
        double motion[7];
        const double identity[7] = {0, 0, 0, 1, 0, 0, 0};
        kicp::to_pose7(relative_motion, motion);
        std::vector<Eigen::Vector3d> out(frame.size());
        int64_t m = 0;
        kicp::check(kicp_preprocess(kicp::default_context(), kicp::xyz(frame), (int64_t)frame.size(),
                                    timestamps.empty() ? nullptr : timestamps.data(), (int64_t)timestamps.size(), motion, identity,
                                    max_range_, min_range_, deskew_ ? 1 : 0, out.empty() ? nullptr : out.front().data(),
                                    (int64_t)out.size(), &m),
                    "kicp_preprocess");
        out.resize(static_cast<size_t>(m));
        return out;
    }
    double max_range_;
    double min_range_;
    bool deskew_;
    int max_num_threads_;
};
}  // namespace kiss_icp
