// kiss_icp/core/Preprocessing.hpp surface (KISS-ICP v1.2.0): Preprocessor(max_range, min_range, deskew, threads).
// Host-side stand-in for the step before the hot path (pipeline/KinematicICP.cpp:54-57); SURVEY.md §8(f)#2.
#pragma once
#include <Eigen/Core>
#include <algorithm>
#include <sophus/se3.hpp>
#include <vector>

namespace kiss_icp {
struct Preprocessor {
    Preprocessor(const double max_range, const double min_range, const bool deskew, const int max_num_threads)
        : max_range_(max_range), min_range_(min_range), deskew_(deskew), max_num_threads_(max_num_threads) {}

    std::vector<Eigen::Vector3d> Preprocess(const std::vector<Eigen::Vector3d> &frame, const std::vector<double> &timestamps,
                                            const Sophus::SE3d &relative_motion) const {
        std::vector<Eigen::Vector3d> deskewed;
        const std::vector<Eigen::Vector3d> *src = &frame;
        if (deskew_ && !timestamps.empty()) {
            const auto mm = std::minmax_element(timestamps.cbegin(), timestamps.cend());
            const double min_time = *mm.first, max_time = *mm.second;
            const Sophus::SE3d::Tangent omega = relative_motion.log();
            deskewed.resize(frame.size());
            for (size_t i = 0; i < frame.size(); ++i) {
                const double stamp = (timestamps[i] - min_time) / (max_time - min_time);
                deskewed[i] = Sophus::SE3d::exp(omega * (stamp - 1.0)) * frame[i];
            }
            src = &deskewed;
        }
        std::vector<Eigen::Vector3d> out;
        out.reserve(src->size());
        for (const auto &p : *src) {
            const double r = p.norm();
            if (r < max_range_ && r > min_range_) out.push_back(p);
        }
        return out;
    }
    double max_range_;
    double min_range_;
    bool deskew_;
    int max_num_threads_;
};
}  // namespace kiss_icp
