// TUM trajectory writer: one line per pose, "timestamp x y z qx qy qz qw" with 6 fixed decimals — the format of
// OfflineNode::writePosesInTumFormat (ros/src/kinematic_icp_ros/nodes/offline_node.cpp:76-97).
#pragma once
#include <cstdio>
#include <sophus/se3.hpp>
#include <string>
#include <utility>
#include <vector>

namespace kicp {
inline bool write_poses_tum(const std::string &path, const std::vector<std::pair<double, Sophus::SE3d>> &poses_with_timestamps) {
    // stdio rather than <fstream>: the facade links libstdc++ statically, whose iostream locale state is not set up
    // inside a dlopen'ed library; "%.6f" is what std::fixed << std::setprecision(6) prints.
    std::FILE *file = std::fopen(path.c_str(), "w");
    if (file == nullptr) return false;
    for (const auto &[timestamp, pose] : poses_with_timestamps) {
        const auto &t = pose.translation();
        const auto &q = pose.unit_quaternion();
        std::fprintf(file, "%.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f\n", timestamp, t.x(), t.y(), t.z(), q.x(), q.y(), q.z(), q.w());
    }
    std::fclose(file);
    return true;
}
}  // namespace kicp
