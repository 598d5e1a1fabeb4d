// kiss_icp/core/VoxelUtils.hpp surface (KISS-ICP v1.2.0): Voxel, PointToVoxel, VoxelDownsample.
// Host-side stand-in for the part of kiss_icp_core that sits BEFORE the hot path (pipeline/KinematicICP.cpp:38-44);
// moving it to the device is row (f)#2 of SURVEY.md §8.  On a machine with KISS-ICP installed, link the real one.
#pragma once
#include <Eigen/Core>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <vector>

namespace kiss_icp {
using Voxel = Eigen::Vector3i;
inline Voxel PointToVoxel(const Eigen::Vector3d &point, const double voxel_size) {
    return Voxel(static_cast<int>(std::floor(point.x() / voxel_size)), static_cast<int>(std::floor(point.y() / voxel_size)),
                 static_cast<int>(std::floor(point.z() / voxel_size)));
}
// first point (input order) per voxel; output in order of first occurrence
inline std::vector<Eigen::Vector3d> VoxelDownsample(const std::vector<Eigen::Vector3d> &frame, const double voxel_size) {
    struct H {
        size_t operator()(uint64_t k) const { return static_cast<size_t>(k * 0x9E3779B97F4A7C15ull); }
    };
    std::unordered_map<uint64_t, int, H> grid;
    grid.reserve(frame.size());
    std::vector<Eigen::Vector3d> out;
    out.reserve(frame.size());
    for (const auto &p : frame) {
        const Voxel v = PointToVoxel(p, voxel_size);
        const uint64_t key = (static_cast<uint64_t>(static_cast<uint32_t>(v.x())) & 0x1FFFFF) |
                             ((static_cast<uint64_t>(static_cast<uint32_t>(v.y())) & 0x1FFFFF) << 21) |
                             ((static_cast<uint64_t>(static_cast<uint32_t>(v.z())) & 0x1FFFFF) << 42);
        if (grid.emplace(key, 1).second) out.push_back(p);
    }
    return out;
}
}  // namespace kiss_icp
