// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// kiss_icp/core/VoxelUtils.hpp (KISS-ICP v1.2.0) surface for oracle/_ref: Eigen-typed wrappers over the restated
// oracle (kicp_oracle.hpp).  KISS-ICP itself is not available offline (kiss-icp.cmake:29-31 fetches it).
#pragma once
#include <Eigen/Core>
#include <vector>

#include "kicp_oracle.hpp"

namespace kiss_icp {
using Voxel = Eigen::Vector3i;
inline Voxel PointToVoxel(const Eigen::Vector3d &point, const double voxel_size) {
    const auto v = kicp_oracle::PointToVoxel({point.x(), point.y(), point.z()}, voxel_size);
    return Voxel(v.x, v.y, v.z);
}
namespace shim {
inline std::vector<kicp_oracle::Vec3> to_oracle(const std::vector<Eigen::Vector3d> &v) {
    std::vector<kicp_oracle::Vec3> r(v.size());
    for (size_t i = 0; i < v.size(); ++i) r[i] = {v[i].x(), v[i].y(), v[i].z()};
    return r;
}
inline std::vector<Eigen::Vector3d> from_oracle(const std::vector<kicp_oracle::Vec3> &v) {
    std::vector<Eigen::Vector3d> r(v.size());
    for (size_t i = 0; i < v.size(); ++i) r[i] = Eigen::Vector3d(v[i].x, v[i].y, v[i].z);
    return r;
}
}  // namespace shim
inline std::vector<Eigen::Vector3d> VoxelDownsample(const std::vector<Eigen::Vector3d> &frame, const double voxel_size) {
    return shim::from_oracle(kicp_oracle::VoxelDownsample(shim::to_oracle(frame), voxel_size));
}
}  // namespace kiss_icp
