// kiss_icp/core/VoxelUtils.hpp surface (KISS-ICP v1.2.0): Voxel, PointToVoxel, VoxelDownsample.
// VoxelDownsample runs on the device (kicp_voxel_downsample, include/kicp.h): the first point in input order of every
// voxel, returned in input order.  (The original returns the same set in robin_map iteration order; a caller that down-samples the
// result again or feeds it to the map's greedy insert — as the pipeline does — therefore sees other points than with the library:
// DESIGN.md section 2 has the measured size of that gap.)
#pragma once
#include <Eigen/Core>
#include <cmath>
#include <cstdint>
#include <vector>

#include "kicp/runtime.hpp"

namespace kiss_icp {
using Voxel = Eigen::Vector3i;
inline Voxel PointToVoxel(const Eigen::Vector3d &point, const double voxel_size) {
    return Voxel(static_cast<int>(std::floor(point.x() / voxel_size)), static_cast<int>(std::floor(point.y() / voxel_size)),
                 static_cast<int>(std::floor(point.z() / voxel_size)));
}
inline std::vector<Eigen::Vector3d> VoxelDownsample(const std::vector<Eigen::Vector3d> &frame, const double voxel_size) {
    std::vector<Eigen::Vector3d> out(frame.size());
    int64_t m = 0;
    kicp::check(kicp_voxel_downsample(kicp::default_context(), kicp::xyz(frame), (int64_t)frame.size(), voxel_size,
                                      out.empty() ? nullptr : out.front().data(), (int64_t)out.size(), &m),
                "kicp_voxel_downsample");
    out.resize(static_cast<size_t>(m));
    return out;
}
}  // namespace kiss_icp
