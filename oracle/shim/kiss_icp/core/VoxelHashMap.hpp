// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// kiss_icp/core/VoxelHashMap.hpp (KISS-ICP v1.2.0) surface for oracle/_ref, backed by the restated CPU map.
#pragma once
#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <tuple>
#include <vector>

#include "VoxelUtils.hpp"
#include "kicp_oracle.hpp"

namespace kiss_icp {
struct VoxelHashMap {
    explicit VoxelHashMap(double voxel_size, double max_distance, unsigned int max_points_per_voxel)
        : voxel_size_(voxel_size),
          max_distance_(max_distance),
          max_points_per_voxel_(max_points_per_voxel),
          impl_(voxel_size, max_distance, max_points_per_voxel) {}

    inline void Clear() { impl_.Clear(); }
    inline bool Empty() const { return impl_.Empty(); }
    void Update(const std::vector<Eigen::Vector3d> &points, const Eigen::Vector3d &origin) {
        impl_.Update(shim::to_oracle(points), kicp_oracle::Vec3{origin.x(), origin.y(), origin.z()});
    }
    void Update(const std::vector<Eigen::Vector3d> &points, const Sophus::SE3d &pose) {
        // KISS: transform every point by the pose, then Update(transformed, pose.translation())
        std::vector<kicp_oracle::Vec3> t(points.size());
        for (size_t i = 0; i < points.size(); ++i) {
            const Eigen::Vector3d q = pose * points[i];
            t[i] = {q.x(), q.y(), q.z()};
        }
        const Eigen::Vector3d &o = pose.translation();
        impl_.Update(t, kicp_oracle::Vec3{o.x(), o.y(), o.z()});
    }
    void AddPoints(const std::vector<Eigen::Vector3d> &points) { impl_.AddPoints(shim::to_oracle(points)); }
    void RemovePointsFarFromLocation(const Eigen::Vector3d &origin) {
        impl_.RemovePointsFarFromLocation({origin.x(), origin.y(), origin.z()});
    }
    std::vector<Eigen::Vector3d> Pointcloud() const { return shim::from_oracle(impl_.Pointcloud()); }
    std::tuple<Eigen::Vector3d, double> GetClosestNeighbor(const Eigen::Vector3d &query) const {
        const auto r = impl_.GetClosestNeighbor({query.x(), query.y(), query.z()});
        return std::make_tuple(Eigen::Vector3d(r.first.x, r.first.y, r.first.z), r.second);
    }

    double voxel_size_;
    double max_distance_;
    unsigned int max_points_per_voxel_;
    kicp_oracle::VoxelHashMap impl_;
};
}  // namespace kiss_icp
