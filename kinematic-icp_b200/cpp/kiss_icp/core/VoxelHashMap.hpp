// kiss_icp::VoxelHashMap — same name, namespace, constructor and methods as KISS-ICP v1.2.0's
// cpp/kiss_icp/core/VoxelHashMap.hpp (the header the reference includes at registration/Registration.hpp:26 and
// pipeline/KinematicICP.hpp:29), with the storage and every operation in HBM behind include/kicp.h.
// The robin_map member `map_` of the original has no counterpart (no code of the reference reads it).
#pragma once
#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <tuple>
#include <vector>

#include "kicp/runtime.hpp"
#include "kiss_icp/core/VoxelUtils.hpp"

namespace kiss_icp {
struct VoxelHashMap {
    explicit VoxelHashMap(double voxel_size, double max_distance, unsigned int max_points_per_voxel)
        : voxel_size_(voxel_size), max_distance_(max_distance), max_points_per_voxel_(max_points_per_voxel) {
        kicp::check(kicp_map_create(kicp::default_context(), voxel_size, max_distance, max_points_per_voxel, &handle_),
                    "kicp_map_create");
    }
    ~VoxelHashMap() { kicp_map_destroy(handle_); }
    VoxelHashMap(const VoxelHashMap &) = delete;
    VoxelHashMap &operator=(const VoxelHashMap &) = delete;
    VoxelHashMap(VoxelHashMap &&o) noexcept
        : voxel_size_(o.voxel_size_), max_distance_(o.max_distance_), max_points_per_voxel_(o.max_points_per_voxel_), handle_(o.handle_) {
        o.handle_ = nullptr;
    }

    inline void Clear() { kicp::check(kicp_map_clear(handle_), "kicp_map_clear"); }
    inline bool Empty() const {
        int32_t e = 1;
        kicp::check(kicp_map_empty(handle_, &e), "kicp_map_empty");
        return e != 0;
    }
    void Update(const std::vector<Eigen::Vector3d> &points, const Eigen::Vector3d &origin) {
        kicp::check(kicp_map_update(handle_, kicp::xyz(points), (int64_t)points.size(), origin.data()), "kicp_map_update");
    }
    void Update(const std::vector<Eigen::Vector3d> &points, const Sophus::SE3d &pose) {
        double p[7];
        kicp::to_pose7(pose, p);
        kicp::check(kicp_map_update_pose(handle_, kicp::xyz(points), (int64_t)points.size(), p), "kicp_map_update_pose");
    }
    void AddPoints(const std::vector<Eigen::Vector3d> &points) {
        kicp::check(kicp_map_add_points(handle_, kicp::xyz(points), (int64_t)points.size()), "kicp_map_add_points");
    }
    void RemovePointsFarFromLocation(const Eigen::Vector3d &origin) {
        kicp::check(kicp_map_remove_far(handle_, origin.data()), "kicp_map_remove_far");
    }
    std::vector<Eigen::Vector3d> Pointcloud() const {
        int64_t n = 0;
        kicp::check(kicp_map_num_points(handle_, &n), "kicp_map_num_points");
        std::vector<Eigen::Vector3d> out(static_cast<size_t>(n));
        if (n > 0) kicp::check(kicp_map_pointcloud(handle_, out.front().data(), n, &n), "kicp_map_pointcloud");
        return out;
    }
    std::tuple<Eigen::Vector3d, double> GetClosestNeighbor(const Eigen::Vector3d &query) const {
        Eigen::Vector3d p;
        double d = 0.0;
        kicp::check(kicp_map_nearest(handle_, query.data(), 1, p.data(), &d), "kicp_map_nearest");
        return std::make_tuple(p, d);
    }

    double voxel_size_;
    double max_distance_;
    unsigned int max_points_per_voxel_;
    kicp_map *handle_ = nullptr;  // the HBM-resident map (replaces tsl::robin_map<Voxel, std::vector<Eigen::Vector3d>> map_)
};
}  // namespace kiss_icp
