// extern "C" test hooks over the C++ facade (so tests can drive kinematic_icp::pipeline::KinematicICP, the PointCloud2 decoder and the TUM
// writer from Python).  Test infrastructure: built as tests/hooks/_build/libkicp_facade_hooks.so by kinematic-icp_b200/cpp/Makefile, linked
// AGAINST the product libraries, never into them.
#include <cstdint>
#include <vector>

#include "kicp/pointcloud2.hpp"
#include "kicp/tum.hpp"
#include "kinematic_icp/pipeline/KinematicICP.hpp"

namespace {
std::vector<Eigen::Vector3d> to_eigen(const double *xyz, int64_t n) {
    std::vector<Eigen::Vector3d> v(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) v[i] = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    return v;
}
}  // namespace

extern "C" {
void *kfac_pipeline_create(double max_range, double min_range, double voxel_size, unsigned max_points_per_voxel,
                           int use_adaptive_threshold, double fixed_threshold, int max_num_iterations, double convergence_criterion,
                           int max_num_threads, int use_adaptive_reg, double fixed_reg, int deskew) {
    kinematic_icp::pipeline::Config c;
    c.max_range = max_range, c.min_range = min_range, c.voxel_size = voxel_size, c.max_points_per_voxel = max_points_per_voxel;
    c.use_adaptive_threshold = use_adaptive_threshold != 0, c.fixed_threshold = fixed_threshold;
    c.max_num_iterations = max_num_iterations, c.convergence_criterion = convergence_criterion, c.max_num_threads = max_num_threads;
    c.use_adaptive_odometry_regularization = use_adaptive_reg != 0, c.fixed_regularization = fixed_reg, c.deskew = deskew != 0;
    try {
        return new kinematic_icp::pipeline::KinematicICP(c);
    } catch (const std::exception &) {
        return nullptr;
    }
}
void kfac_pipeline_destroy(void *h) { delete static_cast<kinematic_icp::pipeline::KinematicICP *>(h); }
void kfac_pipeline_set_pose(void *h, const double *pose7) {
    static_cast<kinematic_icp::pipeline::KinematicICP *>(h)->SetPose(kicp::from_pose7(pose7));
}
int64_t kfac_pipeline_register_frame(void *h, const double *xyz, int64_t n, const double *stamps, int64_t n_stamps,
                                     const double *lidar_to_base7, const double *rel_odom7, double *out_pose7) {
    auto *p = static_cast<kinematic_icp::pipeline::KinematicICP *>(h);
    std::vector<double> ts(stamps, stamps + n_stamps);
    const auto [frame, source] = p->RegisterFrame(to_eigen(xyz, n), ts, kicp::from_pose7(lidar_to_base7), kicp::from_pose7(rel_odom7));
    kicp::to_pose7(p->pose(), out_pose7);
    return static_cast<int64_t>(source.size());
}
// the caller's buffer goes to the device as it is: float64 (dtype 0) or float32 (dtype 1) x,y,z records at `point_step` bytes
// (PointCloud2 layout; 0 = packed), no intermediate std::vector and no host-side widening
int64_t kfac_pipeline_register_frame_raw(void *h, const void *data, int64_t n, int32_t dtype, int32_t point_step, int32_t ox, int32_t oy,
                                         int32_t oz, const double *stamps, int64_t n_stamps, const double *lidar_to_base7,
                                         const double *rel_odom7, double *out_pose7) {
    auto *p = static_cast<kinematic_icp::pipeline::KinematicICP *>(h);
    kicp_frame_input in{};
    in.data = data, in.n = n, in.dtype = dtype, in.point_step = point_step, in.offset_x = ox, in.offset_y = oy, in.offset_z = oz;
    in.stamps = n_stamps ? stamps : nullptr, in.n_stamps = n_stamps;
    const auto [frame, source] = p->RegisterFrame(in, kicp::from_pose7(lidar_to_base7), kicp::from_pose7(rel_odom7));
    kicp::to_pose7(p->pose(), out_pose7);
    return static_cast<int64_t>(source.size());
}
int64_t kfac_pipeline_num_map_points(void *h) {
    return static_cast<int64_t>(static_cast<kinematic_icp::pipeline::KinematicICP *>(h)->LocalMap().size());
}
// host-side stage times (ms, cumulative) of the last RegisterFrame on the process-wide context (library debug export)
extern "C" int kicp_debug_frame_timing(kicp_ctx *, double *);
int kfac_debug_frame_timing(double *out8) { return kicp_debug_frame_timing(kicp::default_context(), out8); }
// PointCloud2-shaped buffer -> ingest layout + per-point stamps (kicp/pointcloud2.hpp).  Fields arrive as parallel arrays, names
// joined by ','.  layout5 = {dtype, point_step, offset_x, offset_y, offset_z}; returns the number of stamps written (0 = none),
// -1 on a decoding error.
int64_t kfac_pc2_decode(const uint8_t *data, uint32_t width, uint32_t height, uint32_t point_step, const char *names_csv,
                        const uint32_t *offsets, const uint8_t *datatypes, const uint32_t *counts, int32_t n_fields, int32_t *layout5,
                        double *stamps_out) {
    try {
        kicp::PointCloud2View m;
        m.data = data, m.width = width, m.height = height, m.point_step = point_step;
        std::string names(names_csv);
        size_t pos = 0;
        for (int32_t i = 0; i < n_fields; ++i) {
            const size_t next = names.find(',', pos);
            kicp::PointFieldView f;
            f.name = names.substr(pos, next == std::string::npos ? std::string::npos : next - pos);
            f.offset = offsets[i], f.datatype = datatypes[i], f.count = counts[i];
            m.fields.push_back(f);
            pos = next == std::string::npos ? names.size() : next + 1;
        }
        const kicp_frame_input in = kicp::make_frame_input(m);
        layout5[0] = in.dtype, layout5[1] = in.point_step, layout5[2] = in.offset_x, layout5[3] = in.offset_y, layout5[4] = in.offset_z;
        const std::vector<double> stamps = kicp::extract_timestamps(m);
        for (size_t i = 0; i < stamps.size(); ++i) stamps_out[i] = stamps[i];
        return static_cast<int64_t>(stamps.size());
    } catch (const std::exception &) {
        return -1;
    }
}
void kfac_process_timestamps(double *stamps, int64_t n, double header_stamp, double *last_processed, double *begin_end2) {
    std::vector<double> v(stamps, stamps + n);
    const kicp::SweepTiming t = kicp::process_timestamps(v, header_stamp, *last_processed);
    for (int64_t i = 0; i < n; ++i) stamps[i] = v[static_cast<size_t>(i)];
    begin_end2[0] = t.begin, begin_end2[1] = t.end;
}
// poses7[n][7] + stamps[n] -> TUM file (offline_node.cpp:76-97)
int kfac_write_tum(const char *path, const double *stamps, const double *poses7, int64_t n) {
    std::vector<std::pair<double, Sophus::SE3d>> v;
    v.reserve(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) v.emplace_back(stamps[i], kicp::from_pose7(poses7 + 7 * i));
    return kicp::write_poses_tum(path, v) ? 0 : 1;
}
}
