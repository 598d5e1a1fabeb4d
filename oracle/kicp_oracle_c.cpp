// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  extern "C" surface of the CPU oracle for ctypes
// (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).  All poses are double[7] =
// {qx,qy,qz,qw,tx,ty,tz}; all point arrays are row-major xyz doubles.
#include <algorithm>
#include <cstring>
#include <vector>

#include "kicp_oracle.hpp"

using namespace kicp_oracle;

namespace {
std::vector<Vec3> to_vec(const double *xyz, int64_t n) {
    std::vector<Vec3> v(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) v[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    return v;
}
void from_vec(const std::vector<Vec3> &v, double *xyz) {
    for (size_t i = 0; i < v.size(); ++i) xyz[3 * i] = v[i].x, xyz[3 * i + 1] = v[i].y, xyz[3 * i + 2] = v[i].z;
}
}  // namespace

extern "C" {

struct kor_stats {
    int32_t iterations;
    int32_t associations;
    double beta;
    double last_dx_norm;
    double sums[64][7];  // JTJ00 JTJ01 JTJ11 JTr0 JTr1 N sumsq (sumsq only on solve 0)
    double dx[64][2];
};

void *kor_map_create(double voxel_size, double max_distance, unsigned max_points_per_voxel) {
    return new VoxelHashMap(voxel_size, max_distance, max_points_per_voxel);
}
void kor_map_destroy(void *h) { delete static_cast<VoxelHashMap *>(h); }
void kor_map_clear(void *h) { static_cast<VoxelHashMap *>(h)->Clear(); }
int kor_map_empty(void *h) { return static_cast<VoxelHashMap *>(h)->Empty() ? 1 : 0; }
int64_t kor_map_num_points(void *h) { return static_cast<int64_t>(static_cast<VoxelHashMap *>(h)->NumPoints()); }
int64_t kor_map_num_voxels(void *h) { return static_cast<int64_t>(static_cast<VoxelHashMap *>(h)->map_.size()); }
void kor_map_add_points(void *h, const double *xyz, int64_t n) { static_cast<VoxelHashMap *>(h)->AddPoints(to_vec(xyz, n)); }
void kor_map_remove_far(void *h, const double *origin) {
    static_cast<VoxelHashMap *>(h)->RemovePointsFarFromLocation({origin[0], origin[1], origin[2]});
}
void kor_map_update_origin(void *h, const double *xyz, int64_t n, const double *origin) {
    static_cast<VoxelHashMap *>(h)->Update(to_vec(xyz, n), Vec3{origin[0], origin[1], origin[2]});
}
void kor_map_update_pose(void *h, const double *xyz, int64_t n, const double *pose7) {
    static_cast<VoxelHashMap *>(h)->Update(to_vec(xyz, n), SE3::from_pose7(pose7));
}
int64_t kor_map_pointcloud(void *h, double *out, int64_t cap) {
    const auto pts = static_cast<VoxelHashMap *>(h)->Pointcloud();
    if (static_cast<int64_t>(pts.size()) > cap) return -static_cast<int64_t>(pts.size());
    from_vec(pts, out);
    return static_cast<int64_t>(pts.size());
}
// Voxel-grouped export, voxels sorted lexicographically by key so the layout is reproducible:
// keys[V][3], counts[V], points[total][3] (each voxel's points in insertion order).
int64_t kor_map_export_voxels(void *h, int32_t *keys, int32_t *counts, double *points) {
    auto *m = static_cast<VoxelHashMap *>(h);
    std::vector<const std::pair<const Voxel, std::vector<Vec3>> *> items;
    items.reserve(m->map_.size());
    for (const auto &kv : m->map_) items.push_back(&kv);
    std::sort(items.begin(), items.end(), [](auto *a, auto *b) {
        if (a->first.x != b->first.x) return a->first.x < b->first.x;
        if (a->first.y != b->first.y) return a->first.y < b->first.y;
        return a->first.z < b->first.z;
    });
    int64_t v = 0, p = 0;
    for (auto *it : items) {
        keys[3 * v] = it->first.x, keys[3 * v + 1] = it->first.y, keys[3 * v + 2] = it->first.z;
        counts[v] = static_cast<int32_t>(it->second.size());
        for (const Vec3 &q : it->second) points[3 * p] = q.x, points[3 * p + 1] = q.y, points[3 * p + 2] = q.z, ++p;
        ++v;
    }
    return v;
}
void kor_map_nearest(void *h, const double *q, int64_t n, double *out_pts, double *out_dist) {
    auto *m = static_cast<VoxelHashMap *>(h);
    for (int64_t i = 0; i < n; ++i) {
        const auto [p, d] = m->GetClosestNeighbor({q[3 * i], q[3 * i + 1], q[3 * i + 2]});
        out_pts[3 * i] = p.x, out_pts[3 * i + 1] = p.y, out_pts[3 * i + 2] = p.z;
        out_dist[i] = d;
    }
}
// mean candidates per query in the occupied voxels of the 27-neighbourhood (c-bar of SURVEY.md §8(d)),
// evaluated at q = T * p; also returns mean occupied voxels.
void kor_map_neighbourhood_stats(void *h, const double *xyz, int64_t n, const double *pose7, double *mean_candidates,
                                 double *mean_occupied) {
    auto *m = static_cast<VoxelHashMap *>(h);
    const SE3 T = SE3::from_pose7(pose7);
    double cand = 0, occ = 0;
    for (int64_t i = 0; i < n; ++i) {
        const Vec3 q = transform(T, {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
        const Voxel v = PointToVoxel(q, m->voxel_size_);
        for (const Voxel &s : kVoxelShifts) {
            auto it = m->map_.find({v.x + s.x, v.y + s.y, v.z + s.z});
            if (it != m->map_.end()) cand += it->second.size(), occ += 1;
        }
    }
    *mean_candidates = n ? cand / n : 0;
    *mean_occupied = n ? occ / n : 0;
}

void kor_register(void *h, const double *frame, int64_t n, const double *last_pose7, const double *rel_odom7, double tau,
                  int max_iter, double conv, int adaptive, double fixed_reg, int threads, double *out_pose7,
                  kor_stats *stats) {
    RegParams p;
    p.max_num_iterations = max_iter;
    p.convergence_criterion = conv;
    p.use_adaptive_odometry_regularization = adaptive != 0;
    p.fixed_regularization = fixed_reg;
    RegStats st;
    const SE3 T = ComputeRobotMotion(to_vec(frame, n), *static_cast<VoxelHashMap *>(h), SE3::from_pose7(last_pose7),
                                     SE3::from_pose7(rel_odom7), tau, p, &st, threads);
    T.to_pose7(out_pose7);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->iterations = st.iterations;
        stats->associations = st.associations;
        stats->beta = st.beta;
        stats->last_dx_norm = st.last_dx_norm;
        for (size_t i = 0; i < st.sums.size() && i < 64; ++i) {
            const IterSums &s = st.sums[i];
            const double v[7] = {s.JTJ00, s.JTJ01, s.JTJ11, s.JTr0, s.JTr1, s.N, s.sumsq};
            std::memcpy(stats->sums[i], v, sizeof(v));
            stats->dx[i][0] = st.dx[2 * i], stats->dx[i][1] = st.dx[2 * i + 1];
        }
    }
}

// iteration order of the (recalled) tsl::robin_map over `n` distinct voxels inserted in the given order after reserve(reserve_count)
void kor_robin_order(const int32_t *keys, int64_t n, int64_t reserve_count, int mask20, int64_t *out) {
    std::vector<Voxel> k(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) k[i] = Voxel{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]};
    const auto order = RobinMapIterationOrder(k, static_cast<size_t>(reserve_count), mask20 != 0);
    for (size_t i = 0; i < order.size(); ++i) out[i] = static_cast<int64_t>(order[i]);
}
void kor_set_downsample_order(int mode) { SetDownsampleOrder(mode); }
int kor_get_downsample_order() { return GetDownsampleOrder(); }
int64_t kor_voxel_downsample(const double *xyz, int64_t n, double voxel_size, double *out) {
    const auto r = VoxelDownsample(to_vec(xyz, n), voxel_size);
    from_vec(r, out);
    return static_cast<int64_t>(r.size());
}
int64_t kor_preprocess(const double *xyz, int64_t n, const double *stamps, int64_t n_stamps, const double *rel_motion7,
                       double max_range, double min_range, int deskew, double *out) {
    std::vector<double> ts(stamps, stamps + n_stamps);
    const auto r = Preprocess(to_vec(xyz, n), ts, SE3::from_pose7(rel_motion7), max_range, min_range, deskew != 0);
    from_vec(r, out);
    return static_cast<int64_t>(r.size());
}

// SE3 helpers for the known-answer tests
void kor_se3_exp(const double *tangent6, double *pose7) { se3_exp(tangent6).to_pose7(pose7); }
void kor_se3_log(const double *pose7, double *tangent6) { se3_log(SE3::from_pose7(pose7), tangent6); }
void kor_se3_compose(const double *a7, const double *b7, double *out7) {
    compose(SE3::from_pose7(a7), SE3::from_pose7(b7)).to_pose7(out7);
}
void kor_se3_inverse(const double *a7, double *out7) { inverse(SE3::from_pose7(a7)).to_pose7(out7); }
void kor_se3_transform(const double *pose7, const double *xyz, int64_t n, double *out) {
    const SE3 T = SE3::from_pose7(pose7);
    for (int64_t i = 0; i < n; ++i) {
        const Vec3 p = transform(T, {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
        out[3 * i] = p.x, out[3 * i + 1] = p.y, out[3 * i + 2] = p.z;
    }
}

// CorrespondenceThreshold
void *kor_threshold_create(double map_err, double max_range, int adaptive, double fixed) {
    return new CorrespondenceThreshold(map_err, max_range, adaptive != 0, fixed);
}
void kor_threshold_destroy(void *h) { delete static_cast<CorrespondenceThreshold *>(h); }
void kor_threshold_update(void *h, const double *err7) {
    static_cast<CorrespondenceThreshold *>(h)->UpdateOdometryError(SE3::from_pose7(err7));
}
double kor_threshold_compute(void *h) { return static_cast<CorrespondenceThreshold *>(h)->ComputeThreshold(); }
void kor_threshold_reset(void *h) { static_cast<CorrespondenceThreshold *>(h)->Reset(); }

}  // extern "C"

// Diagnostic (tests / design analysis): work the exact-pruning traversal does per point — hash probes issued and
// candidate distances evaluated — with the same pruning rule as the CUDA kernel k_assoc_pruned.
extern "C" void kor_pruned_work(void *h, const double *xyz, int64_t n, const double *pose7, int32_t *probes, int32_t *cands,
                                int32_t *empty_centre) {
    auto *m = static_cast<VoxelHashMap *>(h);
    const SE3 T = SE3::from_pose7(pose7);
    const double vs = m->voxel_size_;
    for (int64_t i = 0; i < n; ++i) {
        const Vec3 q = transform(T, {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
        const Voxel v = PointToVoxel(q, vs);
        const double g[3][2] = {{(v.x + 1) * vs - q.x, q.x - v.x * vs}, {(v.y + 1) * vs - q.y, q.y - v.y * vs}, {(v.z + 1) * vs - q.z, q.z - v.z * vs}};
        double best = 1.7976931348623157e308;
        int np = 0, nc = 0;
        for (int k = 0; k < 27; ++k) {
            const Voxel &s = kVoxelShifts[k];
            const int sh[3] = {s.x, s.y, s.z};
            double lb2 = 0;
            for (int a = 0; a < 3; ++a) {
                const double d = sh[a] > 0 ? g[a][0] : (sh[a] < 0 ? g[a][1] : 0.0);
                lb2 += d * d;
            }
            if (lb2 > best * (1 + 1e-6) + 1e-10) continue;
            ++np;
            auto it = m->map_.find({v.x + s.x, v.y + s.y, v.z + s.z});
            if (it == m->map_.end()) {
                if (k == 0 && empty_centre) empty_centre[i] = 1;
                continue;
            }
            if (k == 0 && empty_centre) empty_centre[i] = 0;
            for (const Vec3 &p : it->second) {
                ++nc;
                const Vec3 d = p - q;
                const double d2 = d.x * d.x + d.y * d.y + d.z * d.z;
                if (d2 < best) best = d2;
            }
        }
        probes[i] = np, cands[i] = nc;
    }
}
