// Host-side emulation of the voxel-map kernels (kinematic-icp_b200/csrc/kicp_map_kernels.cuh): the kernel SOURCE is compiled here
// unchanged against tests/emu/cuda_emu.hpp; the launch sequences of kicp_map.cu (map_add_points_impl, kicp_map_remove_far,
// kicp_map_load_voxels, kicp_map_nearest) are restated with host memory in place of the slab; the renumbering of the surviving blocks
// runs the product's own single-pass scan (k_exclusive_sum, kicp_scan.cuh) with its state kept across launches.  Test infrastructure: tests/test_map_kernels_emu_cpu.py compares the map with the CPU oracle bit for bit.
// Build with -ffp-contract=off (the product compiles these kernels with -fmad=false).  Never part of the product library.
#include "cuda_emu.hpp"

#include "../../kinematic-icp_b200/csrc/kicp_map_kernels.cuh"
#include "scan_state.hpp"

namespace {
struct EmuMap {
    double voxel_size, max_distance;
    int cap;
    uint32_t blocks_cap, nslots, num_blocks = 0;
    int64_t num_points = 0;
    std::vector<int4> slots, blk, blk_spare;
    std::vector<double> pts, pts_spare;
    std::vector<int32_t> pend_head;
    MapRW rw() { return MapRW{slots.data(), nslots - 1, blk.data(), pts.data(), pend_head.data(), blocks_cap, cap, voxel_size}; }
    MapView view() const { return MapView{slots.data(), nslots - 1, pts.data(), cap, voxel_size}; }
};
uint32_t pow2_at_least(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t)p;
}
void rebuild_table(EmuMap &m) {  // map_rebuild_table: clear, then one thread per block claims its slot
    std::fill(m.slots.begin(), m.slots.end(), make_int4(-1, -1, -1, (int)KICP_SLOT_EMPTY));
    if (m.num_blocks) {
        int4 *slots = m.slots.data();
        const int4 *blk = m.blk.data();
        const uint32_t mask = m.nslots - 1, nb = m.num_blocks;
        emu::launch_waves((int)((nb + 255) / 256), 256, [=]() { k_table_rebuild(slots, mask, blk, nb); });
    }
}
// enqueue_exclusive_sum of kicp_map.cu
void exclusive_sum(const uint32_t *in, uint32_t *out, uint32_t n) {
    if (n == 0) return;
    const kicp_scan_args sa = emu::scan_state().next(n, kScanTile);
    emu::launch_waves((int)((n + kScanTile - 1) / kScanTile), kScanThreads, [=]() { k_exclusive_sum(in, out, n, sa); });
}
}  // namespace

extern "C" {
// the scan alone (tests/test_map_kernels_emu_cpu.py: sizes around the tile and look-back boundaries, many launches on one state)
void km_emu_exclusive_sum(const uint32_t *in, uint32_t *out, uint32_t n) { exclusive_sum(in, out, n); }

void *km_emu_create(double voxel_size, double max_distance, int32_t cap, uint32_t blocks_cap) {
    EmuMap *m = new EmuMap();
    m->voxel_size = voxel_size, m->max_distance = max_distance, m->cap = cap, m->blocks_cap = blocks_cap;
    m->nslots = std::max<uint32_t>(pow2_at_least((uint64_t)blocks_cap * 4), 1024u);  // load <= 0.25, as map_alloc_storage
    m->slots.assign(m->nslots, make_int4(-1, -1, -1, (int)KICP_SLOT_EMPTY));
    m->blk.assign(blocks_cap, make_int4(0, 0, 0, 0)), m->blk_spare = m->blk;
    m->pts.assign((size_t)blocks_cap * cap * KICP_PSTRIDE, 0.0), m->pts_spare = m->pts;
    m->pend_head.assign(blocks_cap, -1);
    return m;
}
void km_emu_destroy(void *h) { delete static_cast<EmuMap *>(h); }
int64_t km_emu_num_points(void *h) { return static_cast<EmuMap *>(h)->num_points; }
int64_t km_emu_num_voxels(void *h) { return static_cast<EmuMap *>(h)->num_blocks; }

// VoxelHashMap::AddPoints (pose7 == null) or the insert half of Update(points, pose): map_add_points_impl
int km_emu_add_points(void *h, const double *xyz, int64_t n, const double *pose7) {
    EmuMap &m = *static_cast<EmuMap *>(h);
    if (n == 0) return 0;
    if ((uint64_t)m.num_blocks + (uint64_t)n > m.blocks_cap) return -1;  // (the product grows its slab here)
    std::vector<double> xyz_t((size_t)n * 3);
    std::vector<int32_t> next((size_t)n), touched((size_t)n);
    uint32_t counters[8] = {m.num_blocks, 0, 0, 0, 0, 0, 0, 0};
    MapRW rw = m.rw();
    Pose pose{0, 0, 0, 1, 0, 0, 0};
    if (pose7) pose = Pose{pose7[0], pose7[1], pose7[2], pose7[3], pose7[4], pose7[5], pose7[6]};
    const int has_pose = pose7 ? 1 : 0;
    double *xt = xyz_t.data();
    int32_t *nx = next.data(), *tc = touched.data();
    uint32_t *ctr = counters;
    const int grid = (int)((n + 255) / 256);
    emu::launch_waves(grid, 256, [=]() { k_add_find_or_create(rw, xyz, n, has_pose, pose, xt, nx, ctr, tc); });
    const double map_resolution = std::sqrt(m.voxel_size * m.voxel_size / (double)m.cap);
    emu::launch_waves(grid, 256, [=]() { k_add_commit(rw, xt, nx, ctr, tc, map_resolution); });
    if (counters[2]) return -2;
    m.num_blocks = counters[0];
    m.num_points += counters[3];
    return 0;
}

// VoxelHashMap::RemovePointsFarFromLocation: kicp_map_remove_far
int km_emu_remove_far(void *h, const double origin[3]) {
    EmuMap &m = *static_cast<EmuMap *>(h);
    if (m.num_blocks == 0) return 0;
    std::vector<uint32_t> keep(m.num_blocks), new_id(m.num_blocks);
    uint32_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    {
        const int4 *blk = m.blk.data();
        const double *pts = m.pts.data();
        uint32_t *kp = keep.data(), *ctr = counters;
        const uint32_t nb = m.num_blocks;
        const int cap = m.cap;
        const double ox = origin[0], oy = origin[1], oz = origin[2], md2 = m.max_distance * m.max_distance;
        emu::launch_waves((int)((nb + 255) / 256), 256, [=]() { k_mark_far(blk, pts, cap, nb, ox, oy, oz, md2, kp, ctr); });
    }
    if (counters[4] > 0) {
        exclusive_sum(keep.data(), new_id.data(), m.num_blocks);
        const int4 *blk = m.blk.data();
        const double *pts = m.pts.data();
        const uint32_t *kp = keep.data(), *ni = new_id.data();
        int4 *bo = m.blk_spare.data();
        double *po = m.pts_spare.data();
        const uint32_t nb = m.num_blocks;
        const int cap = m.cap;
        emu::launch_waves((int)(((uint64_t)nb * 32 + 255) / 256), 256, [=]() { k_compact_blocks(blk, pts, cap, nb, kp, ni, bo, po); });
        std::swap(m.blk, m.blk_spare);
        std::swap(m.pts, m.pts_spare);
        m.num_blocks -= counters[4];
        m.num_points -= counters[3];
        rebuild_table(m);
    }
    return 0;
}

// VoxelHashMap::Update(points, pose) as kicp_register_frame issues it (kicp_map_update_pose_async): the point count and the pose are
// READ BY THE KERNELS from device-resident words (`d_n`, the registration's result block), nothing comes back to the host in between;
// the eviction always compacts into the spare arrays, over an upper bound of the block count.  status != KICP_OK in the result block
// (a NaN pose) must leave the map untouched.
int km_emu_update_pose_async(void *h, const double *xyz, int64_t n_max, int32_t n_actual, const double pose7[7], int32_t status) {
    EmuMap &m = *static_cast<EmuMap *>(h);
    if (n_max <= 0) return -3;
    if ((uint64_t)m.num_blocks + (uint64_t)n_max > m.blocks_cap) return -1;
    std::vector<double> xyz_t((size_t)n_max * 3);
    std::vector<int32_t> next((size_t)n_max), touched((size_t)n_max);
    uint32_t counters[8] = {m.num_blocks, 0, 0, 0, 0, 0, 0, 0};
    const int d_n_word = n_actual;
    kicp_reg_result res;
    memset(&res, 0, sizeof(res));
    for (int k = 0; k < 7; ++k) res.pose[k] = pose7[k];
    res.status = status;
    MapRW rw = m.rw();
    double *xt = xyz_t.data();
    int32_t *nx = next.data(), *tc = touched.data();
    uint32_t *ctr = counters;
    const int *d_n = &d_n_word;
    const kicp_reg_result *d_res = &res;
    const int pgrid = (int)((n_max + 255) / 256);
    emu::launch_waves(pgrid, 256, [=]() { k_add_find_or_create(rw, xyz, n_max, 1, Pose{0, 0, 0, 1, 0, 0, 0}, xt, nx, ctr, tc, d_n, d_res); });
    const double map_resolution = std::sqrt(m.voxel_size * m.voxel_size / (double)m.cap);
    emu::launch_waves(pgrid, 256, [=]() { k_add_commit(rw, xt, nx, ctr, tc, map_resolution); });
    const uint32_t ub = (uint32_t)std::min<uint64_t>((uint64_t)m.num_blocks + (uint64_t)n_max, m.blocks_cap);
    std::vector<uint32_t> keep(ub), new_id(ub);
    {
        const int4 *blk = m.blk.data();
        const double *pts = m.pts.data();
        uint32_t *kp = keep.data();
        const int cap = m.cap;
        const double md2 = m.max_distance * m.max_distance;
        emu::launch_waves((int)((ub + 255) / 256), 256, [=]() { k_mark_far(blk, pts, cap, ub, 0.0, 0.0, 0.0, md2, kp, ctr, d_res); });
        exclusive_sum(keep.data(), new_id.data(), ub);
        const uint32_t *ni = new_id.data();
        int4 *bo = m.blk_spare.data();
        double *po = m.pts_spare.data();
        emu::launch_waves((int)(((uint64_t)ub * 32 + 255) / 256), 256, [=]() { k_compact_blocks(blk, pts, cap, ub, kp, ni, bo, po); });
    }
    std::swap(m.blk, m.blk_spare);
    std::swap(m.pts, m.pts_spare);
    std::fill(m.slots.begin(), m.slots.end(), make_int4(-1, -1, -1, (int)KICP_SLOT_EMPTY));
    {
        int4 *slots = m.slots.data();
        const int4 *blk = m.blk.data();
        const uint32_t mask = m.nslots - 1;
        emu::launch_waves((int)((ub + 255) / 256), 256, [=]() { k_table_rebuild(slots, mask, blk, ub, ctr); });
    }
    // kicp_map_finish_update
    m.num_blocks = counters[0] - counters[4];
    m.num_points += (int64_t)counters[3] - (int64_t)counters[5];
    return counters[2] ? -2 : 0;
}

// kicp_map_load_voxels: voxel v becomes block v, then the table is rebuilt
int km_emu_load_voxels(void *h, const int32_t *keys, const int32_t *counts, const double *points, int64_t nvox) {
    EmuMap &m = *static_cast<EmuMap *>(h);
    if ((uint64_t)nvox > m.blocks_cap) return -1;
    std::vector<int64_t> offsets((size_t)nvox + 1, 0);
    for (int64_t v = 0; v < nvox; ++v) offsets[v + 1] = offsets[v] + counts[v];
    int4 *blk = m.blk.data();
    double *pts = m.pts.data();
    const int64_t *off = offsets.data();
    const int cap = m.cap;
    const uint32_t nv = (uint32_t)nvox;
    if (nvox) emu::launch_waves((int)(((uint64_t)nv * 32 + 255) / 256), 256, [=]() { k_load_voxels(blk, pts, cap, keys, counts, off, points, nv); });
    m.num_blocks = nv, m.num_points = offsets[nvox];
    rebuild_table(m);
    return 0;
}

// the map as (keys, counts, points grouped by voxel) in block order; returns the number of voxels
int64_t km_emu_export(void *h, int32_t *keys, int32_t *counts, double *points) {
    EmuMap &m = *static_cast<EmuMap *>(h);
    int64_t off = 0;
    for (uint32_t b = 0; b < m.num_blocks; ++b) {
        keys[3 * b] = m.blk[b].x, keys[3 * b + 1] = m.blk[b].y, keys[3 * b + 2] = m.blk[b].z, counts[b] = m.blk[b].w;
        for (int j = 0; j < m.blk[b].w; ++j, ++off)
            for (int d = 0; d < 3; ++d) points[off * 3 + d] = m.pts[((size_t)b * m.cap + j) * KICP_PSTRIDE + d];
    }
    return m.num_blocks;
}

// GetClosestNeighbor for a batch: kicp_map_nearest
int km_emu_nearest(void *h, const double *q, int64_t n, double *out_pts, double *out_dist) {
    EmuMap &m = *static_cast<EmuMap *>(h);
    const MapView v = m.view();
    if (n) emu::launch_waves((int)((n + 255) / 256), 256, [=]() { k_nearest(v, q, n, out_pts, out_dist); });
    return 0;
}
}
