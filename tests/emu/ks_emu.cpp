// Host-side emulation of the voxel-sorted registration kernel (kinematic-icp_b200/csrc/kicp_register_sorted.cu): the kernel
// SOURCE is compiled here unchanged against tests/emu/cuda_emu.hpp and run as a small grid of fibers, on a map laid out exactly
// as the device holds it (kicp_internal.h).  Test infrastructure: tests/test_sorted_engine_emu_cpu.py compares its result with
// the CPU oracle.  Never part of the product library.
#include "cuda_emu.hpp"

#define KS_EMU 1
// warp-level work model: [0] shift-loop steps, [1] candidate-loop steps (2 points each), [2] lane visits, [3] candidate points; +4: passes >= 1
static unsigned long long ks_emu_counters[8];
#define KS_COUNT(counter, value) __atomic_fetch_add(&ks_emu_counters[counter], (unsigned long long)(value), __ATOMIC_RELAXED);
// the few PTX helpers of the kernel file, host versions
struct KsPoint;
static inline unsigned long long ks_gtime_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
static inline unsigned int ks_ld_acquire_gpu(const unsigned int *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline uint32_t ks_ld_acquire_sys(const uint32_t *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
#define ks_ld_map_point(p) ks_emu_ld_point(p)
#define ks_ld_frame_point(p) ks_emu_ld_point(p)
#define ks_st_frame_point(dst, x, y, z) ((dst)[0] = (x), (dst)[1] = (y), (dst)[2] = (z), (dst)[3] = 0.0)
template <class P = KsPoint>
static inline P ks_emu_ld_point(const double *p) {
    P r;
    r.x = p[0], r.y = p[1], r.z = p[2], r.w = p[3];
    return r;
}

#include "../../kinematic-icp_b200/csrc/kicp_register_sorted.cu"

#include <vector>

// Lay the voxels out as the device map does (open addressing, linear probing, meta = block << 8 | count; 32-byte points) and run
// one registration on `grid` CTAs.  keys [nvox][3] int32, counts [nvox] int32, pts [sum counts][3] doubles grouped by voxel.
extern "C" int ks_emu_register(const int32_t *keys, const int32_t *counts, const double *pts, int64_t nvox, int32_t cap, double voxel_size,
                               const void *scan, int64_t n, int32_t f32, const double last[7], const double odom[7], double tau,
                               const kicp_reg_params *params, int32_t grid, kicp_reg_result *result, double *sorted_out /* [n][4] or null */,
                               uint64_t *stats_out /* [11] probes, candidate points, 128-byte lines, the 8 work-model counters; or null */) {
    uint32_t nslots = 1024;
    while ((int64_t)nslots < 4 * nvox) nslots <<= 1;
    std::vector<int4> slots(nslots, make_int4(0, 0, 0, (int)KICP_SLOT_EMPTY));
    std::vector<double> mp((size_t)std::max<int64_t>(nvox, 1) * cap * KICP_PSTRIDE, 0.0);
    int64_t off = 0;
    for (int64_t b = 0; b < nvox; ++b) {
        uint32_t h = voxel_hash(keys[3 * b], keys[3 * b + 1], keys[3 * b + 2]) & (nslots - 1);
        while ((uint32_t)slots[h].w != KICP_SLOT_EMPTY) h = (h + 1) & (nslots - 1);
        slots[h] = make_int4(keys[3 * b], keys[3 * b + 1], keys[3 * b + 2], (int)(((uint32_t)b << 8) | (uint32_t)counts[b]));
        for (int j = 0; j < counts[b]; ++j)
            for (int d = 0; d < 3; ++d) mp[((size_t)b * cap + j) * KICP_PSTRIDE + d] = pts[(off + j) * 3 + d];
        off += counts[b];
    }
    SortedArgs a{};
    SortedState st{};
    std::vector<double> dbg(KICP_MAX_ITERATIONS * 6, 0.0);
    unsigned long long stats[4] = {0, 0, 0, 0};
    a.st = &st, a.result = result, a.dbg = dbg.data(), a.stats = stats;
    a.scan.base = (const unsigned char *)scan, a.scan.n = (int)n, a.scan.d_n = nullptr;
    a.scan.stride = f32 ? 12 : 24, a.scan.ox = 0, a.scan.oy = f32 ? 4 : 8, a.scan.oz = f32 ? 8 : 16, a.scan.f32 = f32;
    a.map = MapView{slots.data(), nslots - 1, mp.data(), cap, voxel_size};
    std::vector<double> partials((size_t)2 * grid * 8, 0.0);
    a.partials = partials.data();
    a.up = UploadArgs{nullptr, 0u, 1};
    a.init.last = Pose{last[0], last[1], last[2], last[3], last[4], last[5], last[6]};
    a.init.odom = Pose{odom[0], odom[1], odom[2], odom[3], odom[4], odom[5], odom[6]};
    a.init.tau = tau, a.init.conv = params->convergence_criterion, a.init.fixed_reg = params->fixed_regularization;
    a.init.adaptive = params->use_adaptive_odometry_regularization ? 1 : 0, a.init.max_iter = params->max_num_iterations, a.init.iters_out = nullptr;
    int e = 0;
    a.pow2_voxel = std::frexp(voxel_size, &e) == 0.5 ? 1 : 0;
    a.collect_stats = 1;
    uint32_t bslots = 1024;
    while ((int64_t)bslots < 2 * n) bslots <<= 1;
    std::vector<unsigned long long> bin_key(bslots, 0xFFFFFFFFFFFFFFFFull);
    std::vector<unsigned int> bin_cnt(bslots, 0u);
    std::vector<uint2> pslot((size_t)std::max<int64_t>(n, 1));
    std::vector<double> sorted((size_t)std::max<int64_t>(n, 1) * 4, -1.0);
    std::vector<unsigned int> nn((size_t)std::max<int64_t>(n, 1), 0x12345678u);
    a.bin_key = bin_key.data(), a.bin_cnt = bin_cnt.data(), a.bin_mask = bslots - 1, a.pslot = pslot.data(), a.sorted = sorted.data(),
    a.nn_g = nn.data();
    a.result_host = nullptr, a.timeout_ns = 600ull * 1000000000ull;
    emu::launch(grid, KS_THREADS, [a]() { k_register_sorted<0>(a); });
    if (stats_out) for (int k = 0; k < 8; ++k) stats_out[3 + k] = ks_emu_counters[k], ks_emu_counters[k] = 0;
    if (stats_out) stats_out[0] = stats[0], stats_out[1] = stats[1], stats_out[2] = stats[2];
    if (sorted_out) memcpy(sorted_out, sorted.data(), (size_t)n * 4 * sizeof(double));
    // what every launch must leave behind for the next one: an empty table, zero counts, zero counters
    for (uint32_t s = 0; s < bslots; ++s)
        if (bin_key[s] != 0xFFFFFFFFFFFFFFFFull || bin_cnt[s] != 0u) return -1;
    if (st.win_ctr || st.arrive || st.exit_ctr || st.cursor || st.abort) return -2;
    return 0;
}
