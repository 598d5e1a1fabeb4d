// Host-side emulation of the registration kernel (kinematic-icp_b200/csrc/kicp_register.cu): the kernel SOURCE is compiled here
// unchanged against tests/emu/cuda_emu.hpp and run as a small grid of fibers — one grid per "rank" for the sharded path, the
// peers' mailboxes being plain host memory — on a map laid out exactly as the device holds it (kicp_internal.h).
// Test infrastructure: tests/test_kernel_emu_cpu.py compares the results with the CPU oracle.  Never part of the product library.
#define EMU_CHAOS 1  // interleaving stress on request (KICP_EMU_CHAOS = seed): this kernel waits through collectives and sleeps only
#include "cuda_emu.hpp"

#define KR_EMU 1
// the few PTX helpers of the kernel file, host versions
static inline unsigned long long gtime_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
static inline uint32_t ld_acquire_sys_u32(const uint32_t *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline unsigned int ld_acquire_gpu_u32(const unsigned int *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
struct Point4;
template <class P = Point4>
static inline P kr_emu_ld_point(const double *p) {
    P r;
    r.x = p[0], r.y = p[1], r.z = p[2], r.w = p[3];
    return r;
}
#define ld_point(p) kr_emu_ld_point(p)

#include "../../kinematic-icp_b200/csrc/kicp_register.cu"


namespace {
struct HostMap {
    std::vector<int4> slots;
    std::vector<double> pts;
    MapView view;
};
// the voxels laid out as the device map holds them: open addressing, linear probing, meta = block << 8 | count; 32-byte points
void build_map(HostMap &m, const int32_t *keys, const int32_t *counts, const double *pts, int64_t nvox, int32_t cap, double voxel_size) {
    uint32_t nslots = 1024;
    while ((int64_t)nslots < 4 * nvox) nslots <<= 1;
    m.slots.assign(nslots, make_int4(0, 0, 0, (int)KICP_SLOT_EMPTY));
    m.pts.assign((size_t)std::max<int64_t>(nvox, 1) * cap * KICP_PSTRIDE, 0.0);
    int64_t off = 0;
    for (int64_t b = 0; b < nvox; ++b) {
        uint32_t h = voxel_hash(keys[3 * b], keys[3 * b + 1], keys[3 * b + 2]) & (nslots - 1);
        while ((uint32_t)m.slots[h].w != KICP_SLOT_EMPTY) h = (h + 1) & (nslots - 1);
        m.slots[h] = make_int4(keys[3 * b], keys[3 * b + 1], keys[3 * b + 2], (int)(((uint32_t)b << 8) | (uint32_t)counts[b]));
        for (int j = 0; j < counts[b]; ++j)
            for (int d = 0; d < 3; ++d) m.pts[((size_t)b * cap + j) * KICP_PSTRIDE + d] = pts[(off + j) * 3 + d];
        off += counts[b];
    }
    m.view = MapView{m.slots.data(), nslots - 1, m.pts.data(), cap, voxel_size};
}
struct RankScratch {
    RegState st;
    std::vector<double> partials;
    std::vector<unsigned> nn_g, nn_g2, todo;
    std::vector<float> nn_l, nn_seed;
};
}  // namespace

// One registration of a frame split over `nranks` emulated GPUs (contiguous index ranges, as kb.shard_range does), every rank a
// grid of `grid` CTAs.  persistent = 1: k_register<true> (nranks > 1: the fused peer-mailbox exchange); persistent = 0 (nranks == 1
// only): one launch per pass, k_reg_init / k_register<false> / k_solve.  nn_cache = 1: certificates carried between passes.
// results: one kicp_reg_result per rank.  stats: [3] probes, candidate points, lines of rank 0.  Returns 0, or a negative number
// when a launch did not leave its counters as the next one needs them.
// late_upload = 1 (persistent, one rank): the kernel is started on a frame buffer full of NaN; an "uploader" thread then copies the
// frame in, segment by segment with pauses, raising each segment's flag after its bytes — the protocol of the host-pointer entry
// points (kicp_register_api.cu: the chunks are issued right after the launch).  A window that read its segment early would see NaN.
extern "C" int kr_emu_register(const int32_t *keys, const int32_t *counts, const double *pts, int64_t nvox, int32_t cap, double voxel_size,
                               const void *scan, int64_t n, int32_t f32, const double last[7], const double odom[7], double tau,
                               const kicp_reg_params *params, int32_t grid, int32_t nranks, int32_t persistent, int32_t nn_cache,
                               int32_t registrations, kicp_reg_result *results, uint64_t *stats, int32_t late_upload, int32_t device_count) {
    static int device_count_word;  // (read by the kernels through a pointer, like the survivor count a frame leaves in device memory)
    device_count_word = device_count;
    if (nranks < 1 || nranks > KICP_MAX_RANKS || (nranks > 1 && !persistent)) return -10;
    HostMap map;
    build_map(map, keys, counts, pts, nvox, cap, voxel_size);
    std::vector<RankScratch> rs(nranks);
    std::vector<P2PMailbox> boxes(nranks);
    memset(boxes.data(), 0, sizeof(P2PMailbox) * nranks);
    const int stride = f32 ? 12 : 24;
    int rc = 0;
    for (int reg = 0; reg < registrations; ++reg) {  // (several in a row: the counters a launch leaves behind, the mailbox parity and tags)
        std::vector<std::thread> ranks;
        std::vector<std::unique_ptr<std::vector<unsigned char>>> late_bufs;  // (kept alive until the threads have joined)
        std::vector<std::unique_ptr<std::vector<uint32_t>>> late_flags;
        for (int r = 0; r < nranks; ++r) {
            const int64_t lo = n * r / nranks, hi = n * (r + 1) / nranks;
            RankScratch &s = rs[r];
            if (reg == 0) {
                memset(&s.st, 0, sizeof(RegState));
                s.partials.assign((size_t)2 * grid * 8, 0.0);
                const size_t m = (size_t)std::max<int64_t>(hi - lo, 1);
                s.nn_g.assign(m, 0x12345678u), s.nn_g2.assign(m, 0x12345678u), s.todo.assign(m, 0u), s.nn_l.assign(m, 0.f), s.nn_seed.assign(m, 0.f);
            }
            KernelArgs a{};
            a.st = &s.st;
            a.scan.base = (const unsigned char *)scan + lo * stride, a.scan.n = (int)(hi - lo), a.scan.d_n = nullptr;
            if (device_count >= 0 && nranks == 1) a.scan.d_n = &device_count_word;  // the frame path: n is only an upper bound
            a.scan.stride = stride, a.scan.ox = 0, a.scan.oy = f32 ? 4 : 8, a.scan.oz = f32 ? 8 : 16, a.scan.f32 = f32;
            a.map = map.view, a.partials = s.partials.data();
            a.px.nranks = 1;
            if (nranks > 1) {
                for (int q = 0; q < nranks; ++q) a.px.peer[q] = &boxes[q];
                a.px.nranks = nranks, a.px.rank = r, a.px.parity = reg & 1, a.px.tag_base = (uint32_t)(reg * KICP_MAX_ITERATIONS + 1);
            }
            a.up = UploadArgs{nullptr, 0u, 1};
            std::vector<unsigned char> *late = nullptr;
            std::vector<uint32_t> *flags = nullptr;
            int64_t wpc = 1;
            if (late_upload && persistent && nranks == 1 && hi > lo) {
                const int64_t windows = (hi - lo + 31) / 32;
                wpc = (windows + KICP_UPLOAD_CHUNKS - 1) / KICP_UPLOAD_CHUNKS;  // as register_host computes it
                late_bufs.emplace_back(new std::vector<unsigned char>((size_t)(hi - lo) * stride, 0xFF));  // all-ones bytes: NaN in float32 and float64
                late_flags.emplace_back(new std::vector<uint32_t>(KICP_UPLOAD_CHUNKS, 0u));
                late = late_bufs.back().get(), flags = late_flags.back().get();
                a.scan.base = late->data();
                a.up = UploadArgs{flags->data(), (uint32_t)(reg + 1), (int)wpc};
            }
            a.init.last = Pose{last[0], last[1], last[2], last[3], last[4], last[5], last[6]};
            a.init.odom = Pose{odom[0], odom[1], odom[2], odom[3], odom[4], odom[5], odom[6]};
            a.init.tau = tau, a.init.conv = params->convergence_criterion, a.init.fixed_reg = params->fixed_regularization;
            a.init.adaptive = params->use_adaptive_odometry_regularization ? 1 : 0, a.init.max_iter = params->max_num_iterations,
            a.init.iters_out = nullptr;
            int e = 0;
            a.pow2_voxel = std::frexp(voxel_size, &e) == 0.5 ? 1 : 0;
            a.collect_stats = 1;
            if (nn_cache && persistent)
                a.nn_g = s.nn_g.data(), a.nn_g2 = s.nn_g2.data(), a.nn_l = s.nn_l.data(), a.nn_seed = s.nn_seed.data(), a.todo = s.todo.data();
            a.result_host = nullptr, a.timeout_ns = 600ull * 1000000000ull;
            memset(s.st.stats, 0, sizeof(s.st.stats));
            if (persistent) {
                // (ranks are concurrent grids: emu::launch keeps one global grid size, identical for all of them)
                ranks.emplace_back([a, grid]() { emu::launch(grid, KR_THREADS, [a]() { k_register<true>(a); }); });
                if (late) {  // the uploader: starts AFTER the launch, like the host-pointer entry points
                    const unsigned char *src = (const unsigned char *)scan + lo * stride;
                    const int64_t npts = hi - lo;
                    const uint32_t seq = (uint32_t)(reg + 1);
                    ranks.emplace_back([late, flags, src, npts, wpc, stride, seq]() {
                        for (int k = 0; k < KICP_UPLOAD_CHUNKS; ++k) {
                            std::this_thread::sleep_for(std::chrono::milliseconds(3));
                            const int64_t p0 = std::min<int64_t>(npts, (int64_t)k * wpc * 32), p1 = std::min<int64_t>(npts, (int64_t)(k + 1) * wpc * 32);
                            if (p1 > p0) memcpy(late->data() + p0 * stride, src + p0 * stride, (size_t)(p1 - p0) * stride);
                            __atomic_store_n(&(*flags)[k], seq, __ATOMIC_RELEASE);
                        }
                    });
                }
            } else {
                RegState *st = &s.st;
                emu::launch(1, 32, [st, a]() { k_reg_init(st, a.init); });
                for (int j = 0; j < a.init.max_iter; ++j) {
                    emu::launch(grid, KR_THREADS, [a]() { k_register<false>(a); });
                    emu::launch(1, 32, [st]() { k_solve(st); });
                }
            }
        }
        for (auto &t : ranks) t.join();
        for (int r = 0; r < nranks; ++r) {
            results[r] = rs[r].st.result;
            const RegState &st = rs[r].st;
            if (persistent && (st.win_ctr || st.arrive || st.exit_ctr || st.a_arrive || st.abort)) rc = -1;
            for (int k = 0; k < KICP_MAX_ITERATIONS; ++k)
                if (persistent && (st.todo_n[k] || st.slow_n[k])) rc = -2;
        }
    }
    if (stats) stats[0] = rs[0].st.stats[0], stats[1] = rs[0].st.stats[1], stats[2] = rs[0].st.stats[2];
    return rc;
}
