// Host-side emulation of the front-end kernels (kinematic-icp_b200/csrc/kicp_frontend_kernels.cuh): the kernel SOURCE is compiled
// here unchanged against tests/emu/cuda_emu.hpp; the launch sequences of kicp_frontend.cu (enqueue_downsample, enqueue_preprocess,
// the ingest of kicp_register_frame) are restated with host memory — the stable selects and the stamps' min / max included: they are
// the product's own kernels (kicp_scan.cuh fused into k_ds_select / k_preprocess_select, k_stamp_minmax), with the scan state kept
// across launches as a context keeps it.  Test infrastructure (tests/test_frontend_kernels_emu_cpu.py); build with
// -ffp-contract=off (the product compiles these kernels with -fmad=false).  Never part of the product library.
#include "cuda_emu.hpp"

#include "../../kinematic-icp_b200/csrc/kicp_frontend_kernels.cuh"
#include "scan_state.hpp"

extern "C" {
// kiss_icp::VoxelDownsample: enqueue_downsample (scratch hash, atomicMin of the input index per voxel, flag, stable select)
// n_actual >= 0: the stage runs as in a frame whose survivor count lives on the device — grid sized for the upper bound n, the
// kernels read the count from a device word (`d_n`) and the tail threads retire; -1: the count is the host's n.
int64_t kf_emu_voxel_downsample(const double *xyz, int64_t n, double voxel_size, double *out, int32_t n_actual) {
    if (n == 0) return 0;
    const int d_n_word = n_actual;
    const int *d_n = n_actual >= 0 ? &d_n_word : nullptr;
    uint32_t nslots = 1024;
    while ((int64_t)nslots < 2 * n) nslots <<= 1;
    std::vector<int4> slots(nslots, make_int4(7, 7, 7, 7));  // stale contents: k_ds_clear has to empty the table
    std::vector<int> first_idx(nslots, 3), slot_of((size_t)n);
    std::vector<P3> dst((size_t)n);
    int count = -1;
    const P3 *src = reinterpret_cast<const P3 *>(xyz);
    int4 *sl = slots.data();
    int *fi = first_idx.data(), *so = slot_of.data(), *cnt = &count;
    P3 *ds = dst.data();
    const int nn = (int)n, grid = (int)((n + 255) / 256);
    emu::launch_waves((int)((nslots + 255) / 256), 256, [=]() { k_ds_clear(sl, fi, (int)nslots); });
    emu::launch_waves(grid, 256, [=]() { k_ds_insert(src, nn, d_n, voxel_size, sl, nslots - 1, fi, so); });
    const kicp_scan_args sa = emu::scan_state().next(n, kScanTile);
    emu::launch_waves((int)((n + kScanTile - 1) / kScanTile), kScanThreads, [=]() { k_ds_select(src, nn, d_n, fi, so, ds, cnt, sa); });
    memcpy(out, dst.data(), (size_t)count * sizeof(P3));
    return count;
}

// kiss_icp::Preprocessor::Preprocess + the transform to the base frame: enqueue_preprocess.  omega = log(relative_motion) as
// (upsilon, omega), which the product computes on the host once per frame.
int64_t kf_emu_preprocess(const double *xyz, int64_t n, const double *stamps, int64_t n_stamps, const double omega[6],
                          const double lidar_to_base[7], double max_range, double min_range, int32_t deskew, double *out) {
    if (n == 0) return 0;
    PreArgs a;
    a.deskew = (deskew && n_stamps > 0) ? 1 : 0;
    a.max_range = max_range, a.min_range = min_range;
    a.lidar_to_base = Pose{lidar_to_base[0], lidar_to_base[1], lidar_to_base[2], lidar_to_base[3], lidar_to_base[4], lidar_to_base[5],
                           lidar_to_base[6]};
    for (int k = 0; k < 6; ++k) a.omega[k] = a.deskew ? omega[k] : 0.0;
    std::vector<double> mm(2 + 2 * kMinMaxMaxGrid, 0.0);
    static unsigned int mm_ticket = 0;  // put back to zero by the kernel itself
    const P3 *src = reinterpret_cast<const P3 *>(xyz);
    const int nn = (int)n;
    double *dmm = mm.data();
    if (a.deskew) {
        unsigned int *tk = &mm_ticket;
        const int grid = std::min(kMinMaxMaxGrid, (nn + 2047) / 2048);
        emu::launch_waves(grid, 256, [=]() { k_stamp_minmax(stamps, nn, dmm + 2, tk, dmm); });
    }
    std::vector<P3> dst((size_t)n);
    int count = -1;
    int *cnt = &count;
    P3 *ds = dst.data();
    const kicp_scan_args sa = emu::scan_state().next(n, kScanTile);
    emu::launch_waves((int)((n + kScanTile - 1) / kScanTile), kScanThreads, [=]() { k_preprocess_select(src, stamps, dmm, nn, a, ds, cnt, sa); });
    memcpy(out, dst.data(), (size_t)count * sizeof(P3));
    return count;
}

// PointCloud2-shaped ingest: float32 or float64 fields at a byte stride widened to packed doubles
void kf_emu_ingest(const unsigned char *raw, int64_t n, int32_t is_f32, int32_t step, int32_t ox, int32_t oy, int32_t oz, double *out) {
    if (n == 0) return;
    P3 *o = reinterpret_cast<P3 *>(out);
    const IngestArgs a{is_f32, step, ox, oy, oz};
    const int nn = (int)n;
    emu::launch_waves((int)((n + 255) / 256), 256, [=]() { k_ingest(raw, nn, a, o); });
}
}
