// Device-side front end of KinematicICP::RegisterFrame: the two operations the reference runs
// immediately before the hot path (pipeline/KinematicICP.cpp:38-44, 54-59), both from KISS-ICP v1.2.0:
//   kiss_icp::VoxelDownsample(frame, voxel_size)          first point (input order) of every voxel
//   kiss_icp::Preprocessor::Preprocess(frame, stamps, T)  de-skew p <- exp((s-1) log T) p, then keep min < |p| < max
// followed by the transform to the robot base frame.  Output order is the input order of the survivors (stable stream
// compaction — the library's own single-pass scan, kicp_scan.cuh, fused into the kernels that decide what survives), which is what
// the CPU restatement produces.  Compiled with -fmad=false like kicp_map.cu: voxel floors and
// range tests are evaluated in plain IEEE order.
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstring>

#include "kicp_device.cuh"

using namespace kicp_dev;

#include "kicp_frontend_kernels.cuh"  // struct P3, IngestArgs, PreArgs and the kernels of this file

// ------------------------------------------------------------------------------------------------------------ host
namespace {
struct Scratch {
    P3 *in = nullptr, *out = nullptr, *out2 = nullptr, *out3 = nullptr;
    double *stamps = nullptr, *d_mm = nullptr;  // d_mm: {min, max} of the stamps, then the per-CTA partials of their reduction
    unsigned int *d_mm_ticket = nullptr;
    unsigned char *raw = nullptr;
    size_t raw_cap = 0;
    int *h_count = nullptr;  // pinned, 4 ints
    uint32_t *h_mapctr = nullptr;  // pinned, 8 words: the map's counters after an asynchronous update
    cudaEvent_t ev_front = nullptr;  // front end finished (the copy stream waits for it before staging the clouds)
    int64_t last_source = 0;       // registration-source points of the previous frame (sizes this frame's grid and staging copy)
    P3 *h_frame = nullptr, *h_source = nullptr;  // pinned staging of the two clouds RegisterFrame returns (cap points each)
    int64_t staged_frame = 0, staged_source = 0;
    double timing[8] = {0};  // host-side stage times of the last kicp_register_frame, milliseconds (debug export)
    int *first_idx = nullptr, *slot_of = nullptr, *d_count = nullptr;
    int4 *slots = nullptr;
    int64_t cap = 0;
    uint32_t nslots = 0;
};
void release(Scratch &s) {
    cudaFree(s.in), cudaFree(s.out), cudaFree(s.out2), cudaFree(s.out3), cudaFree(s.stamps), cudaFree(s.d_mm_ticket);
    cudaFree(s.first_idx), cudaFree(s.slot_of), cudaFree(s.slots), cudaFree(s.d_count), cudaFree(s.d_mm);
    cudaFree(s.raw), cudaFreeHost(s.h_count), cudaFreeHost(s.h_frame), cudaFreeHost(s.h_source), cudaFreeHost(s.h_mapctr);
    if (s.ev_front) cudaEventDestroy(s.ev_front);
    s = Scratch();
}
void free_scratch(kicp_ctx *c) {
    if (!c->frontend) return;
    release(*static_cast<Scratch *>(c->frontend));
    delete static_cast<Scratch *>(c->frontend);
    c->frontend = nullptr;
}
// the scratch belongs to the context (created on first use, freed by kicp_ctx_destroy); like every other object of the
// library it is not thread-safe
Scratch &scratch(kicp_ctx *c) {
    if (!c->frontend) {
        c->frontend = new Scratch();
        c->frontend_free = free_scratch;
    }
    return *static_cast<Scratch *>(c->frontend);
}

int reserve(kicp_ctx *c, Scratch &s, int64_t n) {
    if (n <= s.cap) return KICP_OK;
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    KICP_CUDA(cudaStreamSynchronize(c->copy_stream));
    release(s);
    const int64_t cap = std::max<int64_t>(n + n / 4, 4096);
    uint32_t nslots = 1;
    while (nslots < (uint64_t)cap * 2) nslots <<= 1;
    KICP_CUDA(cudaMalloc(&s.in, cap * sizeof(P3)));
    KICP_CUDA(cudaMalloc(&s.out, cap * sizeof(P3)));
    KICP_CUDA(cudaMalloc(&s.out2, cap * sizeof(P3)));
    KICP_CUDA(cudaMalloc(&s.out3, cap * sizeof(P3)));
    KICP_CUDA(cudaMalloc(&s.d_mm, (2 + 2 * kMinMaxMaxGrid) * sizeof(double)));
    KICP_CUDA(cudaMalloc(&s.d_mm_ticket, sizeof(unsigned int)));
    KICP_CUDA(cudaMemsetAsync(s.d_mm_ticket, 0, sizeof(unsigned int), c->stream));
    KICP_CUDA(cudaMallocHost(&s.h_count, 4 * sizeof(int)));
    KICP_CUDA(cudaMallocHost(&s.h_mapctr, 8 * sizeof(uint32_t)));
    KICP_CUDA(cudaEventCreateWithFlags(&s.ev_front, cudaEventDisableTiming));
    KICP_CUDA(cudaMallocHost(&s.h_frame, cap * sizeof(P3)));
    KICP_CUDA(cudaMallocHost(&s.h_source, cap * sizeof(P3)));
    KICP_CUDA(cudaMalloc(&s.stamps, cap * sizeof(double)));
    KICP_CUDA(cudaMalloc(&s.first_idx, (size_t)nslots * sizeof(int)));
    KICP_CUDA(cudaMalloc(&s.slot_of, cap * sizeof(int)));
    KICP_CUDA(cudaMalloc(&s.slots, (size_t)nslots * sizeof(int4)));
    KICP_CUDA(cudaMalloc(&s.d_count, 4 * sizeof(int)));
    s.cap = cap, s.nslots = nslots;
    return KICP_OK;
}

// Sophus SE3::log of a pose7 (host, once per frame)
void se3_log_host(const double p[7], double out[6]) {
    const double eps = 1e-10;
    const double qx = p[0], qy = p[1], qz = p[2], qw = p[3];
    const double squared_n = qx * qx + qy * qy + qz * qz;
    double two_atan_nbyw_by_n, theta;
    if (squared_n < eps * eps) {
        two_atan_nbyw_by_n = 2.0 / qw - (2.0 / 3.0) * squared_n / (qw * qw * qw);
        theta = 2.0 * squared_n / qw;
    } else {
        const double nn = std::sqrt(squared_n);
        const double at = qw < 0.0 ? std::atan2(-nn, -qw) : std::atan2(nn, qw);
        two_atan_nbyw_by_n = 2.0 * at / nn;
        theta = two_atan_nbyw_by_n * nn;
    }
    const double w[3] = {two_atan_nbyw_by_n * qx, two_atan_nbyw_by_n * qy, two_atan_nbyw_by_n * qz};
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
    double cc;
    if (std::abs(theta) < eps) {
        cc = 1.0 / 12.0;
    } else {
        const double half = 0.5 * theta;
        cc = (1.0 - theta * std::cos(half) / (2.0 * std::sin(half))) / (theta * theta);
    }
    double Vinv[9];
    for (int i = 0; i < 9; ++i) Vinv[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + cc * O2[i];
    for (int i = 0; i < 3; ++i) out[i] = Vinv[3 * i] * p[4] + Vinv[3 * i + 1] * p[5] + Vinv[3 * i + 2] * p[6];
    out[3] = w[0], out[4] = w[1], out[5] = w[2];
}

// VoxelDownsample of src[0, n) (n = *d_n when given, else n_max) into dst, survivor count to d_count_out; all on the stream:
// clear the scratch hash, claim a slot per voxel with the smallest input index, select the points that hold their voxel's index
int enqueue_downsample(kicp_ctx *c, Scratch &s, const P3 *src, int n_max, const int *d_n, double voxel_size, P3 *dst, int *d_count_out) {
    const int threads = 256, blocks = (n_max + threads - 1) / threads;
    k_ds_clear<<<(s.nslots + 255) / 256, 256, 0, c->stream>>>(s.slots, s.first_idx, (int)s.nslots);
    KICP_CHECK_LAUNCH(c);
    k_ds_insert<<<blocks, threads, 0, c->stream>>>(src, n_max, d_n, voxel_size, s.slots, s.nslots - 1, s.first_idx, s.slot_of);
    KICP_CHECK_LAUNCH(c);
    kicp_scan_args sa;
    KICP_TRY(kicp_scan_next(c, n_max, &sa));
    k_ds_select<<<(n_max + kScanTile - 1) / kScanTile, kScanThreads, 0, c->stream>>>(src, n_max, d_n, s.first_idx, s.slot_of, dst, d_count_out,
                                                                                    sa);
    KICP_CHECK_LAUNCH(c);
    return KICP_OK;
}

// Preprocess of s.in[0, n) (stamps still on the host) + transform to the base frame, compacted into dst
int enqueue_preprocess(kicp_ctx *c, Scratch &s, int n, const double *stamps, int64_t n_stamps, const double relative_motion[7],
                       const double lidar_to_base[7], double max_range, double min_range, int deskew, P3 *dst, int *d_count_out) {
    PreArgs a;
    a.deskew = (deskew && n_stamps > 0) ? 1 : 0;  // Preprocessing.cpp: `!deskew_ || timestamps.empty()` returns the frame as is
    a.max_range = max_range, a.min_range = min_range;
    a.lidar_to_base = Pose{lidar_to_base[0], lidar_to_base[1], lidar_to_base[2], lidar_to_base[3], lidar_to_base[4], lidar_to_base[5],
                           lidar_to_base[6]};
    for (int k = 0; k < 6; ++k) a.omega[k] = 0.0;
    if (a.deskew) {
        KICP_CUDA(cudaMemcpyAsync(s.stamps, stamps, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        const int grid = std::min(kMinMaxMaxGrid, (n + 2047) / 2048);
        k_stamp_minmax<<<grid, 256, 0, c->stream>>>(s.stamps, n, s.d_mm + 2, s.d_mm_ticket, s.d_mm);
        KICP_CHECK_LAUNCH(c);
        se3_log_host(relative_motion, a.omega);
    }
    kicp_scan_args sa;
    KICP_TRY(kicp_scan_next(c, n, &sa));
    k_preprocess_select<<<(n + kScanTile - 1) / kScanTile, kScanThreads, 0, c->stream>>>(s.in, s.stamps, s.d_mm, n, a, dst, d_count_out, sa);
    KICP_CHECK_LAUNCH(c);
    return KICP_OK;
}
}  // namespace

extern "C" int kicp_voxel_downsample(kicp_ctx *c, const double *xyz, int64_t n, double voxel_size, double *out_xyz, int64_t cap,
                                     int64_t *m) {
    if (!c || !m || n < 0 || (n > 0 && (!xyz || !out_xyz)) || !(voxel_size > 0.0) || n > 0x3FFFFFFF) return KICP_ERR_INVALID;
    *m = 0;
    if (n == 0) return KICP_OK;
    KICP_CUDA(cudaSetDevice(c->device));
    Scratch &s = scratch(c);
    KICP_TRY(reserve(c, s, n));
    KICP_CUDA(cudaMemcpyAsync(s.in, xyz, (size_t)n * sizeof(P3), cudaMemcpyHostToDevice, c->stream));
    KICP_TRY(enqueue_downsample(c, s, s.in, (int)n, nullptr, voxel_size, s.out, s.d_count));
    int count = 0;
    KICP_CUDA(cudaMemcpyAsync(&count, s.d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    *m = count;
    if (count > cap) return KICP_ERR_CAPACITY;
    KICP_CUDA(cudaMemcpy(out_xyz, s.out, (size_t)count * sizeof(P3), cudaMemcpyDeviceToHost));
    return KICP_OK;
}

extern "C" int kicp_preprocess(kicp_ctx *c, const double *xyz, int64_t n, const double *stamps, int64_t n_stamps,
                               const double relative_motion[7], const double lidar_to_base[7], double max_range, double min_range,
                               int deskew, double *out_xyz, int64_t cap, int64_t *m) {
    if (!c || !m || n < 0 || (n > 0 && (!xyz || !out_xyz)) || !relative_motion || !lidar_to_base || n > 0x3FFFFFFF)
        return KICP_ERR_INVALID;
    if (deskew && n_stamps > 0 && (n_stamps != n || !stamps)) return KICP_ERR_INVALID;
    *m = 0;
    if (n == 0) return KICP_OK;
    KICP_CUDA(cudaSetDevice(c->device));
    Scratch &s = scratch(c);
    KICP_TRY(reserve(c, s, n));
    KICP_CUDA(cudaMemcpyAsync(s.in, xyz, (size_t)n * sizeof(P3), cudaMemcpyHostToDevice, c->stream));
    KICP_TRY(enqueue_preprocess(c, s, (int)n, stamps, n_stamps, relative_motion, lidar_to_base, max_range, min_range, deskew, s.out,
                                s.d_count));
    int count = 0;
    KICP_CUDA(cudaMemcpyAsync(&count, s.d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    *m = count;
    if (count > cap) return KICP_ERR_CAPACITY;
    KICP_CUDA(cudaMemcpy(out_xyz, s.out, (size_t)count * sizeof(P3), cudaMemcpyDeviceToHost));
    return KICP_OK;
}

// ------------------------------------------------------------------------------------------- fused RegisterFrame
// KinematicICP::RegisterFrame (pipeline/KinematicICP.cpp:48-85) with every per-point stage on the device and the frame
// resident in HBM from ingest to map update: one upload of the raw scan, one pose-sized download, plus the two point
// clouds the reference returns by value.  The scalar threshold model (CorrespondenceThreshold) stays with the caller:
// it needs tau before and the pose after, both of which cross this boundary anyway.
extern "C" int kicp_register_frame(kicp_map *map, const kicp_frame_input *in, const double deskew_motion[7], const double lidar_to_base[7],
                                   const double last_pose[7], const double relative_odometry[7], double tau, const kicp_frame_params *fp,
                                   double out_pose[7], double *out_frame, int64_t cap_frame, int64_t *n_frame, double *out_source,
                                   int64_t cap_source, int64_t *n_source, kicp_reg_result *result) {
    if (!map || !in || !deskew_motion || !lidar_to_base || !last_pose || !relative_odometry || !fp || !out_pose) return KICP_ERR_INVALID;
    const int64_t n = in->n;
    if (n < 0 || n > 0x3FFFFFFF || (n > 0 && !in->data) || !(fp->voxel_size > 0.0)) return KICP_ERR_INVALID;
    if (in->dtype != KICP_DTYPE_F64 && in->dtype != KICP_DTYPE_F32) return KICP_ERR_INVALID;
    const int elem = in->dtype == KICP_DTYPE_F32 ? 4 : 8;
    const int step = in->point_step > 0 ? in->point_step : 3 * elem;
    const int ox = in->point_step > 0 ? in->offset_x : 0, oy = in->point_step > 0 ? in->offset_y : elem,
              oz = in->point_step > 0 ? in->offset_z : 2 * elem;
    if (ox < 0 || oy < 0 || oz < 0 || ox + elem > step || oy + elem > step || oz + elem > step) return KICP_ERR_INVALID;
    const int deskew = fp->deskew && in->n_stamps > 0;
    if (deskew && (in->n_stamps != n || !in->stamps)) return KICP_ERR_INVALID;
    if (n_frame) *n_frame = 0;
    if (n_source) *n_source = 0;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&t0]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    kicp_ctx *c = map->ctx;
    KICP_CUDA(cudaSetDevice(c->device));
    Scratch &s = scratch(c);
    int counts[3] = {0, 0, 0};
    if (n > 0) {
        KICP_TRY(reserve(c, s, n));
        const bool packed_f64 = in->dtype == KICP_DTYPE_F64 && step == 24 && ox == 0 && oy == 8 && oz == 16;
        if (packed_f64) {
            KICP_CUDA(cudaMemcpyAsync(s.in, in->data, (size_t)n * sizeof(P3), cudaMemcpyHostToDevice, c->stream));
        } else {
            const size_t raw_bytes = (size_t)n * step;
            if (raw_bytes > s.raw_cap) {
                KICP_CUDA(cudaStreamSynchronize(c->stream));
                cudaFree(s.raw);
                s.raw = nullptr, s.raw_cap = 0;
                KICP_CUDA(cudaMalloc(&s.raw, raw_bytes + raw_bytes / 4));
                s.raw_cap = raw_bytes + raw_bytes / 4;
            }
            KICP_CUDA(cudaMemcpyAsync(s.raw, in->data, raw_bytes, cudaMemcpyHostToDevice, c->stream));
            k_ingest<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(s.raw, (int)n, IngestArgs{in->dtype == KICP_DTYPE_F32, step, ox, oy, oz},
                                                                       s.in);
            KICP_CHECK_LAUNCH(c);
        }
        // s.out = preprocessed frame in base, s.out2 = frame_downsample (0.5 vs), s.out3 = source (1.5 vs); each stage reads
        // the previous stage's survivor count from device memory (KinematicICP.cpp:38-44, 54-62)
        KICP_TRY(enqueue_preprocess(c, s, (int)n, in->stamps, in->n_stamps, deskew_motion, lidar_to_base, fp->max_range, fp->min_range,
                                    fp->deskew, s.out, s.d_count));
        KICP_TRY(enqueue_downsample(c, s, s.out, (int)n, s.d_count, fp->voxel_size * 0.5, s.out2, s.d_count + 1));
        KICP_TRY(enqueue_downsample(c, s, s.out2, (int)n, s.d_count + 1, fp->voxel_size * 1.5, s.out3, s.d_count + 2));
        // A frame is ONE host synchronisation (at its end): every later stage reads the survivor counts and the pose from device
        // memory.  The legacy order (read the counts back, size everything exactly) remains for callers whose output buffers
        // might be too small for the worst case — they are owed KICP_ERR_CAPACITY before the map changes — and as option
        // "frame_sync" for A/B measurements.
        const bool async_frame = !c->frame_sync && (!out_frame || cap_frame >= n) && (!out_source || cap_source >= n);
        if (async_frame) {
            s.timing[0] = ms_since();  // upload + front end enqueued
            const bool stage = fp->stage_clouds || out_frame || out_source;
            if (stage) KICP_CUDA(cudaEventRecord(s.ev_front, c->stream));
            // registration of the source against the local map; the grid is sized from the previous frame's source
            c->reg_n_hint = s.last_source > 0 ? s.last_source + s.last_source / 4 + 256 : 0;
            KICP_TRY(kicp_enqueue_registration_device(map, reinterpret_cast<const double *>(s.out3), n, s.d_count + 2, last_pose,
                                                      relative_odometry, tau, &fp->reg));
            s.timing[1] = ms_since();  // + registration enqueued
            // the two clouds go to pinned staging on the copy stream while the registration runs (sizes: upper bounds)
            s.staged_frame = s.staged_source = 0;
            const int64_t src_copy = std::min<int64_t>(n, std::max<int64_t>(2 * s.last_source, 32768));
            if (stage) {
                KICP_CUDA(cudaStreamWaitEvent(c->copy_stream, s.ev_front, 0));
                KICP_CUDA(cudaMemcpyAsync(s.h_frame, s.out, (size_t)n * sizeof(P3), cudaMemcpyDeviceToHost, c->copy_stream));
                KICP_CUDA(cudaMemcpyAsync(s.h_source, s.out3, (size_t)src_copy * sizeof(P3), cudaMemcpyDeviceToHost, c->copy_stream));
            }
            // local_map_.Update(frame_downsample, new_pose) (KinematicICP.cpp:79), pose and count still on the device; a NaN pose
            // (zero correspondences, which the reference does not defend against) leaves the map untouched
            KICP_TRY(kicp_map_update_pose_async(map, reinterpret_cast<const double *>(s.out2), n, s.d_count + 1, kicp_device_result(c),
                                                s.h_mapctr));
            KICP_CUDA(cudaMemcpyAsync(s.h_count, s.d_count, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            KICP_CUDA(cudaStreamSynchronize(c->stream));  // THE synchronisation of the frame
            s.timing[2] = s.timing[3] = ms_since();      // + registration and map update finished
            for (int k = 0; k < 3; ++k) counts[k] = s.h_count[k];
            s.last_source = counts[2];
            const kicp_reg_result &r = *c->h_result;
            for (int k = 0; k < 7; ++k) out_pose[k] = r.pose[k];
            if (result) *result = r;
            int status = r.status;
            const int mst = kicp_map_finish_update(map, s.h_mapctr);
            if (mst != KICP_OK) status = mst;
            if (n_frame) *n_frame = counts[0];
            if (n_source) *n_source = counts[2];
            if (stage) {
                if (counts[2] > src_copy)  // the source outgrew the previous frame's twofold: fetch the rest
                    KICP_CUDA(cudaMemcpyAsync(s.h_source + src_copy, s.out3 + src_copy, (size_t)(counts[2] - src_copy) * sizeof(P3),
                                              cudaMemcpyDeviceToHost, c->copy_stream));
                KICP_CUDA(cudaStreamSynchronize(c->copy_stream));
                s.staged_frame = counts[0], s.staged_source = counts[2];
                if (out_frame && counts[0]) memcpy(out_frame, s.h_frame, (size_t)counts[0] * sizeof(P3));
                if (out_source && counts[2]) memcpy(out_source, s.h_source, (size_t)counts[2] * sizeof(P3));
            }
            s.timing[4] = ms_since();  // + clouds delivered
            return status;
        }
        KICP_CUDA(cudaMemcpyAsync(s.h_count, s.d_count, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        KICP_CUDA(cudaStreamSynchronize(c->stream));  // legacy order: 12 bytes back, sizes the registration grid exactly
        for (int k = 0; k < 3; ++k) counts[k] = s.h_count[k];
        s.last_source = counts[2];
    }
    s.timing[0] = ms_since();  // upload + front end (ingest, de-skew, filter, transform, both down-samples)
    if (n_frame) *n_frame = counts[0];
    if (n_source) *n_source = counts[2];
    if ((out_frame && counts[0] > cap_frame) || (out_source && counts[2] > cap_source)) return KICP_ERR_CAPACITY;
    // registration of the source against the local map, enqueued first so that the downloads below overlap it
    KICP_TRY(kicp_enqueue_registration_device(map, counts[2] ? reinterpret_cast<const double *>(s.out3) : nullptr, counts[2], nullptr,
                                              last_pose, relative_odometry, tau, &fp->reg));
    s.timing[1] = ms_since();  // + registration enqueued
    // the two clouds go to pinned staging on the copy stream while the registration runs
    const bool stage = fp->stage_clouds || out_frame || out_source;
    s.staged_frame = s.staged_source = 0;
    if (stage && counts[0]) KICP_CUDA(cudaMemcpyAsync(s.h_frame, s.out, (size_t)counts[0] * sizeof(P3), cudaMemcpyDeviceToHost, c->copy_stream));
    if (stage && counts[2])
        KICP_CUDA(cudaMemcpyAsync(s.h_source, s.out3, (size_t)counts[2] * sizeof(P3), cudaMemcpyDeviceToHost, c->copy_stream));
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    s.timing[2] = ms_since();  // + registration finished
    const kicp_reg_result &r = *c->h_result;
    for (int k = 0; k < 7; ++k) out_pose[k] = r.pose[k];
    if (result) *result = r;
    int status = r.status;
    // local_map_.Update(frame_downsample, new_pose) (KinematicICP.cpp:79); a NaN pose (zero correspondences, which the
    // reference does not defend against) is reported and leaves the map untouched
    if (status == KICP_OK && counts[1] > 0)
        status = kicp_map_update_pose_device(map, reinterpret_cast<const double *>(s.out2), counts[1], r.pose);
    s.timing[3] = ms_since();  // + map update
    KICP_CUDA(cudaStreamSynchronize(c->copy_stream));
    if (stage) s.staged_frame = counts[0], s.staged_source = counts[2];
    if (out_frame && counts[0]) memcpy(out_frame, s.h_frame, (size_t)counts[0] * sizeof(P3));
    if (out_source && counts[2]) memcpy(out_source, s.h_source, (size_t)counts[2] * sizeof(P3));
    s.timing[4] = ms_since();  // + clouds delivered
    return status;
}

/* The clouds of the last kicp_register_frame on this context (fp->stage_clouds), in context-owned pinned host memory, valid
 * until the next kicp_register_frame / kicp_preprocess / kicp_voxel_downsample call on the same context. */
extern "C" int kicp_frame_clouds(kicp_ctx *c, const double **frame, int64_t *n_frame, const double **source, int64_t *n_source) {
    if (!c || !frame || !n_frame || !source || !n_source) return KICP_ERR_INVALID;
    Scratch &s = scratch(c);
    *frame = reinterpret_cast<const double *>(s.h_frame), *n_frame = s.staged_frame;
    *source = reinterpret_cast<const double *>(s.h_source), *n_source = s.staged_source;
    return KICP_OK;
}

// debugging aid (not part of the public header): cumulative host-side stage times of the last kicp_register_frame, ms
extern "C" int kicp_debug_frame_timing(kicp_ctx *c, double out[8]) {
    if (!c || !out) return KICP_ERR_INVALID;
    for (int k = 0; k < 8; ++k) out[k] = scratch(c).timing[k];
    return KICP_OK;
}
