// Kernels of the device-side front end (included by kicp_frontend.cu, which is compiled with -fmad=false): VoxelDownsample's
// min-index-per-voxel insert and its order-preserving select, the PointCloud2 ingest, the stamps' min / max, Preprocess (de-skew,
// range filter, base transform) compacting its survivors as it goes.  The selects are the library's own single-pass scans
// (kicp_scan.cuh) fused into the kernels that decide what survives: no flag arrays, no library kernels.  A header of its own so that
// tests/emu can compile the kernels for the host against the SIMT emulator without the CUDA-runtime orchestration of kicp_frontend.cu.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstring>

#include "kicp_device.cuh"
#include "kicp_scan.cuh"

using namespace kicp_dev;

struct P3 {
    double x, y, z;
};

// ------------------------------------------------------------------------------------------------ VoxelDownsample
// `d_n` (optional) is the device-resident point count of a stage whose input was compacted on the device: the grid is sized
// for the host-side upper bound `n_max` and the tail threads retire, so chained stages need no host round trip.
__global__ void k_ds_insert(const P3 *__restrict__ pts, int n_max, const int *__restrict__ d_n, double vs, int4 *slots, uint32_t mask_in,
                            int *first_idx, int *slot_of) {
    __shared__ uint32_t s_mask[32];
    const uint32_t mask = lane_private(mask_in, s_mask);  // divergence safety, see kicp_device.cuh
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = d_n ? min(*d_n, n_max) : n_max;
    if (i >= n) return;
    const P3 p = pts[i];
    const int kx = voxel_coord(p.x, vs), ky = voxel_coord(p.y, vs), kz = voxel_coord(p.z, vs);
    uint32_t h = voxel_hash(kx, ky, kz) & mask;
    volatile int4 *vsl = slots;
    while (true) {
        uint32_t st = (uint32_t)vsl[h].w;
        if (st == KICP_SLOT_EMPTY) {
            const uint32_t old = atomicCAS((unsigned int *)&slots[h].w, KICP_SLOT_EMPTY, KICP_SLOT_LOCKED);
            if (old == KICP_SLOT_EMPTY) {
                vsl[h].x = kx, vsl[h].y = ky, vsl[h].z = kz;
                __threadfence();
                atomicExch((unsigned int *)&slots[h].w, 0u);
                break;
            }
            st = old;
        }
        if (st == KICP_SLOT_LOCKED) continue;
        __threadfence();
        if (vsl[h].x == kx && vsl[h].y == ky && vsl[h].z == kz) break;
        h = (h + 1) & mask;
    }
    atomicMin(&first_idx[h], i);  // the voxel keeps the point with the smallest input index
    slot_of[i] = (int)h;
}

// The survivors of the down-sample — the points whose index is the smallest of their voxel — copied to `dst` in input order; their
// number goes to *d_count_out.  Grid = ceil(n_max / kScanTile) CTAs of kScanThreads threads (kicp_scan.cuh).
__global__ void __launch_bounds__(kScanThreads) k_ds_select(const P3 *__restrict__ src, int n_max, const int *__restrict__ d_n,
                                                            const int *__restrict__ first_idx, const int *__restrict__ slot_of, P3 *dst,
                                                            int *d_count_out, kicp_scan_args sa) {
    __shared__ uint32_t s_warp[kScanWarps + 1];
    const uint32_t tile = scan_take_tile(sa, &s_warp[kScanWarps]);
    const int n = d_n ? min(*d_n, n_max) : n_max;
    const int64_t i0 = scan_first_item(tile);
    bool keep[kScanItems];
    uint32_t rank[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = i0 + 32 * k;
        keep[k] = i < n && first_idx[slot_of[i]] == (int)i;
    }
    const uint32_t warp_total = scan_warp_ranks(keep, rank);
    uint32_t inclusive;
    const uint32_t base = scan_offset(sa, tile, warp_total, s_warp, inclusive);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (keep[k]) dst[base + rank[k]] = src[i0 + 32 * k];
    if (tile == gridDim.x - 1 && threadIdx.x == 0) *d_count_out = (int)inclusive;
}

// empties the scratch hash of the down-sample: every slot free, no first index yet
__global__ void k_ds_clear(int4 *slots, int *first_idx, int nslots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    slots[i] = make_int4(-1, -1, -1, -1);  // w = KICP_SLOT_EMPTY
    first_idx[i] = 0x7FFFFFFF;
}

// ---------------------------------------------------------------------------------------------------------- ingest
// PointCloud2-shaped input (RosUtils.cpp:30-39 reads float32 x,y,z through an iterator with the message's point_step and
// widens to double): the raw bytes are uploaded once and widened here.
struct IngestArgs {
    int is_f32, step, ox, oy, oz;
};
__global__ void k_ingest(const unsigned char *__restrict__ raw, int n, IngestArgs a, P3 *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char *p = raw + (size_t)i * a.step;
    P3 o;
    if (a.is_f32) {
        float x, y, z;  // fields need not be 4-byte aligned inside the message: assemble from bytes
        memcpy(&x, p + a.ox, 4), memcpy(&y, p + a.oy, 4), memcpy(&z, p + a.oz, 4);
        o = P3{(double)x, (double)y, (double)z};
    } else {
        memcpy(&o.x, p + a.ox, 8), memcpy(&o.y, p + a.oy, 8), memcpy(&o.z, p + a.oz, 8);
    }
    out[i] = o;
}

// ------------------------------------------------------------------------------------------------------ Preprocess
struct PreArgs {
    double omega[6];       // log(relative_motion) as (upsilon, omega), computed on the host once per frame
    Pose lidar_to_base;
    double max_range, min_range;  // stamps are normalised as (t - t_min) / (t_max - t_min)
    int deskew;
};

// Sophus SE3::exp applied to a point (same operation order as the CPU restatement)
__device__ void se3_exp_apply(const double a[6], double px, double py, double pz, double &ox, double &oy, double &oz) {
    const double eps = 1e-10;
    const double wx = a[3], wy = a[4], wz = a[5];
    const double theta_sq = wx * wx + wy * wy + wz * wz;
    double theta, imag, real;
    if (theta_sq < eps * eps) {
        theta = 0.0;
        const double theta_po4 = theta_sq * theta_sq;
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
        real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
    } else {
        theta = sqrt(theta_sq);
        const double half_theta = 0.5 * theta;
        imag = sin(half_theta) / theta;
        real = cos(half_theta);
    }
    const double qx = imag * wx, qy = imag * wy, qz = imag * wz, qw = real;
    // V = I + c1 W + c2 W^2 (or the rotation matrix when theta < eps), t = V * upsilon
    double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0}, O2[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
    if (theta < eps) {
        const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
        const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx;
        const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
        V[0] = 1 - (tyy + tzz), V[1] = txy - twz, V[2] = txz + twy;
        V[3] = txy + twz, V[4] = 1 - (txx + tzz), V[5] = tyz - twx;
        V[6] = txz - twy, V[7] = tyz + twx, V[8] = 1 - (txx + tyy);
    } else {
        const double theta2 = theta * theta;
        const double c1 = (1.0 - cos(theta)) / theta2, c2 = (theta - sin(theta)) / (theta2 * theta);
        for (int i = 0; i < 9; ++i) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * O[i] + c2 * O2[i];
    }
    const double tx = V[0] * a[0] + V[1] * a[1] + V[2] * a[2];
    const double ty = V[3] * a[0] + V[4] * a[1] + V[5] * a[2];
    const double tz = V[6] * a[0] + V[7] * a[1] + V[8] * a[2];
    double rx, ry, rz;
    quat_rotate(qx, qy, qz, qw, px, py, pz, rx, ry, rz);
    ox = rx + tx, oy = ry + ty, oz = rz + tz;
}

// {min, max} of the stamps (Preprocessing.cpp normalises them with std::minmax_element): every CTA reduces a grid-stride share into
// `partial[2 * blockIdx.x]`, the CTA that finishes last (ticket) reduces the partials into d_mm and puts the ticket back to zero.
// Grid <= kMinMaxMaxGrid CTAs of 256 threads.
constexpr int kMinMaxMaxGrid = 128;
__global__ void __launch_bounds__(256) k_stamp_minmax(const double *__restrict__ v, int n, double *partial, unsigned int *ticket, double *d_mm) {
    __shared__ double s_lo[8], s_hi[8];
    __shared__ uint32_t s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double lo = DBL_MAX, hi = -DBL_MAX;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = v[i];
        lo = x < lo ? x : lo, hi = x > hi ? x : hi;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const double l2 = __shfl_xor_sync(0xFFFFFFFFu, lo, d), h2 = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
        lo = l2 < lo ? l2 : lo, hi = h2 > hi ? h2 : hi;
    }
    if (lane == 0) s_lo[warp] = lo, s_hi[warp] = hi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) lo = s_lo[w] < lo ? s_lo[w] : lo, hi = s_hi[w] > hi ? s_hi[w] : hi;
        __stcg(&partial[2 * blockIdx.x], lo), __stcg(&partial[2 * blockIdx.x + 1], hi);
        __threadfence();
        const uint32_t t = atomicAdd(ticket, 1u);
        s_last = t == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!*(volatile uint32_t *)&s_last) return;  // the whole CTA leaves or stays
    __threadfence();
    lo = DBL_MAX, hi = -DBL_MAX;
    if (threadIdx.x < gridDim.x) lo = __ldcg(&partial[2 * threadIdx.x]), hi = __ldcg(&partial[2 * threadIdx.x + 1]);
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const double l2 = __shfl_xor_sync(0xFFFFFFFFu, lo, d), h2 = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
        lo = l2 < lo ? l2 : lo, hi = h2 > hi ? h2 : hi;
    }
    __syncthreads();
    if (lane == 0) s_lo[warp] = lo, s_hi[warp] = hi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) lo = s_lo[w] < lo ? s_lo[w] : lo, hi = s_hi[w] > hi ? s_hi[w] : hi;
        d_mm[0] = lo, d_mm[1] = hi;
        atomicExch(ticket, 0u);
    }
}

// One point through Preprocess: de-skew, range test, transform to the base frame.  Returns whether the point is kept.
__device__ __forceinline__ bool preprocess_point(P3 p, const double *__restrict__ stamps, const double *__restrict__ d_mm, int64_t i,
                                                 const PreArgs &a, P3 &out) {
    if (a.deskew) {
        const double t_min = d_mm[0], t_span = d_mm[1] - d_mm[0];
        const double stamp = (stamps[i] - t_min) / t_span;
        double w[6];
        for (int k = 0; k < 6; ++k) w[k] = (stamp - 1.0) * a.omega[k];
        double ox, oy, oz;
        se3_exp_apply(w, p.x, p.y, p.z, ox, oy, oz);
        p.x = ox, p.y = oy, p.z = oz;
    }
    const double r = sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
    double bx, by, bz;
    pose_apply(a.lidar_to_base, p.x, p.y, p.z, bx, by, bz);  // preprocessed_frame_in_base (KinematicICP.cpp:59)
    out = P3{bx, by, bz};
    return r < a.max_range && r > a.min_range;
}

// Preprocess of pts[0, n): the kept points, in the base frame, to `out` in input order and their number to *d_count_out.
// `d_mm` = {min, max} of the stamps, reduced on the device just before (no host round trip).  Grid as k_ds_select.
__global__ void __launch_bounds__(kScanThreads) k_preprocess_select(const P3 *__restrict__ pts, const double *__restrict__ stamps,
                                                                    const double *__restrict__ d_mm, int n, PreArgs a, P3 *out,
                                                                    int *d_count_out, kicp_scan_args sa) {
    __shared__ uint32_t s_warp[kScanWarps + 1];
    const uint32_t tile = scan_take_tile(sa, &s_warp[kScanWarps]);
    const int64_t i0 = scan_first_item(tile);
    bool keep[kScanItems];
    uint32_t rank[kScanItems];
    P3 q[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = i0 + 32 * k;
        keep[k] = false;
        if (i < n) keep[k] = preprocess_point(pts[i], stamps, d_mm, i, a, q[k]);
    }
    const uint32_t warp_total = scan_warp_ranks(keep, rank);
    uint32_t inclusive;
    const uint32_t base = scan_offset(sa, tile, warp_total, s_warp, inclusive);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (keep[k]) out[base + rank[k]] = q[k];
    if (tile == gridDim.x - 1 && threadIdx.x == 0) *d_count_out = (int)inclusive;
}
