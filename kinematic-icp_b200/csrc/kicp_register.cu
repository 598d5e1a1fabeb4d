// kinematic_icp::KinematicRegistration::ComputeRobotMotion on the device
// (reference: cpp/kinematic_icp/registration/Registration.cpp:48-190 + kiss_icp::VoxelHashMap::GetClosestNeighbor).
//
// One registration = k_reg_init, a Morton binning of the scan (keys -> radix sort -> gather, once, at the initial
// guess), then up to max_num_iterations launches of k_assoc.  k_assoc fuses, for every scan point,
//   q = T p                                          Registration.cpp:74
//   27-voxel probe + nearest neighbour               GetClosestNeighbor (KISS-ICP v1.2.0)
//   gate d < tau                                     Registration.cpp:75
//   r = T p - n,  J = [R e_x | R (-p_y, p_x, 0)]     Registration.cpp:86-93
//   sum of J^T J, J^T r, N, |r|^2                    Registration.cpp:95-118, 48-60
// and its last CTA solves the 2x2 system, applies the unicycle motion model and decides convergence
// (Registration.cpp:119-125, 159-167, 181-184) — so the iteration loop never returns to the host.
// The correspondence list of the reference is never materialised: association and linearisation use the same T.
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>

#include "kicp_device.cuh"

using namespace kicp_dev;

#define KICP_WARPS 8     // warps per CTA in k_assoc
#ifndef KICP_MINB
#define KICP_MINB 2   // resident CTAs per SM the pruned kernels are compiled for (register cap = 65536 / (256 * MINB));
                      // measured: 2 (128 registers, no spills, 4-wide scans) beats 3 and 4
#endif
#define KICP_CH 192      // candidates staged in shared memory per pass (27 voxels x 20 points = 540 worst case)

struct RegState {
    double q[4];  // current estimate: unit quaternion (x, y, z, w) ...
    double t[3];  // ... translation ...
    double R[9];  // ... and the rotation matrix of q, row-major
    double tau, conv, fixed_reg, beta;
    int adaptive, max_iter;
    int iter, done, status;
    unsigned int ticket, window_counter;
    unsigned int generation;  // grid-barrier generation of the persistent kernel (= iterations completed)
    int fused_tail;
    int *iters_out;  // optional: where to publish the iteration count when the registration finishes (profiling)
    double acc[8];  // JTJ00 JTJ01 JTJ11 JTr0 JTr1 N sum|r|^2 (unused)
    double dbg[KICP_MAX_ITERATIONS][4];  // per iteration, ns: windows phase of CTA 0, barrier wait of the last CTA, partial sum, solve
    kicp_reg_result result;
};

// ------------------------------------------------------------------------------------------ SE3 helpers (Sophus)
__device__ void quat_to_matrix(const double q[4], double R[9]) {  // Eigen::Quaternion::toRotationMatrix
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz), R[1] = txy - twz, R[2] = txz + twy;
    R[3] = txy + twz, R[4] = 1 - (txx + tzz), R[5] = tyz - twx;
    R[6] = txz - twy, R[7] = tyz + twx, R[8] = 1 - (txx + tyy);
}

// Sophus SE3 product: q = normalize(a.q * b.q) (the SO3(quaternion) ctor normalises), t = a.t + a.q * b.t
__device__ void se3_compose(const double aq[4], const double at[3], const double bq[4], const double bt[3], double oq[4],
                            double ot[3]) {
    const double ax = aq[0], ay = aq[1], az = aq[2], aw = aq[3];
    const double bx = bq[0], by = bq[1], bz = bq[2], bw = bq[3];
    double w = aw * bw - ax * bx - ay * by - az * bz;
    double x = aw * bx + ax * bw + ay * bz - az * by;
    double y = aw * by + ay * bw + az * bx - ax * bz;
    double z = aw * bz + az * bw + ax * by - ay * bx;
    const double len = sqrt(x * x + y * y + z * z + w * w);
    x /= len, y /= len, z /= len, w /= len;
    double rx, ry, rz;
    quat_rotate(ax, ay, az, aw, bt[0], bt[1], bt[2], rx, ry, rz);
    oq[0] = x, oq[1] = y, oq[2] = z, oq[3] = w;
    ot[0] = at[0] + rx, ot[1] = at[1] + ry, ot[2] = at[2] + rz;
}

// Sophus SE3::exp for the tangent the motion model produces: (ux, uy, 0, 0, 0, theta)
__device__ void se3_exp_planar(double ux, double uy, double theta_in, double oq[4], double ot[3]) {
    const double eps = 1e-10;  // Sophus::Constants<double>::epsilon()
    const double wx = 0.0, wy = 0.0, wz = theta_in;
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
        double sh, ch;
        sincos(half_theta, &sh, &ch);
        imag = sh / theta;
        real = ch;
    }
    oq[0] = imag * wx, oq[1] = imag * wy, oq[2] = imag * wz, oq[3] = real;
    // V = I + (1-cos)/th^2 W + (th - sin)/th^3 W^2, or V = R when theta < eps;  W = hat(0, 0, wz)
    double V[9];
    if (theta < eps) {
        quat_to_matrix(oq, V);
    } else {
        double st_, ct_;
        sincos(theta, &st_, &ct_);
        const double c1 = (1.0 - ct_) / (theta * theta);
        const double c2 = (theta - st_) / (theta * theta * theta);
        const double w2 = wz * wz;
        V[0] = 1.0 + c2 * (-w2), V[1] = c1 * (-wz), V[2] = 0.0;
        V[3] = c1 * wz, V[4] = 1.0 + c2 * (-w2), V[5] = 0.0;
        V[6] = 0.0, V[7] = 0.0, V[8] = 1.0;
    }
    ot[0] = V[0] * ux + V[1] * uy;
    ot[1] = V[3] * ux + V[4] * uy;
    ot[2] = V[6] * ux + V[7] * uy;
}

// ---------------------------------------------------------------------------------------------------- kernels
struct RegArgs {
    Pose last, odom;
    double tau, conv, fixed_reg;
    int adaptive, max_iter, fused_tail;
    int *iters_out;
};

// current_estimate = last_robot_pose * relative_wheel_odometry   (Registration.cpp:156)
__global__ void k_reg_init(RegState *st, RegArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double lq[4] = {a.last.qx, a.last.qy, a.last.qz, a.last.qw}, lt[3] = {a.last.tx, a.last.ty, a.last.tz};
    const double oq[4] = {a.odom.qx, a.odom.qy, a.odom.qz, a.odom.qw}, ot[3] = {a.odom.tx, a.odom.ty, a.odom.tz};
    se3_compose(lq, lt, oq, ot, st->q, st->t);
    quat_to_matrix(st->q, st->R);
    st->tau = a.tau, st->conv = a.conv, st->fixed_reg = a.fixed_reg, st->beta = 0.0;
    st->adaptive = a.adaptive, st->max_iter = a.max_iter, st->fused_tail = a.fused_tail;
    st->iter = 0, st->done = a.max_iter <= 0 ? 1 : 0, st->status = KICP_OK;
    st->ticket = 0, st->window_counter = 0, st->generation = 0;
    st->iters_out = a.iters_out;
    if (a.iters_out) *a.iters_out = 0;
    for (int k = 0; k < 8; ++k) st->acc[k] = 0.0;
    kicp_reg_result *r = &st->result;
    for (int k = 0; k < 4; ++k) r->pose[k] = st->q[k];
    for (int k = 0; k < 3; ++k) r->pose[4 + k] = st->t[k];
    r->beta = 0.0, r->last_dx_norm = 0.0, r->iterations = 0, r->status = KICP_OK;
}

__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
    v &= 1023u;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// Morton key of the voxel each point falls in at the initial guess.  The key only orders the scan so that
// consecutive points share voxel neighbourhoods; k_assoc re-derives every voxel from the current estimate, so a
// stale or aliased key costs locality, never correctness.
__global__ void k_morton_keys(const RegState *st, const double *__restrict__ xyz, int n, double voxel_size,
                              uint32_t *__restrict__ keys, int32_t *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    const double *R = st->R, *t = st->t;
    const double qx = R[0] * px + R[1] * py + R[2] * pz + t[0];
    const double qy = R[3] * px + R[4] * py + R[5] * pz + t[1];
    const double qz = R[6] * px + R[7] * py + R[8] * pz + t[2];
    const int ox = voxel_coord(t[0], voxel_size) - 512, oy = voxel_coord(t[1], voxel_size) - 512,
              oz = voxel_coord(t[2], voxel_size) - 512;
    const uint32_t ux = (uint32_t)(voxel_coord(qx, voxel_size) - ox), uy = (uint32_t)(voxel_coord(qy, voxel_size) - oy),
                   uz = (uint32_t)(voxel_coord(qz, voxel_size) - oz);
    keys[i] = spread10(ux) | (spread10(uy) << 1) | (spread10(uz) << 2);
    idx[i] = i;
}

__global__ void k_gather(const double *__restrict__ xyz, const int32_t *__restrict__ idx, int n, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = idx[i];
    out[3 * i] = xyz[3 * s], out[3 * i + 1] = xyz[3 * s + 1], out[3 * i + 2] = xyz[3 * s + 2];
}

// ComputePerturbation's tail + motion model + pose update + convergence test (one thread).
// `acc` holds the (all-reduced) sums of this iteration.
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ void solve_and_update(RegState *st) {
    double s[8];
    for (int k = 0; k < 8; ++k) s[k] = __ldcg(&st->acc[k]);  // written by other CTAs' atomics: read through L2
    const int j = st->iter;
    kicp_reg_result *res = &st->result;
    if (j < KICP_MAX_ITERATIONS)
        for (int k = 0; k < 8; ++k) res->sums[j][k] = s[k];
    const double N = s[5];
    if (j == 0) {
        // ComputeOdometryRegularization (Registration.cpp:48-60): beta = 1 / (mean |T0 p - n|^2 + DBL_MIN), computed
        // once from the first association; the fixed value otherwise (:171-177)
        st->beta = st->adaptive ? 1.0 / (s[6] / N + DBL_MIN) : st->fixed_reg;
        res->beta = st->beta;
    }
    // JTJ /= N; JTr /= N; JTJ += diag(beta, 0); dx = -(JTJ^-1 JTr)     (Registration.cpp:119-125)
    const double a = s[0] / N + st->beta, b = s[1] / N, d = s[2] / N + 0.0;
    const double r0 = s[3] / N, r1 = s[4] / N;
    const double invdet = 1.0 / (a * d - b * b);
    const double i00 = d * invdet, i01 = -b * invdet, i10 = -b * invdet, i11 = a * invdet;
    const double dx0 = -(i00 * r0 + i01 * r1), dx1 = -(i10 * r0 + i11 * r1);
    // motion_model (Registration.cpp:159-167), epsilon = DBL_MIN
    double sn, cs;
    sincos(dx1, &sn, &cs);
    const double ux = dx0 * sn / (dx1 + DBL_MIN);
    const double uy = dx0 * (1.0 - cs) / (dx1 + DBL_MIN);
    double dq[4], dt[3], nq[4], nt[3];
    se3_exp_planar(ux, uy, dx1, dq, dt);
    se3_compose(st->q, st->t, dq, dt, nq, nt);  // current_estimate = current_estimate * delta_motion  (:182)
    for (int k = 0; k < 4; ++k) st->q[k] = nq[k];
    for (int k = 0; k < 3; ++k) st->t[k] = nt[k];
    quat_to_matrix(st->q, st->R);
    const double dxn = sqrt(dx0 * dx0 + dx1 * dx1);
    if (j < KICP_MAX_ITERATIONS) res->dx[j][0] = dx0, res->dx[j][1] = dx1;
    res->last_dx_norm = dxn;
    res->iterations = j + 1;
    for (int k = 0; k < 4; ++k) res->pose[k] = st->q[k];
    for (int k = 0; k < 3; ++k) res->pose[4 + k] = st->t[k];
    st->iter = j + 1;
    int done = (dxn < st->conv) || (j + 1 >= st->max_iter);  // break BEFORE re-association (:184)
    if (!(N > 0.0)) {  // the reference has no guard: the pose is NaN from here on; stop early and say so
        st->status = KICP_WARN_NO_CORRESPONDENCES;
        done = 1;
    }
    if (st->status == KICP_ERR_NCCL) done = 1;  // a peer of the fused exchange never arrived
    res->status = st->status;
    st->done = done;
    if (st->iters_out) *st->iters_out = j + 1;
    for (int k = 0; k < 8; ++k) st->acc[k] = 0.0;
}

__global__ void k_solve(RegState *st) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && !st->done) solve_and_update(st);
}

// The fused association + linearisation + reduction kernel.  One warp owns a window of 32 consecutive (Morton-
// sorted) scan points.  Lanes that fall in the same voxel form a group; for each group the warp probes the 27
// neighbour voxels (lanes 0..26, one hash probe each), stages their points in shared memory, and scans them with
// 32 / 2^ceil(log2 P) lanes per query point so small groups still use every lane.
__global__ void __launch_bounds__(KICP_WARPS * 32) k_assoc(RegState *st, const double *__restrict__ scan, int n, MapView map) {
    if (st->done) return;
    extern __shared__ double smem_d[];
    __shared__ double s_T[12];
    __shared__ double s_part[KICP_WARPS][8];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xFFFFFFFFu;
    // per-warp shared memory: candidate coordinates (SoA), their global point indices, query broadcast, results
    double *sx = smem_d + (size_t)wid * (3 * KICP_CH + 3 * 32 + 32);
    double *sy = sx + KICP_CH, *sz = sy + KICP_CH;
    double *sq = sz + KICP_CH;      // [32][3] query points of the current group
    double *sres_d = sq + 96;       // [32]    winning squared distance per query slot
    int *sg = (int *)(smem_d + (size_t)KICP_WARPS * (3 * KICP_CH + 3 * 32 + 32)) + wid * (KICP_CH + 32);
    int *sres_g = sg + KICP_CH;     // [32]    winning global point index per query slot

    if (threadIdx.x < 9) s_T[threadIdx.x] = st->R[threadIdx.x];
    if (threadIdx.x >= 9 && threadIdx.x < 12) s_T[threadIdx.x] = st->t[threadIdx.x - 9];
    __syncthreads();
    __shared__ MapView s_map[32];
    const MapRegs mr = map_regs(map, s_map);  // per-thread copy of the map view (divergence safety, kicp_device.cuh)
    const double tau = st->tau, vs = map.voxel_size;
    const int num_windows = (n + 31) >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0, cntN = 0, ssq = 0;

    while (true) {
        int w = 0;
        if (lane == 0) w = (int)atomicAdd(&st->window_counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= num_windows) break;
        const int i = w * 32 + lane;
        const bool valid = i < n;
        double px = 0, py = 0, pz = 0;
        if (valid) px = scan[3 * (size_t)i], py = scan[3 * (size_t)i + 1], pz = scan[3 * (size_t)i + 2];
        const double qx = s_T[0] * px + s_T[1] * py + s_T[2] * pz + s_T[9];
        const double qy = s_T[3] * px + s_T[4] * py + s_T[5] * pz + s_T[10];
        const double qz = s_T[6] * px + s_T[7] * py + s_T[8] * pz + s_T[11];
        const int vx = voxel_coord(qx, vs), vy = voxel_coord(qy, vs), vz = voxel_coord(qz, vs);
        const unsigned vmask = __ballot_sync(FULL, valid);
        unsigned gmask = 0;
        if (valid) gmask = __match_any_sync(vmask, vx) & __match_any_sync(vmask, vy) & __match_any_sync(vmask, vz);
        double best = DBL_MAX;
        int bestg = -1;
        unsigned remaining = vmask;
        while (remaining) {
            const int leader = __ffs(remaining) - 1;
            const unsigned gm = __shfl_sync(FULL, gmask, leader);
            const int cvx = __shfl_sync(FULL, vx, leader), cvy = __shfl_sync(FULL, vy, leader), cvz = __shfl_sync(FULL, vz, leader);
            const int P = __popc(gm);
            const bool member = (gm >> lane) & 1u;
            const int myslot = __popc(gm & lt_mask);
            if (member) sq[3 * myslot] = qx, sq[3 * myslot + 1] = qy, sq[3 * myslot + 2] = qz;
            // 27-neighbour probe, KISS shift order (lane k <-> voxel_shifts[k])
            int cnt = 0;
            unsigned blk = 0;
            if (lane < 27) {
                const uint32_t meta = map_probe(mr, cvx + shift_x(lane), cvy + shift_y(lane), cvz + shift_z(lane));
                if (meta != KICP_SLOT_EMPTY) cnt = (int)(meta & 0xFFu), blk = meta >> 8;
            }
            __syncwarp();
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += y;
            }
            const int off = incl - cnt;
            const int C = __shfl_sync(FULL, incl, 31);
            // lanes per query point: 32 / 2^ceil(log2 P)
            const int lg = P <= 1 ? 0 : 32 - __clz(P - 1);
            const int L = 32 >> lg;
            const int slot = lane >> (5 - lg), sub = lane & (L - 1);
            const bool worker = slot < P;
            __syncwarp();
            double wqx = 0, wqy = 0, wqz = 0;
            if (worker) wqx = sq[3 * slot], wqy = sq[3 * slot + 1], wqz = sq[3 * slot + 2];
            double wbest = DBL_MAX;
            int wc = 0x7FFFFFFF, wg = -1;
            for (int base = 0; base < C; base += KICP_CH) {
                for (int j = 0; j < cnt; ++j) {
                    const int c = off + j - base;
                    if (c >= 0 && c < KICP_CH) sg[c] = (int)(blk * (unsigned)mr.cap) + j;
                }
                __syncwarp();
                const int m = min(KICP_CH, C - base);
                for (int c = lane; c < m; c += 32) {
                    const double *gp = mr.pts + (size_t)sg[c] * KICP_PSTRIDE;
                    sx[c] = __ldg(gp), sy[c] = __ldg(gp + 1), sz[c] = __ldg(gp + 2);
                }
                __syncwarp();
                if (worker) {
                    for (int c = sub; c < m; c += L) {
                        const double dx = sx[c] - wqx, dy = sy[c] - wqy, dz = sz[c] - wqz;
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        if (d2 < wbest) wbest = d2, wc = base + c, wg = sg[c];  // strict <: first minimum wins
                    }
                }
                __syncwarp();
            }
            for (int o = L >> 1; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(FULL, wbest, o);
                const int oc = __shfl_xor_sync(FULL, wc, o), og = __shfl_xor_sync(FULL, wg, o);
                if (od < wbest || (od == wbest && oc < wc)) wbest = od, wc = oc, wg = og;
            }
            if (worker && sub == 0) sres_d[slot] = wbest, sres_g[slot] = wg;
            __syncwarp();
            if (member) best = sres_d[myslot], bestg = sres_g[myslot];
            __syncwarp();
            remaining &= ~gm;
        }
        if (valid && bestg >= 0) {
            const double *gp = mr.pts + (size_t)bestg * KICP_PSTRIDE;
            const double rx = qx - __ldg(gp), ry = qy - __ldg(gp + 1), rz = qz - __ldg(gp + 2);  // r = T p - n
            const double rr = rx * rx + ry * ry + rz * rz;
            if (sqrt(rr) < tau) {  // distance < max_correspondance_distance   (Registration.cpp:75)
                // J = [R e_x | R (-p_y, p_x, 0)]      (Registration.cpp:89-91)
                const double c0x = s_T[0], c0y = s_T[3], c0z = s_T[6];
                const double c1x = s_T[1] * px - s_T[0] * py, c1y = s_T[4] * px - s_T[3] * py, c1z = s_T[7] * px - s_T[6] * py;
                a00 += c0x * c0x + c0y * c0y + c0z * c0z;
                a01 += c0x * c1x + c0y * c1y + c0z * c1z;
                a11 += c1x * c1x + c1y * c1y + c1z * c1z;
                b0 += c0x * rx + c0y * ry + c0z * rz;
                b1 += c1x * rx + c1y * ry + c1z * rz;
                cntN += 1.0;
                ssq += rr;
            }
        }
        (void)best;
    }

    // warp shuffle reduction -> one partial per warp -> one set of atomics per CTA
    double v[7] = {a00, a01, a11, b0, b1, cntN, ssq};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) s_part[wid][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double s = 0.0;
        for (int k = 0; k < KICP_WARPS; ++k) s += s_part[k][threadIdx.x];
        atomicAdd(&st->acc[threadIdx.x], s);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(&st->ticket, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        st->ticket = 0;
        st->window_counter = 0;
        if (st->fused_tail) solve_and_update(st);
    }
}



// One candidate step: squared distance of the 32-byte record at `p` to the query; strict `<` keeps the first minimum
// (candidates are visited in the reference's order).  Only (d2, pointer) are tracked; the winner is re-read once.
__device__ __forceinline__ void candidate_step(const double2 xy, const double2 zw, const double *p, double qx, double qy, double qz,
                                               double &best, const double *&bestp) {
    const double dx = xy.x - qx, dy = xy.y - qy, dz = zw.x - qz;
    const double d2 = dx * dx + dy * dy + dz * dz;
    if (d2 < best) best = d2, bestp = p;
}

// Scan the `cnt` points of one voxel, KICP_SCANW per step: 2*KICP_SCANW independent 16-byte loads are in flight per lane
// before the first distance is formed (the path is bound by dependent memory round trips).  Indices past
// the end are clamped to the last point, which is harmless: re-evaluating a point cannot change a strict minimum.
// The loop is per-lane (no warp vote): lanes with fewer points leave earlier and reconverge behind it.
#ifndef KICP_SCANW
#define KICP_SCANW 4
#endif
__device__ __forceinline__ void scan_voxel(const double *vp, int cnt, double qx, double qy, double qz, double &best,
                                           const double *&bestp) {
    for (int j = 0; j < cnt; j += KICP_SCANW) {
        const double *p[KICP_SCANW];
        double2 a[KICP_SCANW], b[KICP_SCANW];
#pragma unroll
        for (int u = 0; u < KICP_SCANW; ++u) {
            p[u] = vp + (size_t)min(j + u, cnt - 1) * KICP_PSTRIDE;
            a[u] = __ldg(reinterpret_cast<const double2 *>(p[u]));
            b[u] = __ldg(reinterpret_cast<const double2 *>(p[u]) + 1);
        }
#pragma unroll
        for (int u = 0; u < KICP_SCANW; ++u) candidate_step(a[u], b[u], p[u], qx, qy, qz, best, bestp);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_assoc_pruned: one thread per scan point, exact pruning of the 27-voxel neighbourhood.
//
// GetClosestNeighbor returns the nearest of all points stored in the 27 voxels around the query.  A voxel whose
// CUBE is farther from the query than the best distance found so far cannot contain that point, so it can be
// skipped without changing the result: with q in voxel v, the gap to the cube of v + s is
//     lb^2 = sum over axes of { (upper face - q)^2 if s = +1, (q - lower face)^2 if s = -1, 0 if s = 0 }.
// The voxels are visited in the reference's own order (centre, faces, edges, corners), so the strict `<` keeps the
// same point on exact ties; a voxel is skipped only when lb^2 exceeds the current best by a safety margin that
// covers the rounding of the face coordinates (1e-6 relative + 1e-10 absolute, orders of magnitude above 1 ulp).
// Typical result on the 0.5 m / 1.0 m maps: ~3 voxels and ~15-25 candidate distances per point instead of 27 / 145+.
//
// The loop is organised around memory-level parallelism, because the path is latency-bound once pruned: the
// centre voxel first, then rounds of up to four neighbour voxels whose home hash slots are loaded together, and
// every voxel's points are read four at a time (eight independent 16-byte loads per lane).  Map data is read
// straight from L1/L2 (the map fits the 126 MB L2).
// ---------------------------------------------------------------------------------------------------------------
// PERSISTENT = true: launched cooperatively with every CTA resident; all IRLS iterations run inside this one launch,
// separated by a grid barrier whose last arriver solves the 2x2 system and updates the pose (no launches after
// convergence, no per-iteration launch gap).  PERSISTENT = false: one launch per iteration (used by the NCCL-sharded
// path, where the allreduce sits between association and solve).
// Chunked upload overlapped with the first iteration: chunk c (windows [c*windows_per_chunk, ...)) may be read once
// flags[c] == seq — the flag is copied by the same copy stream right after the chunk's data.
struct UploadArgs {
    const uint32_t *flags;  // nullptr: the scan is already resident
    uint32_t seq;
    int windows_per_chunk;  // in 32-point windows
};

// Host side of the same upload: the chunks still to be issued on the copy stream.  With the "launch_first" option only the
// first chunk goes out before the persistent kernel is launched and the rest right after the launch call; measured on B200
// this does not pay (cfg4 e2e +-0, cfg3 -6 %: the first pass is bound by the 6 MB copy itself, not by the ~16 driver
// calls in front of the launch), so the default issues every chunk first.
struct HostUpload {
    const double *src;
    double *dst;
    int64_t n, wpc;
    int issued;
};
static int issue_chunks(kicp_ctx *c, HostUpload *hu, int upto) {
    for (int k = hu->issued; k < upto; ++k) {
        const int64_t lo = std::min<int64_t>(hu->n, k * hu->wpc * 32), hi = std::min<int64_t>(hu->n, (k + 1) * hu->wpc * 32);
        hu->issued = k + 1;
        if (hi > lo)
            KICP_CUDA(cudaMemcpyAsync(hu->dst + 3 * lo, hu->src + 3 * lo, (size_t)(hi - lo) * 3 * sizeof(double), cudaMemcpyHostToDevice,
                                      c->copy_stream));
        KICP_CUDA(cudaMemcpyAsync(c->d_chunk_flags + k, c->h_chunk_tags + k, sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream));
    }
    return KICP_OK;
}

struct P2PArgs {
    P2PMailbox *peer[KICP_MAX_RANKS];
    int nranks, rank, parity;
    unsigned long long tag_base;
};

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Cross-GPU exchange fused into the grid barrier (executed by the last CTA of every rank): write the local sums of
// iteration `it` into every peer's mailbox over NVLink, raise the flags, wait for all peers, and replace st->acc by the
// sum over ranks IN RANK ORDER — every rank adds the same values in the same order, so all ranks solve for the
// identical pose.  A rank that never shows up turns into an error status after ~2 s instead of a hung GPU.
__device__ __forceinline__ void p2p_exchange(RegState *st, const P2PArgs &px, unsigned it) {
    const unsigned long long tag = px.tag_base + it + 1ull;
    const int t = threadIdx.x;
    if (t < 32) {  // warp 0 only; the loops over ranks are warp-synchronous (see kicp_device.cuh on divergence safety)
        const double mine = t < 7 ? __ldcg(&st->acc[t]) : 0.0;
        for (int r = 0; r < px.nranks; ++r) {
            if (t < 8) {
                volatile double *dst = &px.peer[r]->data[px.parity][it][px.rank][t];
                *dst = mine;
            }
            __syncwarp();
        }
        __threadfence_system();
        __syncwarp();
        if (t < px.nranks) st_release_sys(&px.peer[t]->flag[px.parity][it][px.rank], tag);
        if (t < px.nranks) {
            const unsigned long long *f = &px.peer[px.rank]->flag[px.parity][it][t];
            const long long t0 = clock64();
            while (ld_acquire_sys(f) < tag) {
                if (clock64() - t0 > 4000000000ll) {  // ~2 s: a peer is missing
                    st->status = KICP_ERR_NCCL;
                    break;
                }
            }
        }
        __syncwarp();
        double s = 0.0;
        for (int r = 0; r < px.nranks; ++r) {
            if (t < 8) s += *(volatile double *)&px.peer[px.rank]->data[px.parity][it][r][t];
            __syncwarp();
        }
        if (t < 8) __stcg(&st->acc[t], s);
    }
    __syncthreads();
}

// Warp-shuffle reduction of the 7 per-thread sums -> one partial per warp -> one partial per CTA, written with plain
// stores to partials[blockIdx.x][8].  The last CTA to arrive (ticket) sums the partials in a fixed order — no contended
// floating-point atomics, and the result does not depend on the arrival order — and then either runs the multi-launch
// tail or releases the grid barrier of the persistent kernel after solving for the next pose.
// Returns true when the kernel must return (non-persistent launch, or converged).
template <bool PERSISTENT>
__device__ __forceinline__ bool reduce_and_finish(RegState *st, double *partials, double (&v)[7], unsigned it, double (*s_part)[8],
                                                  int *s_last, const P2PArgs &px) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) s_part[wid][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double s = 0.0;
        if (threadIdx.x < 7)
            for (int k = 0; k < KICP_WARPS; ++k) s += s_part[k][threadIdx.x];
        __stcg(&partials[(size_t)blockIdx.x * 8 + threadIdx.x], s);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(&st->ticket, 1u);
        *s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    const bool last = *s_last != 0;
    const unsigned long long tb0 = gtime_ns();
    if (last) {
        __threadfence();
        if (wid < 7) {  // warp `wid` sums column `wid` over the CTAs, lanes striding, fixed tree
            double s = 0.0;
            for (unsigned b0 = 0; b0 < gridDim.x; b0 += 32) {  // warp-uniform trip count
                const unsigned b = b0 + lane;
                if (b < gridDim.x) s += __ldcg(&partials[(size_t)b * 8 + wid]);
                __syncwarp();
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
            if (lane == 0) __stcg(&st->acc[wid], s);
        }
        __syncthreads();
        if (PERSISTENT && px.nranks > 1) p2p_exchange(st, px, it);
        if (threadIdx.x == 0) {
            __threadfence();
            st->ticket = 0;
            st->window_counter = 0;
            const unsigned long long tb1 = gtime_ns();
            const int jdbg = st->iter;
            if (PERSISTENT || st->fused_tail) solve_and_update(st);
            if (jdbg < KICP_MAX_ITERATIONS) st->dbg[jdbg][2] = (double)(tb1 - tb0), st->dbg[jdbg][3] = (double)(gtime_ns() - tb1);
            if (PERSISTENT) {
                __threadfence();
                atomicExch(&st->generation, it + 1u);
            }
        }
    }
    if (!PERSISTENT) return true;
    if (threadIdx.x == 0) {
        if (!last) {
            while (*(volatile unsigned *)&st->generation <= it) {
            }
            __threadfence();
        }
        s_last[1] = __ldcg(&st->done);  // a different word than the ticket flag s_last[0] (other warps may still read it)
    }
    __syncthreads();
    return s_last[1] != 0;
}

template <bool PERSISTENT>
__global__ void __launch_bounds__(KICP_WARPS * 32, KICP_MINB) k_assoc_pruned(RegState *st, const double *__restrict__ scan, int n, MapView map, double *partials,
                                                                                 P2PArgs px, UploadArgs up) {
    if (st->done) return;
    __shared__ double s_T[12];
    __shared__ double s_part[KICP_WARPS][8];
    __shared__ int s_last[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xFFFFFFFFu;
    __shared__ MapView s_map[32];
    const MapRegs mr = map_regs(map, s_map);  // per-thread copy of the map view (divergence safety, kicp_device.cuh)
    __shared__ double s_tau[32];
    if (threadIdx.x < 32) s_tau[threadIdx.x] = st->tau;
    __syncthreads();
    const double tau = ((const volatile double *)s_tau)[lane], vs = map.voxel_size;
    const int num_windows = (n + 31) >> 5;
    // the iteration counter is read inside thread 0's spin loop of the grid barrier: keep it in a per-thread register
    // (loaded from a lane-dependent shared address, see kicp_device.cuh) rather than in a uniform register
    __shared__ unsigned s_zero[32];
    if (threadIdx.x < 32) s_zero[threadIdx.x] = 0u;
    __syncthreads();
    unsigned it = ((const volatile unsigned *)s_zero)[lane];
  for (;; ++it) {
    // the current estimate: written by k_reg_init or by the last CTA of the previous iteration -> read through L2
    if (threadIdx.x < 9) s_T[threadIdx.x] = __ldcg(&st->R[threadIdx.x]);
    if (threadIdx.x >= 9 && threadIdx.x < 12) s_T[threadIdx.x] = __ldcg(&st->t[threadIdx.x - 9]);
    __syncthreads();
    double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0, cntN = 0, ssq = 0;
    const unsigned long long t_iter0 = gtime_ns();

    // dynamic window fetch (measured 15-20 % faster than a static round-robin at cfg4: windows differ in cost)
    while (true) {
        int w = 0;
        if (lane == 0) w = (int)atomicAdd(&st->window_counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= num_windows) break;
        if (PERSISTENT && up.flags != nullptr && it == 0u) {
            // first pass over a frame that is still being uploaded: wait until this window's chunk has landed
            if (lane == 0) {
                const uint32_t *f = up.flags + min(w / up.windows_per_chunk, KICP_UPLOAD_CHUNKS - 1);
                const long long t0 = clock64();
                while (ld_acquire_sys_u32(f) != up.seq) {
                    if (clock64() - t0 > 4000000000ll) {  // ~2 s: the copy never arrived
                        st->status = KICP_ERR_CUDA;
                        break;
                    }
                }
            }
            __syncwarp();
        }
        const int i = w * 32 + lane;
        const bool valid = i < n;
        double px = 0, py = 0, pz = 0;
        if (valid) px = scan[3 * (size_t)i], py = scan[3 * (size_t)i + 1], pz = scan[3 * (size_t)i + 2];
        const double qx = s_T[0] * px + s_T[1] * py + s_T[2] * pz + s_T[9];
        const double qy = s_T[3] * px + s_T[4] * py + s_T[5] * pz + s_T[10];
        const double qz = s_T[6] * px + s_T[7] * py + s_T[8] * pz + s_T[11];
        const int vx = voxel_coord(qx, vs), vy = voxel_coord(qy, vs), vz = voxel_coord(qz, vs);
        // squared gaps to the six faces of the query voxel
        double t;
        t = (double)(vx + 1) * vs - qx; const double gxp = t * t;
        t = qx - (double)vx * vs;       const double gxm = t * t;
        t = (double)(vy + 1) * vs - qy; const double gyp = t * t;
        t = qy - (double)vy * vs;       const double gym = t * t;
        t = (double)(vz + 1) * vs - qz; const double gzp = t * t;
        t = qz - (double)vz * vs;       const double gzm = t * t;

        double best = DBL_MAX;
        const double *bestp = nullptr;
        const uint32_t tmask = mr.mask;
        const int4 *tslots = mr.slots;
        const double *tpts = mr.pts;
        const size_t tstride = (size_t)mr.cap * KICP_PSTRIDE;
        // round 0: the query's own voxel — it usually yields a best distance that prunes most of the other 26
        if (valid) {
            const uint32_t meta = map_probe(mr, vx, vy, vz);
            if (meta != KICP_SLOT_EMPTY) scan_voxel(tpts + (size_t)(meta >> 8) * tstride, (int)(meta & 0xFFu), qx, qy, qz, best, bestp);
        }
        __syncwarp();
        unsigned mask = valid ? 0x07FFFFFEu : 0u;  // shifts still to consider, bit k <-> voxel_shifts[k]
        while (__any_sync(FULL, mask != 0u)) {  // warp-uniform loop; inside a round every lane walks its own voxels
            if (mask) {
                // drop every shift whose cube is provably too far, then take the next (up to) 4 in KISS order
                const double bound = best * (1.0 + 1e-6) + 1e-10;
                mask &= (kX0 | (gxp <= bound ? kXP : 0u) | (gxm <= bound ? kXM : 0u)) &
                        (kY0 | (gyp <= bound ? kYP : 0u) | (gym <= bound ? kYM : 0u)) &
                        (kZ0 | (gzp <= bound ? kZP : 0u) | (gzm <= bound ? kZM : 0u));
                int kx[4], ky[4], kz[4];
                uint32_t hh[4];
                double lb[4];
                bool use[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    use[u] = false;
                    kx[u] = ky[u] = kz[u] = 0, hh[u] = 0, lb[u] = 0.0;
                    while (mask) {
                        const int k = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int sx = shift_x(k), sy = shift_y(k), sz = shift_z(k);
                        const double lb2 = (sx > 0 ? gxp : (sx < 0 ? gxm : 0.0)) + (sy > 0 ? gyp : (sy < 0 ? gym : 0.0)) +
                                           (sz > 0 ? gzp : (sz < 0 ? gzm : 0.0));
                        if (lb2 > bound) continue;
                        use[u] = true, lb[u] = lb2;
                        kx[u] = vx + sx, ky[u] = vy + sy, kz[u] = vz + sz;
                        hh[u] = voxel_hash(kx[u], ky[u], kz[u]) & tmask;
                        break;
                    }
                }
                // four independent home-slot loads in flight, then resolve the (rare) longer probe chains
                int4 s0[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (use[u]) s0[u] = __ldg(&tslots[hh[u]]);
                uint32_t metas[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t meta = KICP_SLOT_EMPTY;
                    if (use[u]) {
                        int4 sl = s0[u];
                        uint32_t h = hh[u];
                        while (true) {
                            if ((uint32_t)sl.w == KICP_SLOT_EMPTY) break;
                            if (sl.x == kx[u] && sl.y == ky[u] && sl.z == kz[u]) {
                                meta = (uint32_t)sl.w;
                                break;
                            }
                            h = (h + 1) & tmask;
                            sl = __ldg(&tslots[h]);
                        }
                    }
                    metas[u] = meta;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    // re-check against the best found meanwhile; KISS order is kept, so strict < resolves ties identically
                    if (metas[u] != KICP_SLOT_EMPTY && !(lb[u] > best * (1.0 + 1e-6) + 1e-10))
                        scan_voxel(tpts + (size_t)(metas[u] >> 8) * tstride, (int)(metas[u] & 0xFFu), qx, qy, qz, best, bestp);
                }
            }
            __syncwarp();
        }
        const bool have = bestp != nullptr;
        double bx = 0, by = 0, bz = 0;
        if (have) {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(bestp)), b = __ldg(reinterpret_cast<const double2 *>(bestp) + 1);
            bx = a.x, by = a.y, bz = b.x;
        }
        if (have) {
            const double rx = qx - bx, ry = qy - by, rz = qz - bz;  // r = T p - n
            const double rr = rx * rx + ry * ry + rz * rz;
            if (sqrt(rr) < tau) {  // distance < max_correspondance_distance   (Registration.cpp:75)
                const double c0x = s_T[0], c0y = s_T[3], c0z = s_T[6];
                const double c1x = s_T[1] * px - s_T[0] * py, c1y = s_T[4] * px - s_T[3] * py, c1z = s_T[7] * px - s_T[6] * py;
                a00 += c0x * c0x + c0y * c0y + c0z * c0z;
                a01 += c0x * c1x + c0y * c1y + c0z * c1z;
                a11 += c1x * c1x + c1y * c1y + c1z * c1z;
                b0 += c0x * rx + c0y * ry + c0z * rz;
                b1 += c1x * rx + c1y * ry + c1z * rz;
                cntN += 1.0;
                ssq += rr;
            }
        }
    }

    if (blockIdx.x == 0 && threadIdx.x == 0 && it < KICP_MAX_ITERATIONS) st->dbg[it][0] = (double)(gtime_ns() - t_iter0);
    double v[7] = {a00, a01, a11, b0, b1, cntN, ssq};
    if (reduce_and_finish<PERSISTENT>(st, partials, v, it, s_part, s_last, px)) return;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// k_assoc_group4 (variant 2): the same exact-pruning search with FOUR LANES PER SCAN POINT (a warp-window = 8 points).
// The pruned path is bound by the serial latency of one window (cfg2: 43 us per iteration with less than one window
// per warp), so the search of one point is spread over a 4-lane group:
//   - every lane of the group probes a different neighbour voxel (4 hash probes in flight per point, 1 per lane);
//   - the group strides over a voxel's points (lane s takes points s, s+4, ...: one coalesced 128-byte row per step);
//   - the best squared distance is min-reduced over the group (2 shuffles) between rounds so pruning stays tight.
// The winner is the minimum of (d2, order key) with order key = (KISS shift index, index in voxel), i.e. the first
// minimum in the reference's visiting order.  Windows are 4x smaller, so the end-of-launch tail shrinks as well.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double group4_min(double x) {
    const unsigned FULL = 0xFFFFFFFFu;
    x = fmin(x, __shfl_xor_sync(FULL, x, 1));
    x = fmin(x, __shfl_xor_sync(FULL, x, 2));
    return x;
}

template <bool PERSISTENT>
__global__ void __launch_bounds__(KICP_WARPS * 32, KICP_MINB) k_assoc_group4(RegState *st, const double *__restrict__ scan, int n, MapView map,
                                                                      double *partials, int pow2_voxel, P2PArgs px) {
    if (st->done) return;
    __shared__ double s_T[12];
    __shared__ double s_part[KICP_WARPS][8];
    __shared__ int s_last[2];
    const int lane = threadIdx.x & 31;
    const int sub = lane & 3;
    const unsigned FULL = 0xFFFFFFFFu;
    __shared__ MapView s_map[32];
    const MapRegs mr = map_regs(map, s_map);  // per-thread copy of the map view (divergence safety, kicp_device.cuh)
    __shared__ double s_tau[32];
    if (threadIdx.x < 32) s_tau[threadIdx.x] = st->tau;
    __syncthreads();
    const double tau = ((const volatile double *)s_tau)[lane], vs = map.voxel_size, inv_vs = 1.0 / map.voxel_size;
    const int num_windows = (n + 7) >> 3;
    __shared__ unsigned s_zero[32];
    if (threadIdx.x < 32) s_zero[threadIdx.x] = 0u;
    __syncthreads();
    unsigned it = ((const volatile unsigned *)s_zero)[lane];
    const uint32_t tmask = mr.mask;
    const int4 *tslots = mr.slots;
    const double *tpts = mr.pts;
    const size_t tstride = (size_t)mr.cap * KICP_PSTRIDE;
  for (;; ++it) {
    if (threadIdx.x < 9) s_T[threadIdx.x] = __ldcg(&st->R[threadIdx.x]);
    if (threadIdx.x >= 9 && threadIdx.x < 12) s_T[threadIdx.x] = __ldcg(&st->t[threadIdx.x - 9]);
    __syncthreads();
    double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0, cntN = 0, ssq = 0;
    const unsigned long long t_iter0 = gtime_ns();

    while (true) {
        int w = 0;
        if (lane == 0) w = (int)atomicAdd(&st->window_counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= num_windows) break;
        const int i = w * 8 + (lane >> 2);
        const bool valid = i < n;
        double px = 0, py = 0, pz = 0;
        if (valid) px = scan[3 * (size_t)i], py = scan[3 * (size_t)i + 1], pz = scan[3 * (size_t)i + 2];
        const double qx = s_T[0] * px + s_T[1] * py + s_T[2] * pz + s_T[9];
        const double qy = s_T[3] * px + s_T[4] * py + s_T[5] * pz + s_T[10];
        const double qz = s_T[6] * px + s_T[7] * py + s_T[8] * pz + s_T[11];
        // PointToVoxel: floor(q / voxel_size); for a power-of-two voxel size the product with the (exact) reciprocal
        // is the same double as the quotient, so the cheaper form is used
        int vx, vy, vz;
        if (pow2_voxel) {
            vx = (int)floor(qx * inv_vs), vy = (int)floor(qy * inv_vs), vz = (int)floor(qz * inv_vs);
        } else {
            vx = voxel_coord(qx, vs), vy = voxel_coord(qy, vs), vz = voxel_coord(qz, vs);
        }
        double t;
        t = (double)(vx + 1) * vs - qx; const double gxp = t * t;
        t = qx - (double)vx * vs;       const double gxm = t * t;
        t = (double)(vy + 1) * vs - qy; const double gyp = t * t;
        t = qy - (double)vy * vs;       const double gym = t * t;
        t = (double)(vz + 1) * vs - qz; const double gzp = t * t;
        t = qz - (double)vz * vs;       const double gzm = t * t;

        double best = DBL_MAX;            // this lane's best
        unsigned bestkey = 0xFFFFFFFFu;   // (shift index << 8) | index in voxel of this lane's best
        const double *bestp = nullptr;
        // round 0: the query's own voxel, its points strided over the 4 lanes
        if (valid) {
            const uint32_t meta = map_probe(mr, vx, vy, vz);
            if (meta != KICP_SLOT_EMPTY) {
                const double *vp = tpts + (size_t)(meta >> 8) * tstride;
                const int cnt = (int)(meta & 0xFFu);
                for (int j = sub; j < cnt; j += 4) {
                    const double *p0 = vp + (size_t)j * KICP_PSTRIDE;
                    const double2 a = __ldg(reinterpret_cast<const double2 *>(p0)), b = __ldg(reinterpret_cast<const double2 *>(p0) + 1);
                    const double dx = a.x - qx, dy = a.y - qy, dz = b.x - qz;
                    const double d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 < best) best = d2, bestkey = (unsigned)j, bestp = p0;
                }
            }
        }
        __syncwarp();
        double gbest = group4_min(best);
        unsigned mask = valid ? 0x07FFFFFEu : 0u;  // identical in the 4 lanes of a group
        while (__any_sync(FULL, mask != 0u)) {
            uint32_t meta = KICP_SLOT_EMPTY;
            int myk = 0;
            double lb[4] = {0.0, 0.0, 0.0, 0.0};
            if (mask) {
                const double bound = gbest * (1.0 + 1e-6) + 1e-10;
                mask &= (kX0 | (gxp <= bound ? kXP : 0u) | (gxm <= bound ? kXM : 0u)) &
                        (kY0 | (gyp <= bound ? kYP : 0u) | (gym <= bound ? kYM : 0u)) &
                        (kZ0 | (gzp <= bound ? kZP : 0u) | (gzm <= bound ? kZM : 0u));
                // the next (up to) 4 shifts in KISS order that survive the exact bound; lane `sub` probes the sub-th one
                int kxs = 0, kys = 0, kzs = 0;
                bool use = false;
                int found = 0;
                while (mask && found < 4) {
                    const int k = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int sx = shift_x(k), sy = shift_y(k), sz = shift_z(k);
                    const double lb2 = (sx > 0 ? gxp : (sx < 0 ? gxm : 0.0)) + (sy > 0 ? gyp : (sy < 0 ? gym : 0.0)) +
                                       (sz > 0 ? gzp : (sz < 0 ? gzm : 0.0));
                    if (lb2 > bound) continue;
                    if (found == 0) lb[0] = lb2;
                    if (found == 1) lb[1] = lb2;
                    if (found == 2) lb[2] = lb2;
                    if (found == 3) lb[3] = lb2;
                    if (found == sub) use = true, myk = k, kxs = vx + sx, kys = vy + sy, kzs = vz + sz;
                    ++found;
                }
                if (use) {
                    uint32_t h = voxel_hash(kxs, kys, kzs) & tmask;
                    while (true) {
                        const int4 sl = __ldg(&tslots[h]);
                        if ((uint32_t)sl.w == KICP_SLOT_EMPTY) break;
                        if (sl.x == kxs && sl.y == kys && sl.z == kzs) {
                            meta = (uint32_t)sl.w;
                            break;
                        }
                        h = (h + 1) & tmask;
                    }
                }
            }
            __syncwarp();
            // every lane of the group learns the 4 probe results, then the group scans the found voxels in KISS order
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t mu = __shfl_sync(FULL, meta, (lane & ~3) | u);
                const int ku = __shfl_sync(FULL, myk, (lane & ~3) | u);
                if (mu != KICP_SLOT_EMPTY && !(lb[u] > best * (1.0 + 1e-6) + 1e-10)) {
                    const double *vp = tpts + (size_t)(mu >> 8) * tstride;
                    const int cnt = (int)(mu & 0xFFu);
                    for (int j = sub; j < cnt; j += 4) {
                        const double *p0 = vp + (size_t)j * KICP_PSTRIDE;
                        const double2 a = __ldg(reinterpret_cast<const double2 *>(p0)), b = __ldg(reinterpret_cast<const double2 *>(p0) + 1);
                        const double dx = a.x - qx, dy = a.y - qy, dz = b.x - qz;
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        if (d2 < best) best = d2, bestkey = ((unsigned)ku << 8) | (unsigned)j, bestp = p0;
                    }
                }
                __syncwarp();
            }
            gbest = group4_min(best);
        }
        // the group's winner: minimum of (d2, order key) — the first minimum in the reference's visiting order
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            const double od = __shfl_xor_sync(FULL, best, o);
            const unsigned ok = __shfl_xor_sync(FULL, bestkey, o);
            const unsigned long long op = __shfl_xor_sync(FULL, (unsigned long long)bestp, o);
            if (od < best || (od == best && ok < bestkey)) best = od, bestkey = ok, bestp = (const double *)op;
        }
        if (sub == 0 && bestp != nullptr) {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(bestp)), b = __ldg(reinterpret_cast<const double2 *>(bestp) + 1);
            const double rx = qx - a.x, ry = qy - a.y, rz = qz - b.x;  // r = T p - n
            const double rr = rx * rx + ry * ry + rz * rz;
            if (sqrt(rr) < tau) {  // distance < max_correspondance_distance   (Registration.cpp:75)
                const double c0x = s_T[0], c0y = s_T[3], c0z = s_T[6];
                const double c1x = s_T[1] * px - s_T[0] * py, c1y = s_T[4] * px - s_T[3] * py, c1z = s_T[7] * px - s_T[6] * py;
                a00 += c0x * c0x + c0y * c0y + c0z * c0z;
                a01 += c0x * c1x + c0y * c1y + c0z * c1z;
                a11 += c1x * c1x + c1y * c1y + c1z * c1z;
                b0 += c0x * rx + c0y * ry + c0z * rz;
                b1 += c1x * rx + c1y * ry + c1z * rz;
                cntN += 1.0;
                ssq += rr;
            }
        }
        __syncwarp();
    }

    if (blockIdx.x == 0 && threadIdx.x == 0 && it < KICP_MAX_ITERATIONS) st->dbg[it][0] = (double)(gtime_ns() - t_iter0);
    double v[7] = {a00, a01, a11, b0, b1, cntN, ssq};
    if (reduce_and_finish<PERSISTENT>(st, partials, v, it, s_part, s_last, px)) return;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// k_assoc_hybrid (variant 3, default): one pass = 32-point windows (k_assoc_pruned's body, the efficient one for the
// bulk) followed by 8-point group windows (k_assoc_group4's body) for the last KICP_TAIL_PERCENT of the points.  Work
// units are handed out in increasing order from one counter, so the small units come last and the end-of-pass tail is
// the duration of an 8-point window (about 5 us) instead of a 32-point one (about 18 us).
// ---------------------------------------------------------------------------------------------------------------
#ifndef KICP_TAIL_PERCENT
#define KICP_TAIL_PERCENT 15
#endif
template <bool PERSISTENT>
__global__ void __launch_bounds__(KICP_WARPS * 32, KICP_MINB) k_assoc_hybrid(RegState *st, const double *__restrict__ scan, int n, MapView map,
                                                                              double *partials, int pow2_voxel, P2PArgs px, UploadArgs up) {
    if (st->done) return;
    __shared__ double s_T[12];
    __shared__ double s_part[KICP_WARPS][8];
    __shared__ int s_last[2];
    const int lane = threadIdx.x & 31;
    const int sub = lane & 3;
    const unsigned FULL = 0xFFFFFFFFu;
    __shared__ MapView s_map[32];
    const MapRegs mr = map_regs(map, s_map);  // per-thread copy of the map view (divergence safety, kicp_device.cuh)
    __shared__ double s_tau[32];
    if (threadIdx.x < 32) s_tau[threadIdx.x] = st->tau;
    __syncthreads();
    const double tau = ((const volatile double *)s_tau)[lane], vs = map.voxel_size, inv_vs = 1.0 / map.voxel_size;
    // bulk: 32-point windows over the first points; tail: 8-point units over the rest
    const int all32 = (n + 31) >> 5;
    const int num_big = n >= 4096 ? (int)((long long)all32 * (100 - KICP_TAIL_PERCENT) / 100) : all32;
    const int base8 = min(n, num_big * 32);
    const int num_units = num_big + ((n - base8 + 7) >> 3);
    __shared__ unsigned s_zero[32];
    if (threadIdx.x < 32) s_zero[threadIdx.x] = 0u;
    __syncthreads();
    unsigned it = ((const volatile unsigned *)s_zero)[lane];
    const uint32_t tmask = mr.mask;
    const int4 *tslots = mr.slots;
    const double *tpts = mr.pts;
    const size_t tstride = (size_t)mr.cap * KICP_PSTRIDE;
  for (;; ++it) {
    if (threadIdx.x < 9) s_T[threadIdx.x] = __ldcg(&st->R[threadIdx.x]);
    if (threadIdx.x >= 9 && threadIdx.x < 12) s_T[threadIdx.x] = __ldcg(&st->t[threadIdx.x - 9]);
    __syncthreads();
    double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0, cntN = 0, ssq = 0;
    const unsigned long long t_iter0 = gtime_ns();

    while (true) {
        int w = 0;
        if (lane == 0) w = (int)atomicAdd(&st->window_counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= num_units) break;
        if (PERSISTENT && up.flags != nullptr && it == 0u) {
            // first pass over a frame that is still being uploaded: wait until this unit's chunk has landed
            if (lane == 0) {
                const int w32 = w < num_big ? w : (base8 + (w - num_big) * 8) >> 5;
                const uint32_t *f = up.flags + min(w32 / up.windows_per_chunk, KICP_UPLOAD_CHUNKS - 1);
                const long long t0 = clock64();
                while (ld_acquire_sys_u32(f) != up.seq) {
                    if (clock64() - t0 > 4000000000ll) {
                        st->status = KICP_ERR_CUDA;
                        break;
                    }
                }
            }
            __syncwarp();
        }
        if (w < num_big) {
        const int i = w * 32 + lane;
        const bool valid = i < n;
        double px = 0, py = 0, pz = 0;
        if (valid) px = scan[3 * (size_t)i], py = scan[3 * (size_t)i + 1], pz = scan[3 * (size_t)i + 2];
        const double qx = s_T[0] * px + s_T[1] * py + s_T[2] * pz + s_T[9];
        const double qy = s_T[3] * px + s_T[4] * py + s_T[5] * pz + s_T[10];
        const double qz = s_T[6] * px + s_T[7] * py + s_T[8] * pz + s_T[11];
        const int vx = voxel_coord(qx, vs), vy = voxel_coord(qy, vs), vz = voxel_coord(qz, vs);
        // squared gaps to the six faces of the query voxel
        double t;
        t = (double)(vx + 1) * vs - qx; const double gxp = t * t;
        t = qx - (double)vx * vs;       const double gxm = t * t;
        t = (double)(vy + 1) * vs - qy; const double gyp = t * t;
        t = qy - (double)vy * vs;       const double gym = t * t;
        t = (double)(vz + 1) * vs - qz; const double gzp = t * t;
        t = qz - (double)vz * vs;       const double gzm = t * t;

        double best = DBL_MAX;
        const double *bestp = nullptr;
        const uint32_t tmask = mr.mask;
        const int4 *tslots = mr.slots;
        const double *tpts = mr.pts;
        const size_t tstride = (size_t)mr.cap * KICP_PSTRIDE;
        // round 0: the query's own voxel — it usually yields a best distance that prunes most of the other 26
        if (valid) {
            const uint32_t meta = map_probe(mr, vx, vy, vz);
            if (meta != KICP_SLOT_EMPTY) scan_voxel(tpts + (size_t)(meta >> 8) * tstride, (int)(meta & 0xFFu), qx, qy, qz, best, bestp);
        }
        __syncwarp();
        unsigned mask = valid ? 0x07FFFFFEu : 0u;  // shifts still to consider, bit k <-> voxel_shifts[k]
        while (__any_sync(FULL, mask != 0u)) {  // warp-uniform loop; inside a round every lane walks its own voxels
            if (mask) {
                // drop every shift whose cube is provably too far, then take the next (up to) 4 in KISS order
                const double bound = best * (1.0 + 1e-6) + 1e-10;
                mask &= (kX0 | (gxp <= bound ? kXP : 0u) | (gxm <= bound ? kXM : 0u)) &
                        (kY0 | (gyp <= bound ? kYP : 0u) | (gym <= bound ? kYM : 0u)) &
                        (kZ0 | (gzp <= bound ? kZP : 0u) | (gzm <= bound ? kZM : 0u));
                int kx[4], ky[4], kz[4];
                uint32_t hh[4];
                double lb[4];
                bool use[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    use[u] = false;
                    kx[u] = ky[u] = kz[u] = 0, hh[u] = 0, lb[u] = 0.0;
                    while (mask) {
                        const int k = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int sx = shift_x(k), sy = shift_y(k), sz = shift_z(k);
                        const double lb2 = (sx > 0 ? gxp : (sx < 0 ? gxm : 0.0)) + (sy > 0 ? gyp : (sy < 0 ? gym : 0.0)) +
                                           (sz > 0 ? gzp : (sz < 0 ? gzm : 0.0));
                        if (lb2 > bound) continue;
                        use[u] = true, lb[u] = lb2;
                        kx[u] = vx + sx, ky[u] = vy + sy, kz[u] = vz + sz;
                        hh[u] = voxel_hash(kx[u], ky[u], kz[u]) & tmask;
                        break;
                    }
                }
                // four independent home-slot loads in flight, then resolve the (rare) longer probe chains
                int4 s0[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (use[u]) s0[u] = __ldg(&tslots[hh[u]]);
                uint32_t metas[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t meta = KICP_SLOT_EMPTY;
                    if (use[u]) {
                        int4 sl = s0[u];
                        uint32_t h = hh[u];
                        while (true) {
                            if ((uint32_t)sl.w == KICP_SLOT_EMPTY) break;
                            if (sl.x == kx[u] && sl.y == ky[u] && sl.z == kz[u]) {
                                meta = (uint32_t)sl.w;
                                break;
                            }
                            h = (h + 1) & tmask;
                            sl = __ldg(&tslots[h]);
                        }
                    }
                    metas[u] = meta;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    // re-check against the best found meanwhile; KISS order is kept, so strict < resolves ties identically
                    if (metas[u] != KICP_SLOT_EMPTY && !(lb[u] > best * (1.0 + 1e-6) + 1e-10))
                        scan_voxel(tpts + (size_t)(metas[u] >> 8) * tstride, (int)(metas[u] & 0xFFu), qx, qy, qz, best, bestp);
                }
            }
            __syncwarp();
        }
        const bool have = bestp != nullptr;
        double bx = 0, by = 0, bz = 0;
        if (have) {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(bestp)), b = __ldg(reinterpret_cast<const double2 *>(bestp) + 1);
            bx = a.x, by = a.y, bz = b.x;
        }
        if (have) {
            const double rx = qx - bx, ry = qy - by, rz = qz - bz;  // r = T p - n
            const double rr = rx * rx + ry * ry + rz * rz;
            if (sqrt(rr) < tau) {  // distance < max_correspondance_distance   (Registration.cpp:75)
                const double c0x = s_T[0], c0y = s_T[3], c0z = s_T[6];
                const double c1x = s_T[1] * px - s_T[0] * py, c1y = s_T[4] * px - s_T[3] * py, c1z = s_T[7] * px - s_T[6] * py;
                a00 += c0x * c0x + c0y * c0y + c0z * c0z;
                a01 += c0x * c1x + c0y * c1y + c0z * c1z;
                a11 += c1x * c1x + c1y * c1y + c1z * c1z;
                b0 += c0x * rx + c0y * ry + c0z * rz;
                b1 += c1x * rx + c1y * ry + c1z * rz;
                cntN += 1.0;
                ssq += rr;
            }
        }

        } else {
        const int i = base8 + (w - num_big) * 8 + (lane >> 2);
        const bool valid = i < n;
        double px = 0, py = 0, pz = 0;
        if (valid) px = scan[3 * (size_t)i], py = scan[3 * (size_t)i + 1], pz = scan[3 * (size_t)i + 2];
        const double qx = s_T[0] * px + s_T[1] * py + s_T[2] * pz + s_T[9];
        const double qy = s_T[3] * px + s_T[4] * py + s_T[5] * pz + s_T[10];
        const double qz = s_T[6] * px + s_T[7] * py + s_T[8] * pz + s_T[11];
        // PointToVoxel: floor(q / voxel_size); for a power-of-two voxel size the product with the (exact) reciprocal
        // is the same double as the quotient, so the cheaper form is used
        int vx, vy, vz;
        if (pow2_voxel) {
            vx = (int)floor(qx * inv_vs), vy = (int)floor(qy * inv_vs), vz = (int)floor(qz * inv_vs);
        } else {
            vx = voxel_coord(qx, vs), vy = voxel_coord(qy, vs), vz = voxel_coord(qz, vs);
        }
        double t;
        t = (double)(vx + 1) * vs - qx; const double gxp = t * t;
        t = qx - (double)vx * vs;       const double gxm = t * t;
        t = (double)(vy + 1) * vs - qy; const double gyp = t * t;
        t = qy - (double)vy * vs;       const double gym = t * t;
        t = (double)(vz + 1) * vs - qz; const double gzp = t * t;
        t = qz - (double)vz * vs;       const double gzm = t * t;

        double best = DBL_MAX;            // this lane's best
        unsigned bestkey = 0xFFFFFFFFu;   // (shift index << 8) | index in voxel of this lane's best
        const double *bestp = nullptr;
        // round 0: the query's own voxel, its points strided over the 4 lanes
        if (valid) {
            const uint32_t meta = map_probe(mr, vx, vy, vz);
            if (meta != KICP_SLOT_EMPTY) {
                const double *vp = tpts + (size_t)(meta >> 8) * tstride;
                const int cnt = (int)(meta & 0xFFu);
                for (int j = sub; j < cnt; j += 4) {
                    const double *p0 = vp + (size_t)j * KICP_PSTRIDE;
                    const double2 a = __ldg(reinterpret_cast<const double2 *>(p0)), b = __ldg(reinterpret_cast<const double2 *>(p0) + 1);
                    const double dx = a.x - qx, dy = a.y - qy, dz = b.x - qz;
                    const double d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 < best) best = d2, bestkey = (unsigned)j, bestp = p0;
                }
            }
        }
        __syncwarp();
        double gbest = group4_min(best);
        unsigned mask = valid ? 0x07FFFFFEu : 0u;  // identical in the 4 lanes of a group
        while (__any_sync(FULL, mask != 0u)) {
            uint32_t meta = KICP_SLOT_EMPTY;
            int myk = 0;
            double lb[4] = {0.0, 0.0, 0.0, 0.0};
            if (mask) {
                const double bound = gbest * (1.0 + 1e-6) + 1e-10;
                mask &= (kX0 | (gxp <= bound ? kXP : 0u) | (gxm <= bound ? kXM : 0u)) &
                        (kY0 | (gyp <= bound ? kYP : 0u) | (gym <= bound ? kYM : 0u)) &
                        (kZ0 | (gzp <= bound ? kZP : 0u) | (gzm <= bound ? kZM : 0u));
                // the next (up to) 4 shifts in KISS order that survive the exact bound; lane `sub` probes the sub-th one
                int kxs = 0, kys = 0, kzs = 0;
                bool use = false;
                int found = 0;
                while (mask && found < 4) {
                    const int k = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int sx = shift_x(k), sy = shift_y(k), sz = shift_z(k);
                    const double lb2 = (sx > 0 ? gxp : (sx < 0 ? gxm : 0.0)) + (sy > 0 ? gyp : (sy < 0 ? gym : 0.0)) +
                                       (sz > 0 ? gzp : (sz < 0 ? gzm : 0.0));
                    if (lb2 > bound) continue;
                    if (found == 0) lb[0] = lb2;
                    if (found == 1) lb[1] = lb2;
                    if (found == 2) lb[2] = lb2;
                    if (found == 3) lb[3] = lb2;
                    if (found == sub) use = true, myk = k, kxs = vx + sx, kys = vy + sy, kzs = vz + sz;
                    ++found;
                }
                if (use) {
                    uint32_t h = voxel_hash(kxs, kys, kzs) & tmask;
                    while (true) {
                        const int4 sl = __ldg(&tslots[h]);
                        if ((uint32_t)sl.w == KICP_SLOT_EMPTY) break;
                        if (sl.x == kxs && sl.y == kys && sl.z == kzs) {
                            meta = (uint32_t)sl.w;
                            break;
                        }
                        h = (h + 1) & tmask;
                    }
                }
            }
            __syncwarp();
            // every lane of the group learns the 4 probe results, then the group scans the found voxels in KISS order
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t mu = __shfl_sync(FULL, meta, (lane & ~3) | u);
                const int ku = __shfl_sync(FULL, myk, (lane & ~3) | u);
                if (mu != KICP_SLOT_EMPTY && !(lb[u] > best * (1.0 + 1e-6) + 1e-10)) {
                    const double *vp = tpts + (size_t)(mu >> 8) * tstride;
                    const int cnt = (int)(mu & 0xFFu);
                    for (int j = sub; j < cnt; j += 4) {
                        const double *p0 = vp + (size_t)j * KICP_PSTRIDE;
                        const double2 a = __ldg(reinterpret_cast<const double2 *>(p0)), b = __ldg(reinterpret_cast<const double2 *>(p0) + 1);
                        const double dx = a.x - qx, dy = a.y - qy, dz = b.x - qz;
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        if (d2 < best) best = d2, bestkey = ((unsigned)ku << 8) | (unsigned)j, bestp = p0;
                    }
                }
                __syncwarp();
            }
            gbest = group4_min(best);
        }
        // the group's winner: minimum of (d2, order key) — the first minimum in the reference's visiting order
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            const double od = __shfl_xor_sync(FULL, best, o);
            const unsigned ok = __shfl_xor_sync(FULL, bestkey, o);
            const unsigned long long op = __shfl_xor_sync(FULL, (unsigned long long)bestp, o);
            if (od < best || (od == best && ok < bestkey)) best = od, bestkey = ok, bestp = (const double *)op;
        }
        if (sub == 0 && bestp != nullptr) {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(bestp)), b = __ldg(reinterpret_cast<const double2 *>(bestp) + 1);
            const double rx = qx - a.x, ry = qy - a.y, rz = qz - b.x;  // r = T p - n
            const double rr = rx * rx + ry * ry + rz * rz;
            if (sqrt(rr) < tau) {  // distance < max_correspondance_distance   (Registration.cpp:75)
                const double c0x = s_T[0], c0y = s_T[3], c0z = s_T[6];
                const double c1x = s_T[1] * px - s_T[0] * py, c1y = s_T[4] * px - s_T[3] * py, c1z = s_T[7] * px - s_T[6] * py;
                a00 += c0x * c0x + c0y * c0y + c0z * c0z;
                a01 += c0x * c1x + c0y * c1y + c0z * c1z;
                a11 += c1x * c1x + c1y * c1y + c1z * c1z;
                b0 += c0x * rx + c0y * ry + c0z * rz;
                b1 += c1x * rx + c1y * ry + c1z * rz;
                cntN += 1.0;
                ssq += rr;
            }
        }
        __syncwarp();

        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && it < KICP_MAX_ITERATIONS) st->dbg[it][0] = (double)(gtime_ns() - t_iter0);
    double v[7] = {a00, a01, a11, b0, b1, cntN, ssq};
    if (reduce_and_finish<PERSISTENT>(st, partials, v, it, s_part, s_last, px)) return;
  }
}

// ------------------------------------------------------------------------------------------------------- host
// Kernel variant and binning granularity are per-context options (kicp_ctx_set_option); the defaults are the
// measured best (profiles/) and can be overridden with KICP_ASSOC=staged|pruned and KICP_SORT_BITS=0..30 so that
// bench.py / ncu can compare variants on the same inputs.
extern "C" int kicp_ctx_set_option(kicp_ctx *c, const char *name, int32_t value) {
    if (!c || !name) return KICP_ERR_INVALID;
    if (!strcmp(name, "assoc_variant")) {
        if (value < 0 || value > 3) return KICP_ERR_INVALID;
        c->assoc_variant = value;
    } else if (!strcmp(name, "persistent")) {
        if (value != 0 && value != 1) return KICP_ERR_INVALID;
        c->persistent = value;
    } else if (!strcmp(name, "sort_bits")) {
        if (value < 0 || value > 30) return KICP_ERR_INVALID;
        c->sort_bits = value;
    } else if (!strcmp(name, "launch_first")) {
        if (value != 0 && value != 1) return KICP_ERR_INVALID;
        c->launch_first = value;
    } else if (!strcmp(name, "group4_below")) {
        if (value < 0) return KICP_ERR_INVALID;
        c->group4_below = value;
    } else {
        return KICP_ERR_INVALID;
    }
    return KICP_OK;
}

static size_t assoc_smem_bytes() {
    return (size_t)KICP_WARPS * ((3 * KICP_CH + 3 * 32 + 32) * sizeof(double) + (KICP_CH + 32) * sizeof(int));
}

static int reg_reserve(kicp_ctx *c, int64_t n) {
    if (!c->d_state) {
        KICP_CUDA(cudaMalloc(&c->d_state, sizeof(RegState)));
        KICP_CUDA(cudaMalloc(&c->d_partials, (size_t)c->sm_count * 16 * 8 * sizeof(double)));
        KICP_CUDA(cudaFuncSetAttribute(k_assoc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)assoc_smem_bytes()));
        int per_sm = 0;
        KICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assoc, KICP_WARPS * 32, assoc_smem_bytes()));
        c->assoc_ctas_per_sm = std::max(per_sm, 1);
        KICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assoc_pruned<false>, KICP_WARPS * 32, 0));
        c->pruned_ctas_per_sm = std::max(per_sm, 1);
        KICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assoc_pruned<true>, KICP_WARPS * 32, 0));
        c->persistent_ctas_per_sm = std::max(per_sm, 1);
        KICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assoc_group4<true>, KICP_WARPS * 32, 0));
        c->group4_ctas_per_sm = std::max(per_sm, 1);
        KICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assoc_group4<false>, KICP_WARPS * 32, 0));
        c->group4_ctas_per_sm = std::min(c->group4_ctas_per_sm, std::max(per_sm, 1));
        KICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assoc_hybrid<true>, KICP_WARPS * 32, 0));
        c->hybrid_ctas_per_sm = std::max(per_sm, 1);
        KICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assoc_hybrid<false>, KICP_WARPS * 32, 0));
        c->hybrid_ctas_per_sm = std::min(c->hybrid_ctas_per_sm, std::max(per_sm, 1));
        int coop = 0;
        KICP_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
        if (!coop) c->persistent = 0;
    }
    if (n <= c->scratch_cap) return KICP_OK;
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_sorted), cudaFree(c->d_keys), cudaFree(c->d_keys_alt), cudaFree(c->d_idx), cudaFree(c->d_idx_alt);
    cudaFree(c->d_sort_tmp);
    c->d_sorted = nullptr, c->d_keys = c->d_keys_alt = nullptr, c->d_idx = c->d_idx_alt = nullptr, c->d_sort_tmp = nullptr;
    const int64_t cap = std::max<int64_t>(n + n / 4, 4096);
    KICP_CUDA(cudaMalloc(&c->d_sorted, (size_t)cap * 3 * sizeof(double)));
    KICP_CUDA(cudaMalloc(&c->d_keys, (size_t)cap * sizeof(uint32_t)));
    KICP_CUDA(cudaMalloc(&c->d_keys_alt, (size_t)cap * sizeof(uint32_t)));
    KICP_CUDA(cudaMalloc(&c->d_idx, (size_t)cap * sizeof(int32_t)));
    KICP_CUDA(cudaMalloc(&c->d_idx_alt, (size_t)cap * sizeof(int32_t)));
    size_t bytes = 0;
    KICP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, c->d_keys, c->d_keys_alt, c->d_idx, c->d_idx_alt, (int)cap, 0, 30,
                                              c->stream));
    KICP_CUDA(cudaMalloc(&c->d_sort_tmp, bytes));
    c->sort_tmp_bytes = bytes;
    c->scratch_cap = cap;
    return KICP_OK;
}

static int check_params(const kicp_reg_params *p) {
    if (!p) return KICP_ERR_INVALID;
    if (p->max_num_iterations > KICP_MAX_ITERATIONS) {
        kicp_set_error("max_num_iterations exceeds KICP_MAX_ITERATIONS");
        return KICP_ERR_INVALID;
    }
    return KICP_OK;
}

// Enqueue one full registration on the context stream.  `sharded` inserts the 8-double allreduce between the
// association and the solve of every iteration.
static int enqueue_registration(kicp_map *m, kicp_scan *scan, const double last[7], const double odom[7], double tau,
                                const kicp_reg_params *p, kicp_reg_result *result, bool sharded, const UploadArgs *upload = nullptr,
                                HostUpload *host_upload = nullptr) {
    if (!m || !scan || !last || !odom) return KICP_ERR_INVALID;
    KICP_TRY(check_params(p));
    kicp_ctx *c = m->ctx;
    if (scan->ctx != c) return KICP_ERR_INVALID;
    if (scan->n > 0x7FFFFFE0ll) return KICP_ERR_CAPACITY;
    if (sharded && !c->nccl_comm && !c->p2p_ready) {
        kicp_set_error("kicp_register_sharded: neither kicp_comm_p2p_init nor kicp_comm_init has been called on this context");
        return KICP_ERR_INVALID;
    }
    KICP_CUDA(cudaSetDevice(c->device));
    const int n = (int)scan->n;
    KICP_TRY(reg_reserve(c, n));
    RegArgs a;
    a.last = Pose{last[0], last[1], last[2], last[3], last[4], last[5], last[6]};
    a.odom = Pose{odom[0], odom[1], odom[2], odom[3], odom[4], odom[5], odom[6]};
    a.tau = tau, a.conv = p->convergence_criterion, a.fixed_reg = p->fixed_regularization;
    a.adaptive = p->use_adaptive_odometry_regularization ? 1 : 0;
    // an empty map returns the prediction (Registration.cpp:157): no association, no solve
    a.max_iter = m->num_blocks == 0 ? 0 : p->max_num_iterations;
    // small scans leave most warps without a 32-point window: the 8-point-window kernel spreads them over the whole machine
    const int variant = (c->assoc_variant == 1 && !sharded && n > 0 && n <= c->group4_below) ? 2 : c->assoc_variant;
    a.fused_tail = (sharded && !(c->p2p_ready && variant >= 1 && c->persistent)) ? 0 : 1;
    a.iters_out = nullptr;
    kicp_ctx::ProfReg *pr = nullptr;
    if (c->profiling && (int64_t)c->prof.size() < c->prof_cap) {
        c->prof.emplace_back();
        pr = &c->prof.back();
        pr->d_iters = c->d_prof_iters + (c->prof.size() - 1);
        a.iters_out = pr->d_iters;
        KICP_CUDA(cudaEventCreate(&pr->prep0));
        KICP_CUDA(cudaEventCreate(&pr->prep1));
        KICP_CUDA(cudaEventRecord(pr->prep0, c->stream));
    }
    k_reg_init<<<1, 32, 0, c->stream>>>(c->d_state, a);
    KICP_CHECK_LAUNCH(c);
    if (a.max_iter > 0) {
        const double *d_pts = scan->d_xyz;
        const int sbits = c->sort_bits;
        if (upload && (sbits > 0 || !((variant == 1 || variant == 3) && c->persistent && (!sharded || c->p2p_ready)))) {
            // this configuration reads the whole frame up front: wait for the upload instead of overlapping it
            if (host_upload) KICP_TRY(issue_chunks(c, host_upload, KICP_UPLOAD_CHUNKS));
            cudaEvent_t ev;
            KICP_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            KICP_CUDA(cudaEventRecord(ev, c->copy_stream));
            KICP_CUDA(cudaStreamWaitEvent(c->stream, ev, 0));
            KICP_CUDA(cudaEventDestroy(ev));
            upload = nullptr;
        }
        if (n > 0 && sbits > 0) {
            const int threads = 256, blocks = (n + threads - 1) / threads;
            k_morton_keys<<<blocks, threads, 0, c->stream>>>(c->d_state, scan->d_xyz, n, m->voxel_size, c->d_keys, c->d_idx);
            KICP_CHECK_LAUNCH(c);
            size_t bytes = c->sort_tmp_bytes;
            KICP_CUDA(cub::DeviceRadixSort::SortPairs(c->d_sort_tmp, bytes, c->d_keys, c->d_keys_alt, c->d_idx, c->d_idx_alt, n,
                                                      30 - sbits, 30, c->stream));
            c->launches += 3;  // CUB's histogram + onesweep passes (library kernels, not counted individually)
            k_gather<<<blocks, threads, 0, c->stream>>>(scan->d_xyz, c->d_idx_alt, n, c->d_sorted);
            KICP_CHECK_LAUNCH(c);
            d_pts = c->d_sorted;
        }
        if (pr) KICP_CUDA(cudaEventRecord(pr->prep1, c->stream));
        const int num_windows = (n + 31) / 32;
        // persistent-style grid: every CTA is resident and pulls windows from a device-side counter
        const bool group4 = variant == 2;
        const bool hybrid = variant == 3;
        const bool pruned = variant >= 1;
        const bool p2p = sharded && c->p2p_ready && pruned;
        const bool persistent = pruned && c->persistent && (!sharded || p2p);
        P2PArgs px{};
        px.nranks = 1;
        if (p2p) {
            for (int r = 0; r < c->nranks; ++r) px.peer[r] = c->p2p_peer[r];
            px.nranks = c->nranks, px.rank = c->rank;
            px.parity = (int)(c->p2p_seq & 1ull);
            px.tag_base = (c->p2p_seq + 1ull) * 128ull;
            c->p2p_seq++;
        }
        const int per_sm = hybrid ? c->hybrid_ctas_per_sm : group4 ? c->group4_ctas_per_sm
                                  : (persistent ? c->persistent_ctas_per_sm : (pruned ? c->pruned_ctas_per_sm : c->assoc_ctas_per_sm));
        const int units = group4 ? (n + 7) / 8 : num_windows;  // warp-windows of 8 or 32 points
        int grid = std::max(1, std::min((units + KICP_WARPS - 1) / KICP_WARPS, c->sm_count * std::min(per_sm, 16)));
        int pow2_voxel = 0;
        {
            int e = 0;
            pow2_voxel = std::frexp(m->voxel_size, &e) == 0.5 ? 1 : 0;
        }
        const bool dbg = getenv("KICP_DEBUG_SYNC") != nullptr;
        if (persistent) {
            // one cooperative launch runs every iteration (grid barrier inside the kernel)
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (pr) {
                KICP_CUDA(cudaEventCreate(&e0));
                KICP_CUDA(cudaEventCreate(&e1));
                pr->it.push_back(e0), pr->it.push_back(e1);
                pr->persistent = true;
                KICP_CUDA(cudaEventRecord(e0, c->stream));
            }
            RegState *st_arg = c->d_state;
            const double *pts_arg = d_pts;
            int n_arg = n;
            MapView mv = m->view();
            double *part_arg = c->d_partials;
            int pow2_arg = pow2_voxel;
            void *args_g4[] = {&st_arg, &pts_arg, &n_arg, &mv, &part_arg, &pow2_arg, &px};
            UploadArgs up_arg = upload ? *upload : UploadArgs{nullptr, 0u, 1};
            if (group4 || sbits > 0) up_arg.flags = nullptr;
            void *args_hy[] = {&st_arg, &pts_arg, &n_arg, &mv, &part_arg, &pow2_arg, &px, &up_arg};
            void *args_pr[] = {&st_arg, &pts_arg, &n_arg, &mv, &part_arg, &px, &up_arg};
            void **args = hybrid ? args_hy : (group4 ? args_g4 : args_pr);
            KICP_CUDA(cudaLaunchCooperativeKernel(hybrid   ? (const void *)k_assoc_hybrid<true>
                                                  : group4 ? (const void *)k_assoc_group4<true>
                                                           : (const void *)k_assoc_pruned<true>,
                                                  dim3(grid), dim3(KICP_WARPS * 32), args, 0, c->stream));
            c->launches++;
            if (host_upload) KICP_TRY(issue_chunks(c, host_upload, KICP_UPLOAD_CHUNKS));  // the kernel is already waiting on the flags
            if (pr) KICP_CUDA(cudaEventRecord(e1, c->stream));
            if (dbg) {
                cudaError_t e = cudaStreamSynchronize(c->stream);
                fprintf(stderr, "[kicp] persistent launch (grid %d, n %d): %s\n", grid, n, cudaGetErrorString(e));
            }
        } else {
            for (int j = 0; j < a.max_iter; ++j) {
                cudaEvent_t e0 = nullptr, e1 = nullptr;
                if (pr) {
                    KICP_CUDA(cudaEventCreate(&e0));
                    KICP_CUDA(cudaEventCreate(&e1));
                    pr->it.push_back(e0), pr->it.push_back(e1);
                    KICP_CUDA(cudaEventRecord(e0, c->stream));
                }
                if (hybrid)
                    k_assoc_hybrid<false><<<grid, KICP_WARPS * 32, 0, c->stream>>>(c->d_state, d_pts, n, m->view(), c->d_partials,
                                                                                   pow2_voxel, px, UploadArgs{nullptr, 0u, 1});
                else if (group4)
                    k_assoc_group4<false><<<grid, KICP_WARPS * 32, 0, c->stream>>>(c->d_state, d_pts, n, m->view(), c->d_partials,
                                                                                   pow2_voxel, px);
                else if (pruned)
                    k_assoc_pruned<false><<<grid, KICP_WARPS * 32, 0, c->stream>>>(c->d_state, d_pts, n, m->view(), c->d_partials, px,
                                                                                   UploadArgs{nullptr, 0u, 1});
                else
                    k_assoc<<<grid, KICP_WARPS * 32, assoc_smem_bytes(), c->stream>>>(c->d_state, d_pts, n, m->view());
                KICP_CHECK_LAUNCH(c);
                if (pr) KICP_CUDA(cudaEventRecord(e1, c->stream));
                if (dbg) {  // debugging aid: locate a misbehaving launch
                    fprintf(stderr, "[kicp] assoc launch %d (variant %d, grid %d, n %d) ...", j, pruned ? 1 : 0, grid, n);
                    cudaError_t e = cudaStreamSynchronize(c->stream);
                    RegState hs;
                    cudaMemcpy(&hs, c->d_state, offsetof(RegState, result), cudaMemcpyDeviceToHost);
                    fprintf(stderr, " %s iter=%d done=%d ticket=%u wc=%u\n", cudaGetErrorString(e), hs.iter, hs.done, hs.ticket,
                            hs.window_counter);
                }
                if (sharded) {
                    KICP_TRY(kicp_comm_allreduce8(c, c->d_state->acc));
                    k_solve<<<1, 32, 0, c->stream>>>(c->d_state);
                    KICP_CHECK_LAUNCH(c);
                }
            }
        }
    }
    if (host_upload) KICP_TRY(issue_chunks(c, host_upload, KICP_UPLOAD_CHUNKS));  // paths that did not need the frame yet
    if (result)
        KICP_CUDA(cudaMemcpyAsync(result, &c->d_state->result, sizeof(kicp_reg_result), cudaMemcpyDeviceToHost, c->stream));
    return KICP_OK;
}

// debugging aid (not part of the public header): per-iteration device timings of the last registration
extern "C" int kicp_debug_last_timing(kicp_ctx *c, double *out /* [KICP_MAX_ITERATIONS][4] */) {
    if (!c || !c->d_state) return KICP_ERR_INVALID;
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    KICP_CUDA(cudaMemcpy(out, (const char *)c->d_state + offsetof(RegState, dbg), sizeof(double) * KICP_MAX_ITERATIONS * 4,
                         cudaMemcpyDeviceToHost));
    return KICP_OK;
}

extern "C" int kicp_ctx_profile_begin(kicp_ctx *c) {
    if (!c) return KICP_ERR_INVALID;
    KICP_CUDA(cudaSetDevice(c->device));
    if (!c->d_prof_iters) {
        c->prof_cap = 1 << 16;
        KICP_CUDA(cudaMalloc(&c->d_prof_iters, (size_t)c->prof_cap * sizeof(int32_t)));
    }
    c->prof.clear();
    c->prof.reserve(4096);
    c->profiling = true;
    return KICP_OK;
}

extern "C" int kicp_ctx_profile_end(kicp_ctx *c, kicp_profile *out) {
    if (!c || !out) return KICP_ERR_INVALID;
    KICP_CUDA(cudaSetDevice(c->device));
    c->profiling = false;
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    kicp_profile p{};
    std::vector<int32_t> iters(c->prof.size());
    if (!iters.empty())
        KICP_CUDA(cudaMemcpy(iters.data(), c->d_prof_iters, iters.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < c->prof.size(); ++r) {
        kicp_ctx::ProfReg &pr = c->prof[r];
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pr.prep0, pr.prep1) == cudaSuccess) p.prep_ms += ms;
        p.assoc_iterations += iters[r];
        for (size_t k = 0; k + 1 < pr.it.size(); k += 2) {
            ms = 0.f;
            cudaEventElapsedTime(&ms, pr.it[k], pr.it[k + 1]);
            if (pr.persistent || (int)(k / 2) < iters[r]) {
                p.assoc_ms += ms, p.assoc_launches++;
            } else {
                p.idle_ms += ms, p.idle_launches++;
            }
            cudaEventDestroy(pr.it[k]), cudaEventDestroy(pr.it[k + 1]);
        }
        cudaEventDestroy(pr.prep0), cudaEventDestroy(pr.prep1);
        p.registrations++;
    }
    cudaGetLastError();
    c->prof.clear();
    *out = p;
    return KICP_OK;
}

extern "C" int kicp_register_scan_async(kicp_map *map, kicp_scan *scan, const double last[7], const double odom[7], double tau,
                                        const kicp_reg_params *params, kicp_reg_result *result) {
    return enqueue_registration(map, scan, last, odom, tau, params, result, false);
}
extern "C" int kicp_register_scan_sharded_async(kicp_map *map, kicp_scan *scan, const double last[7], const double odom[7],
                                                double tau, const kicp_reg_params *params, kicp_reg_result *result) {
    return enqueue_registration(map, scan, last, odom, tau, params, result, true);
}

int kicp_enqueue_registration_device(kicp_map *m, const double *d_xyz, int64_t n, const double last[7], const double odom[7],
                                     double tau, const kicp_reg_params *p) {
    if (!m || n < 0 || (n > 0 && !d_xyz)) return KICP_ERR_INVALID;
    kicp_scan view;  // non-owning alias of the caller's device buffer
    view.ctx = m->ctx, view.d_xyz = const_cast<double *>(d_xyz), view.cap = n, view.n = n;
    return enqueue_registration(m, &view, last, odom, tau, p, m->ctx->h_result, false);
}

static int register_host(kicp_map *map, const double *frame_xyz, int64_t n, const double last[7], const double odom[7],
                         double tau, const kicp_reg_params *params, double out_pose[7], kicp_reg_result *result, bool sharded) {
    if (!map || n < 0 || (n > 0 && !frame_xyz) || !out_pose) return KICP_ERR_INVALID;
    kicp_ctx *c = map->ctx;
    KICP_CUDA(cudaSetDevice(c->device));
    if (!c->upload_scan) KICP_TRY(kicp_scan_create(c, n, &c->upload_scan));
    kicp_scan *s = c->upload_scan;
    KICP_TRY(kicp_scan_reserve(s, n));
    s->n = n;
    // Upload in KICP_UPLOAD_CHUNKS pieces on the copy stream, each followed by a 4-byte flag copy; the persistent kernel's
    // first pass waits per chunk on the flag, so the association starts while later chunks are still on the bus.
    UploadArgs up{nullptr, 0u, 1};
    const UploadArgs *upp = nullptr;
    HostUpload hu{frame_xyz, s->d_xyz, n, 1, KICP_UPLOAD_CHUNKS};
    HostUpload *hup = nullptr;
    if (n > 0) {
        const int64_t windows = (n + 31) / 32;
        const int64_t wpc = (windows + KICP_UPLOAD_CHUNKS - 1) / KICP_UPLOAD_CHUNKS;
        const uint32_t seq = ++c->upload_seq ? c->upload_seq : ++c->upload_seq;  // never 0
        for (int k = 0; k < KICP_UPLOAD_CHUNKS; ++k) c->h_chunk_tags[k] = seq;
        hu.wpc = wpc, hu.issued = 0;
        hup = &hu;
        up = UploadArgs{c->d_chunk_flags, seq, (int)wpc};
        upp = &up;
        // first chunk now; the others follow the kernel launch (launch_first) or precede it (the older order, kept for A/B)
        KICP_TRY(issue_chunks(c, hup, c->launch_first ? 1 : KICP_UPLOAD_CHUNKS));
        if (!c->overlap_upload) {
            KICP_TRY(issue_chunks(c, hup, KICP_UPLOAD_CHUNKS));
            KICP_CUDA(cudaStreamSynchronize(c->copy_stream));
            upp = nullptr;
        }
    }
    KICP_TRY(enqueue_registration(map, s, last, odom, tau, params, c->h_result, sharded, upp, hup));
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    KICP_CUDA(cudaStreamSynchronize(c->copy_stream));
    for (int k = 0; k < 7; ++k) out_pose[k] = c->h_result->pose[k];
    if (result) *result = *c->h_result;
    return c->h_result->status;
}

extern "C" int kicp_register(kicp_map *map, const double *frame_xyz, int64_t n, const double last[7], const double odom[7],
                             double tau, const kicp_reg_params *params, double out_pose[7], kicp_reg_result *result) {
    return register_host(map, frame_xyz, n, last, odom, tau, params, out_pose, result, false);
}
extern "C" int kicp_register_sharded(kicp_map *map, const double *frame_xyz, int64_t n_local, const double last[7],
                                     const double odom[7], double tau, const kicp_reg_params *params, double out_pose[7],
                                     kicp_reg_result *result) {
    return register_host(map, frame_xyz, n_local, last, odom, tau, params, out_pose, result, true);
}
