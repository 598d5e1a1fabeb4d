// kinematic_icp::KinematicRegistration::ComputeRobotMotion on the device
// (reference: cpp/kinematic_icp/registration/Registration.cpp:48-190 + kiss_icp::VoxelHashMap::GetClosestNeighbor).
//
// One registration = ONE cooperative launch of k_register<true>.  Every IRLS iteration ("pass") fuses, for every scan point,
//   q = T p                                          Registration.cpp:74
//   27-voxel probe + nearest neighbour               GetClosestNeighbor (KISS-ICP v1.2.0)
//   gate d < tau                                     Registration.cpp:75
//   r = T p - n,  J = [R e_x | R (-p_y, p_x, 0)]     Registration.cpp:86-93
//   sum of J^T J, J^T r, N, |r|^2                    Registration.cpp:95-118, 48-60
// and ends in a grid barrier after which EVERY CTA sums the per-CTA partials in the same fixed order and solves the 2x2
// system, applies the unicycle motion model and decides convergence redundantly (Registration.cpp:119-125, 159-167,
// 181-184): identical inputs give identical poses, so there is no serial section and no broadcast.  The correspondence
// list of the reference is never materialised: association and linearisation use the same T.
//
// THE SEARCH (pooled, warp-cooperative).  A warp owns a window of 32 consecutive scan points ("owners").  The work of the
// window is turned into flat streams that all 32 lanes consume together, so no lane waits for the slowest owner:
//   tasks      (owner, neighbour shift k): one hash probe each.  Batch -1 is the 32 own voxels (k = 0); its result gives
//              every owner an exact pruning bound, after which the surviving neighbour voxels of ALL owners form one
//              stream that is processed 32 tasks at a time (lane = task);
//   lines      a found voxel's points are one contiguous run of 32-byte records, 4 per 128-byte line.  The lines of a
//              batch are numbered by a warp prefix sum; a quad (4 lanes) takes one line, each lane loads ONE point with a
//              single 256-bit load (a warp instruction touches 8 full lines), KR_G line-rounds are in flight together;
//   reduction  d^2 goes to the owner's slot in shared memory with an atomicMin on the IEEE bit pattern (non-negative
//              doubles order like integers); exact ties resolve by a second atomicMin on the visiting-order key
//              (shift index, index in voxel), i.e. the first minimum in the reference's order wins.
// Exact pruning: with q in voxel v the cube of v + s is at least lb^2 = sum of the squared face gaps along the shifted axes
// away, so a voxel with lb^2 > best (1 + 1e-6) + 1e-10 cannot hold the answer (the margin covers the rounding of the face
// coordinates by 9 orders of magnitude).  Everything that could tie or win is still evaluated, so the result is the
// reference's, bit for bit in the choice of the neighbour.
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "kicp_device.cuh"

using namespace kicp_dev;

#define KR_WARPS 8                    // warps per CTA
#define KR_THREADS (KR_WARPS * 32)
#ifndef KR_MINB
#define KR_MINB 2                     // resident CTAs per SM the kernel is compiled for
#endif
#define KR_LCAP 160                   // line-map entries per chunk (32 tasks x 5 lines at 20 points per voxel)
#ifndef KR_G
#define KR_G 4                        // line-rounds (of 8 lines = 32 points) in flight together
#endif
#define KR_DBLMAX_BITS 0x7FEFFFFFFFFFFFFFull

// Pose + solver state of one registration.  The persistent kernel keeps one replica per CTA in shared memory; the
// multi-launch (NCCL) path keeps it in RegState.
struct PoseState {
    double q[4];  // current estimate: unit quaternion (x, y, z, w) ...
    double t[3];  // ... translation ...
    double R[9];  // ... and the rotation matrix of q, row-major
    double tau, conv, fixed_reg, beta;
    int adaptive, max_iter;
    int iter, done, status;
};

struct RegState {
    PoseState pose;                 // multi-launch path only
    unsigned int win_ctr;           // window tickets handed out so far (monotonic inside a registration)
    unsigned int arrive;            // grid-barrier arrivals so far (monotonic inside a registration)
    unsigned int exit_ctr;          // CTAs that have left the kernel; the last one zeroes the three counters
    unsigned int ticket;            // multi-launch path: last-CTA detection
    int abort;                      // a device-side wait gave up (status code); every CTA leaves after the current pass
    int *iters_out;                 // optional: where to publish the iteration count (profiling)
    double acc[8];                  // multi-launch path: JTJ00 JTJ01 JTJ11 JTr0 JTr1 N sum|r|^2 (unused)
    unsigned long long stats[4];    // optional work counters: probes, candidate points evaluated, lines, windows
    double dbg[KICP_MAX_ITERATIONS][4];  // per pass, ns (CTA 0): windows phase, barrier wait, partial sum (+ exchange), solve
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

// ---------------------------------------------------------------------------------------------------- arguments
struct RegArgs {
    Pose last, odom;
    double tau, conv, fixed_reg;
    int adaptive, max_iter;
    int *iters_out;
};

// The frame as it lies in HBM: float64 or float32 x,y,z fields at a byte stride (std::vector<Eigen::Vector3d> is
// {f64, 24, 0, 8, 16}; a PointCloud2 message is f32 at point_step with its field offsets, RosUtils.cpp:30-39).
struct ScanView {
    const unsigned char *base;
    int n;            // number of points (an upper bound when d_n is given)
    const int *d_n;   // optional device-resident count produced by an earlier stage on the same stream
    int stride, ox, oy, oz;
    int f32;
};

// Chunked upload overlapped with the first pass: chunk c (windows [c*windows_per_chunk, ...)) may be read once
// flags[c] == seq — the flag is copied by the same copy stream right after the chunk's data.
struct UploadArgs {
    const uint32_t *flags;  // nullptr: the scan is already resident
    uint32_t seq;
    int windows_per_chunk;  // in 32-point windows
};

struct P2PArgs {
    P2PMailbox *peer[KICP_MAX_RANKS];
    int nranks, rank, parity;
    uint32_t tag_base;  // + pass index = the tag of this registration's words
};

struct KernelArgs {
    RegState *st;
    ScanView scan;
    MapView map;
    double *partials;  // [2][grid][8]
    P2PArgs px;
    UploadArgs up;
    RegArgs init;
    int pow2_voxel;
    int collect_stats;
    unsigned long long timeout_ns;  // device-side waits (upload flags, peers) give up after this long
};

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// one stored map point {x, y, z, pad}: a single 256-bit load (LDG.E.256, sm_100)
struct __align__(32) Point4 {
    double x, y, z, w;
};
__device__ __forceinline__ Point4 ld_point(const double *p) {
    Point4 r;
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

// current_estimate = last_robot_pose * relative_wheel_odometry   (Registration.cpp:156)
__device__ void pose_init(PoseState *ps, const RegArgs &a) {
    const double lq[4] = {a.last.qx, a.last.qy, a.last.qz, a.last.qw}, lt[3] = {a.last.tx, a.last.ty, a.last.tz};
    const double oq[4] = {a.odom.qx, a.odom.qy, a.odom.qz, a.odom.qw}, ot[3] = {a.odom.tx, a.odom.ty, a.odom.tz};
    double q[4], t[3], R[9];
    se3_compose(lq, lt, oq, ot, q, t);
    quat_to_matrix(q, R);
    for (int k = 0; k < 4; ++k) ps->q[k] = q[k];
    for (int k = 0; k < 3; ++k) ps->t[k] = t[k];
    for (int k = 0; k < 9; ++k) ps->R[k] = R[k];
    ps->tau = a.tau, ps->conv = a.conv, ps->fixed_reg = a.fixed_reg, ps->beta = 0.0;
    ps->adaptive = a.adaptive, ps->max_iter = a.max_iter;
    ps->iter = 0, ps->done = a.max_iter <= 0 ? 1 : 0, ps->status = KICP_OK;
}
__device__ void result_init(kicp_reg_result *r, const PoseState *ps) {
    for (int k = 0; k < 4; ++k) r->pose[k] = ps->q[k];
    for (int k = 0; k < 3; ++k) r->pose[4 + k] = ps->t[k];
    r->beta = 0.0, r->last_dx_norm = 0.0, r->iterations = 0, r->status = ps->status;
}

// Multi-launch path and the "nothing to do" case (empty map / max_iter <= 0): state in global memory.
__global__ void k_reg_init(RegState *st, RegArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    pose_init(&st->pose, a);
    result_init(&st->result, &st->pose);
    st->ticket = 0, st->win_ctr = 0, st->arrive = 0, st->exit_ctr = 0, st->abort = 0;
    st->iters_out = a.iters_out;
    if (a.iters_out) *a.iters_out = 0;
    for (int k = 0; k < 8; ++k) st->acc[k] = 0.0;
}

// ComputePerturbation's tail + motion model + pose update + convergence test (one thread).  `s` holds the (all-reduced)
// sums of this iteration; `res` (optional) receives the public result fields.
__device__ void solve_and_update(PoseState *ps, const double *s, kicp_reg_result *res, int *iters_out) {
    const int j = ps->iter;
    if (res && j < KICP_MAX_ITERATIONS)
        for (int k = 0; k < 8; ++k) res->sums[j][k] = k < 7 ? s[k] : 0.0;
    const double N = s[5];
    if (j == 0) {
        // ComputeOdometryRegularization (Registration.cpp:48-60): beta = 1 / (mean |T0 p - n|^2 + DBL_MIN), computed
        // once from the first association; the fixed value otherwise (:171-177)
        ps->beta = ps->adaptive ? 1.0 / (s[6] / N + DBL_MIN) : ps->fixed_reg;
        if (res) res->beta = ps->beta;
    }
    // JTJ /= N; JTr /= N; JTJ += diag(beta, 0); dx = -(JTJ^-1 JTr)     (Registration.cpp:119-125)
    const double a = s[0] / N + ps->beta, b = s[1] / N, d = s[2] / N + 0.0;
    const double r0 = s[3] / N, r1 = s[4] / N;
    const double invdet = 1.0 / (a * d - b * b);
    const double i00 = d * invdet, i01 = -b * invdet, i10 = -b * invdet, i11 = a * invdet;
    const double dx0 = -(i00 * r0 + i01 * r1), dx1 = -(i10 * r0 + i11 * r1);
    // motion_model (Registration.cpp:159-167), epsilon = DBL_MIN
    double sn, cs;
    sincos(dx1, &sn, &cs);
    const double ux = dx0 * sn / (dx1 + DBL_MIN);
    const double uy = dx0 * (1.0 - cs) / (dx1 + DBL_MIN);
    double dq[4], dt[3], nq[4], nt[3], cq[4], ct[3], nR[9];
    for (int k = 0; k < 4; ++k) cq[k] = ps->q[k];
    for (int k = 0; k < 3; ++k) ct[k] = ps->t[k];
    se3_exp_planar(ux, uy, dx1, dq, dt);
    se3_compose(cq, ct, dq, dt, nq, nt);  // current_estimate = current_estimate * delta_motion  (:182)
    quat_to_matrix(nq, nR);
    for (int k = 0; k < 4; ++k) ps->q[k] = nq[k];
    for (int k = 0; k < 3; ++k) ps->t[k] = nt[k];
    for (int k = 0; k < 9; ++k) ps->R[k] = nR[k];
    const double dxn = sqrt(dx0 * dx0 + dx1 * dx1);
    ps->iter = j + 1;
    int done = (dxn < ps->conv) || (j + 1 >= ps->max_iter);  // break BEFORE re-association (:184)
    if (!(N > 0.0)) {  // the reference has no guard: the pose is NaN from here on; stop early and say so
        ps->status = KICP_WARN_NO_CORRESPONDENCES;
        done = 1;
    }
    ps->done = done;
    if (res) {
        if (j < KICP_MAX_ITERATIONS) res->dx[j][0] = dx0, res->dx[j][1] = dx1;
        res->last_dx_norm = dxn;
        res->iterations = j + 1;
        for (int k = 0; k < 4; ++k) res->pose[k] = nq[k];
        for (int k = 0; k < 3; ++k) res->pose[4 + k] = nt[k];
        res->status = ps->status;
    }
    if (iters_out) *iters_out = j + 1;
}

// Multi-launch path: the solve between the NCCL allreduce and the next association launch.
__global__ void k_solve(RegState *st) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && !st->pose.done) {
        double s[8];
        for (int k = 0; k < 8; ++k) s[k] = st->acc[k];
        solve_and_update(&st->pose, s, &st->result, st->iters_out);
        for (int k = 0; k < 8; ++k) st->acc[k] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------- per-warp shared state
struct __align__(16) WarpSm {
    double2 qxy[32];                 // owner's query point (map frame)
    double qz[32];
    unsigned long long best[32];     // bit pattern of the smallest d^2 found so far (atomicMin)
    unsigned long long kg[32];       // (visiting-order key << 32) | global point index of the FIRST such minimum (atomicMin)
    int vx[32], vy[32], vz[32];      // owner's voxel
    unsigned int tend[32];           // inclusive prefix sum of the owners' surviving-neighbour counts
    unsigned int nmask[32];          // owner's surviving neighbour shifts, bit k <-> voxel_shifts[k]
    unsigned short lmap[KR_LCAP];    // line -> (task lane, line index inside the voxel)
};

__device__ __forceinline__ void load_scan_point(const ScanView &sv, int i, double &x, double &y, double &z) {
    const unsigned char *p = sv.base + (size_t)i * (size_t)sv.stride;
    if (sv.f32) {
        x = (double)__ldg(reinterpret_cast<const float *>(p + sv.ox));
        y = (double)__ldg(reinterpret_cast<const float *>(p + sv.oy));
        z = (double)__ldg(reinterpret_cast<const float *>(p + sv.oz));
    } else {
        x = __ldg(reinterpret_cast<const double *>(p + sv.ox));
        y = __ldg(reinterpret_cast<const double *>(p + sv.oy));
        z = __ldg(reinterpret_cast<const double *>(p + sv.oz));
    }
}

// per-lane work counters (option "stats") -> warp sum -> one atomic per warp
__device__ __forceinline__ void stats_flush(RegState *st, unsigned long long probes, unsigned long long cands, unsigned long long lines) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        probes += __shfl_xor_sync(0xFFFFFFFFu, probes, d);
        cands += __shfl_xor_sync(0xFFFFFFFFu, cands, d);
        lines += __shfl_xor_sync(0xFFFFFFFFu, lines, d);
    }
    if ((threadIdx.x & 31) == 0) atomicAdd(&st->stats[0], probes), atomicAdd(&st->stats[1], cands), atomicAdd(&st->stats[2], lines);
}

// ---------------------------------------------------------------------------------------------------------------
// k_register.  PERSISTENT = true: cooperative launch, every CTA resident, all IRLS iterations inside the launch
// (single GPU, and the sharded path with the exchange over NVLink peer memory fused into the barrier).
// PERSISTENT = false: one pass per launch; the last CTA leaves the local sums in st->acc for the NCCL allreduce.
// ---------------------------------------------------------------------------------------------------------------
template <bool PERSISTENT>
__global__ void __launch_bounds__(KR_THREADS, KR_MINB) k_register(const KernelArgs a) {
    __shared__ WarpSm s_warp[KR_WARPS];
    __shared__ PoseState s_ps;
    __shared__ double s_part[KR_WARPS][8];
    __shared__ double s_sum[8];
    __shared__ MapView s_map[32];
    __shared__ int s_flag[2];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int quad = lane >> 2, sub = lane & 3;
    const unsigned FULL = 0xFFFFFFFFu;
    WarpSm &sm = s_warp[wid];
    RegState *const st = a.st;
    const MapRegs mr = map_regs(a.map, s_map);  // per-thread copy of the map view (divergence safety, kicp_device.cuh)
    const double vs = a.map.voxel_size, inv_vs = 1.0 / a.map.voxel_size;
    const int n = a.scan.d_n ? min(__ldg(a.scan.d_n), a.scan.n) : a.scan.n;
    const int num_windows = (n + 31) >> 5;
    const unsigned total_warps = gridDim.x * KR_WARPS;

    if (threadIdx.x == 0) {
        if (PERSISTENT) {
            pose_init(&s_ps, a.init);
            if (blockIdx.x == 0) {
                result_init(&st->result, &s_ps);
                if (a.init.iters_out) *a.init.iters_out = 0;
            }
        } else {
            s_ps = st->pose;
        }
    }
    __syncthreads();
    if (!PERSISTENT && s_ps.done) return;

    unsigned long long n_probe = 0, n_cand = 0, n_line = 0;

    for (unsigned it = 0; !s_ps.done; ++it) {
        double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0, cntN = 0, ssq = 0;
        const unsigned long long t_iter0 = gtime_ns();
        const unsigned ticket_base = PERSISTENT ? it * ((unsigned)num_windows + total_warps) : 0u;
        const double tau = s_ps.tau;

        // dynamic window fetch (windows differ in cost); the next ticket is requested before the current window is
        // processed so that the atomic's round trip is off the critical path
        unsigned tk = 0;
        if (lane == 0) tk = atomicAdd(&st->win_ctr, 1u);
        int w = (int)(__shfl_sync(FULL, tk, 0) - ticket_base);
        while (w < num_windows) {
            if (lane == 0) tk = atomicAdd(&st->win_ctr, 1u);
            if (PERSISTENT && a.up.flags != nullptr && it == 0u) {
                // first pass over a frame that is still being uploaded: wait until this window's chunk has landed
                if (lane == 0) {
                    const uint32_t *f = a.up.flags + min(w / a.up.windows_per_chunk, KICP_UPLOAD_CHUNKS - 1);
                    const unsigned long long deadline = gtime_ns() + a.timeout_ns;
                    const uint32_t want = a.up.seq;
                    while (ld_acquire_sys_u32(f) != want) {
                        if (gtime_ns() > deadline) {  // the copy never arrived
                            atomicExch(&st->abort, KICP_ERR_CUDA);
                            break;
                        }
                    }
                }
                __syncwarp();
            }
            // ---------------------------------------------------------------- owners: q = T p, voxel, face gaps
            const int i = w * 32 + lane;
            const bool valid = i < n;
            double px = 0, py = 0, pz = 0;
            if (valid) load_scan_point(a.scan, i, px, py, pz);
            const double qx = s_ps.R[0] * px + s_ps.R[1] * py + s_ps.R[2] * pz + s_ps.t[0];
            const double qy = s_ps.R[3] * px + s_ps.R[4] * py + s_ps.R[5] * pz + s_ps.t[1];
            const double qz = s_ps.R[6] * px + s_ps.R[7] * py + s_ps.R[8] * pz + s_ps.t[2];
            // PointToVoxel: floor(q / voxel_size); for a power-of-two voxel size the product with the (exact) reciprocal
            // is the same double as the quotient, so the cheaper form is used
            int vx, vy, vz;
            if (a.pow2_voxel) {
                vx = (int)floor(qx * inv_vs), vy = (int)floor(qy * inv_vs), vz = (int)floor(qz * inv_vs);
            } else {
                vx = voxel_coord(qx, vs), vy = voxel_coord(qy, vs), vz = voxel_coord(qz, vs);
            }
            sm.qxy[lane] = make_double2(qx, qy), sm.qz[lane] = qz;
            sm.vx[lane] = vx, sm.vy[lane] = vy, sm.vz[lane] = vz;
            sm.best[lane] = KR_DBLMAX_BITS, sm.kg[lane] = ~0ull;
            unsigned long long seen = KR_DBLMAX_BITS;  // sm.best[lane] as this lane last saw it
            unsigned total = 0;                        // neighbour tasks of the window
            int nbatch = 0;
            __syncwarp();

            for (int b = -1; b < nbatch; ++b) {
                // ------------------------------------------------------------ this lane's task: (owner o, shift k)
                int o = lane, k = 0;
                bool act = valid;
                if (b >= 0) {
                    const unsigned s = (unsigned)b * 32u + (unsigned)lane;
                    act = s < total;
                    int lo = 0;  // owner = number of owners whose tasks end at or before s
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1)
                        if (sm.tend[lo + step - 1] <= s) lo += step;
                    o = lo;
                    const unsigned before = o ? sm.tend[o - 1] : 0u;
                    const unsigned mk = sm.nmask[o];
                    k = act ? (int)__fns(mk, 0u, (int)(s - before) + 1) : 0;
                }
                // ------------------------------------------------------------ hash probe (warp-uniform loop)
                uint32_t meta = KICP_SLOT_EMPTY;
                {
                    int kx = 0, ky = 0, kz = 0;
                    uint32_t h = 0;
                    if (act) {
                        kx = sm.vx[o] + shift_x(k), ky = sm.vy[o] + shift_y(k), kz = sm.vz[o] + shift_z(k);
                        h = voxel_hash(kx, ky, kz) & mr.mask;
                    }
                    bool pend = act;
                    while (__any_sync(FULL, pend)) {
                        if (pend) {
                            const int4 sl = __ldg(&mr.slots[h]);
                            if ((uint32_t)sl.w == KICP_SLOT_EMPTY) {
                                pend = false;
                            } else if (sl.x == kx && sl.y == ky && sl.z == kz) {
                                meta = (uint32_t)sl.w, pend = false;
                            } else {
                                h = (h + 1) & mr.mask;
                            }
                        }
                        __syncwarp();
                    }
                }
                // ------------------------------------------------------------ number the 128-byte lines of the batch
                const int cnt = meta == KICP_SLOT_EMPTY ? 0 : (int)(meta & 0xFFu);
                const int nl = (cnt + 3) >> 2;
                int incl = nl;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int y = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += y;
                }
                const int loff = incl - nl;
                const int ltot = __shfl_sync(FULL, incl, 31);
                const int maxnl = __reduce_max_sync(FULL, nl);
                const unsigned okpack = ((unsigned)o << 5) | (unsigned)k;
                if (a.collect_stats) n_probe += act ? 1 : 0, n_cand += cnt, n_line += nl;

                for (int lbase = 0; lbase < ltot; lbase += KR_LCAP) {
                    // line map of this chunk: entry = task lane | (line inside the voxel << 5)
                    for (int li = 0; li < maxnl; ++li) {
                        const int pos = loff + li - lbase;
                        if (li < nl && pos >= 0 && pos < KR_LCAP) sm.lmap[pos] = (unsigned short)(lane | (li << 5));
                    }
                    __syncwarp();
                    const int nr = min(KR_LCAP, ltot - lbase);
                    for (int r0 = 0; r0 < nr; r0 += 8 * KR_G) {
                        double d2[KR_G];
                        unsigned long long kgv[KR_G];
                        int own[KR_G];
                        unsigned hasm = 0;
#pragma unroll
                        for (int g = 0; g < KR_G; ++g) {
                            const int line = r0 + g * 8 + quad;
                            const bool lv = line < nr;
                            const unsigned e = lv ? (unsigned)sm.lmap[line] : 0u;
                            const int t = (int)(e & 31u), li = (int)(e >> 5);
                            const uint32_t m = __shfl_sync(FULL, meta, t);
                            const unsigned ok = __shfl_sync(FULL, okpack, t);
                            const int j = li * 4 + sub;
                            const bool has = lv && j < (int)(m & 0xFFu);
                            const unsigned gidx = (m >> 8) * (unsigned)mr.cap + (unsigned)j;
                            const int oo = (int)(ok >> 5);
                            own[g] = oo;
                            kgv[g] = ((unsigned long long)(((ok & 31u) << 8) | (unsigned)j) << 32) | (unsigned long long)gidx;
                            d2[g] = DBL_MAX;
                            if (has) {
                                const Point4 c = ld_point(mr.pts + (size_t)gidx * KICP_PSTRIDE);
                                const double2 qq = sm.qxy[oo];
                                const double dx = c.x - qq.x, dy = c.y - qq.y, dz = c.z - sm.qz[oo];
                                d2[g] = dx * dx + dy * dy + dz * dz;
                                hasm |= 1u << g;
                            }
                        }
                        // phase 1: lower the owners' best; remember who could still be (or tie with) the minimum
                        unsigned lem = 0;
#pragma unroll
                        for (int g = 0; g < KR_G; ++g) {
                            if (hasm & (1u << g)) {
                                const unsigned long long mine = (unsigned long long)__double_as_longlong(d2[g]);
                                const unsigned long long br = sm.best[own[g]];
                                if (mine <= br) {
                                    lem |= 1u << g;
                                    if (mine < br) atomicMin(&sm.best[own[g]], mine);
                                }
                            }
                        }
                        if (__any_sync(FULL, lem != 0u)) {
                            __syncwarp();
                            // every lane, as an owner: a strictly smaller best invalidates the recorded first minimum
                            const unsigned long long cur = sm.best[lane];
                            if (cur != seen) sm.kg[lane] = ~0ull, seen = cur;
                            __syncwarp();
                            // phase 2: among the candidates AT the minimum, the first in visiting order wins
#pragma unroll
                            for (int g = 0; g < KR_G; ++g) {
                                if (lem & (1u << g)) {
                                    if ((unsigned long long)__double_as_longlong(d2[g]) == sm.best[own[g]])
                                        atomicMin(&sm.kg[own[g]], kgv[g]);
                                }
                            }
                        }
                        __syncwarp();
                    }
                    __syncwarp();
                }

                if (b < 0) {
                    // ------------------------------------------------------------ exact pruning of the 26 neighbours
                    __syncwarp();
                    const double best = __longlong_as_double((long long)sm.best[lane]);
                    const double bound = best * (1.0 + 1e-6) + 1e-10;
                    double t;
                    t = (double)(vx + 1) * vs - qx; const double gxp = t * t;
                    t = qx - (double)vx * vs;       const double gxm = t * t;
                    t = (double)(vy + 1) * vs - qy; const double gyp = t * t;
                    t = qy - (double)vy * vs;       const double gym = t * t;
                    t = (double)(vz + 1) * vs - qz; const double gzp = t * t;
                    t = qz - (double)vz * vs;       const double gzm = t * t;
                    unsigned mask = valid ? 0x07FFFFFEu : 0u;
                    mask &= (kX0 | (gxp <= bound ? kXP : 0u) | (gxm <= bound ? kXM : 0u)) &
                            (kY0 | (gyp <= bound ? kYP : 0u) | (gym <= bound ? kYM : 0u)) &
                            (kZ0 | (gzp <= bound ? kZP : 0u) | (gzm <= bound ? kZM : 0u));
                    // edges and corners: the summed gap decides
#pragma unroll
                    for (int kk = 7; kk < 27; ++kk) {
                        const double lb2 = (shift_x(kk) > 0 ? gxp : (shift_x(kk) < 0 ? gxm : 0.0)) +
                                           (shift_y(kk) > 0 ? gyp : (shift_y(kk) < 0 ? gym : 0.0)) +
                                           (shift_z(kk) > 0 ? gzp : (shift_z(kk) < 0 ? gzm : 0.0));
                        if (lb2 > bound) mask &= ~(1u << kk);
                    }
                    const int no = __popc(mask);
                    int tin = no;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int y = __shfl_up_sync(FULL, tin, d);
                        if (lane >= d) tin += y;
                    }
                    sm.tend[lane] = (unsigned)tin, sm.nmask[lane] = mask;
                    total = (unsigned)__shfl_sync(FULL, tin, 31);
                    nbatch = (int)((total + 31u) >> 5);
                    __syncwarp();
                }
            }
            __syncwarp();
            // ---------------------------------------------------------------- gate, residual, Jacobian, sums
            const unsigned long long bb = sm.best[lane];
            const unsigned gwin = (unsigned)(sm.kg[lane] & 0xFFFFFFFFull);
            if (valid && bb != KR_DBLMAX_BITS) {
                const Point4 c = ld_point(mr.pts + (size_t)gwin * KICP_PSTRIDE);
                const double rx = qx - c.x, ry = qy - c.y, rz = qz - c.z;  // r = T p - n
                const double rr = rx * rx + ry * ry + rz * rz;
                if (sqrt(rr) < tau) {  // distance < max_correspondance_distance   (Registration.cpp:75)
                    // J = [R e_x | R (-p_y, p_x, 0)]      (Registration.cpp:89-91)
                    const double c0x = s_ps.R[0], c0y = s_ps.R[3], c0z = s_ps.R[6];
                    const double c1x = s_ps.R[1] * px - s_ps.R[0] * py, c1y = s_ps.R[4] * px - s_ps.R[3] * py,
                                 c1z = s_ps.R[7] * px - s_ps.R[6] * py;
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
            w = (int)(__shfl_sync(FULL, tk, 0) - ticket_base);
        }

        const unsigned long long t_win = gtime_ns();
        // ---------------------------------------------------------------- warp -> CTA partial (plain stores, fixed order)
        double v[7] = {a00, a01, a11, b0, b1, cntN, ssq};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], d);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 7; ++k) s_part[wid][k] = v[k];
            s_part[wid][7] = 0.0;
        }
        __syncthreads();
        double *const part = a.partials + (size_t)(PERSISTENT ? (it & 1u) : 0u) * gridDim.x * 8;
        if (threadIdx.x < 8) {
            double s = 0.0;
            for (int k = 0; k < KR_WARPS; ++k) s += s_part[k][threadIdx.x];
            __stcg(&part[(size_t)blockIdx.x * 8 + threadIdx.x], s);
            __threadfence();
        }
        __syncthreads();

        if (!PERSISTENT) {
            // one pass per launch: the last CTA sums the partials into st->acc (the NCCL allreduce and k_solve follow)
            if (threadIdx.x == 0) {
                const unsigned ticket = atomicAdd(&st->ticket, 1u);
                s_flag[0] = (ticket == gridDim.x - 1);
            }
            __syncthreads();
            if (s_flag[0]) {
                __threadfence();
                if (threadIdx.x < 8) {
                    double s = 0.0;
                    for (unsigned bb = 0; bb < gridDim.x; ++bb) s += __ldcg(&part[(size_t)bb * 8 + threadIdx.x]);
                    st->acc[threadIdx.x] = s;
                }
                if (threadIdx.x == 0) st->ticket = 0, st->win_ctr = 0;
            }
            if (a.collect_stats) stats_flush(st, n_probe, n_cand, n_line);
            return;
        }

        // ---------------------------------------------------------------- grid barrier: arrive, then everyone reduces
        const bool multi = a.px.nranks > 1;
        unsigned long long t_arr = 0, t_red = 0;
        if (threadIdx.x == 0) {
            atomicAdd(&st->arrive, 1u);
            if (!multi || blockIdx.x == 0) {
                const unsigned target = (it + 1u) * gridDim.x;
                const unsigned long long deadline = t_win + a.timeout_ns;  // per-thread register: no uniform read inside the spin
                while (ld_acquire_gpu_u32(&st->arrive) < target) {
                    if (gtime_ns() > deadline) {  // a CTA of this grid never arrived: give up instead of hanging
                        atomicExch(&st->abort, KICP_ERR_CUDA);
                        break;
                    }
                }
            }
            t_arr = gtime_ns();
        }
        __syncthreads();
        if (!multi || blockIdx.x == 0) {
            // column k = thread & 7, rows strided by 32: fixed summation tree, identical in every CTA
            const int col = threadIdx.x & 7, row0 = threadIdx.x >> 3;
            double s = 0.0;
            const unsigned nrow = (gridDim.x + KR_THREADS / 8 - 1) / (KR_THREADS / 8);  // uniform trip count, tail predicated
            for (unsigned kr = 0; kr < nrow; ++kr) {
                const unsigned bb = kr * (KR_THREADS / 8) + (unsigned)row0;
                if (bb < gridDim.x) s += __ldcg(&part[(size_t)bb * 8 + col]);
            }
            s += __shfl_xor_sync(FULL, s, 8);
            s += __shfl_xor_sync(FULL, s, 16);
            if (lane < 8) s_part[wid][lane] = s;
            __syncthreads();
            if (threadIdx.x < 8) {
                double tsum = 0.0;
                for (int k = 0; k < KR_WARPS; ++k) tsum += s_part[k][threadIdx.x];
                s_sum[threadIdx.x] = tsum;
            }
            __syncthreads();
        }
        if (multi) {
            // Exchange fused into the barrier (NCCL-LL style): CTA 0 writes its 8 local sums as sixteen 8-byte words
            // {32 data bits, 32-bit tag} into EVERY rank's mailbox over NVLink — an aligned 8-byte store arrives whole, so no
            // fence and no separate flag are needed; every CTA of every rank polls its OWN GPU's mailbox until all ranks'
            // words carry the tag and adds them IN RANK ORDER (same values, same order -> the same pose on every rank).
            const uint32_t tag = a.px.tag_base + it;
            if (wid == 0) {
                if (blockIdx.x == 0) {
                    const double val = s_sum[(lane & 15) >> 1];
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(val);
                    const uint32_t half = (lane & 1) ? (uint32_t)(bits >> 32) : (uint32_t)bits;
                    const unsigned long long word = ((unsigned long long)tag << 32) | half;
                    // lanes 16..31 repeat the stores of lanes 0..15 (same address, same value): no divergent section
                    for (int r = 0; r < a.px.nranks; ++r) {
                        st_relaxed_sys_u64(&a.px.peer[r]->ll[a.px.parity][it][a.px.rank][lane & 15], word);
                        __syncwarp();
                    }
                }
                __syncwarp();
                double tot = 0.0;
                bool timed_out = false;
                const unsigned long long deadline = gtime_ns() + a.timeout_ns;
                for (int r = 0; r < a.px.nranks; ++r) {
                    unsigned long long wv = 0;
                    const unsigned long long *src = &a.px.peer[a.px.rank]->ll[a.px.parity][it][r][lane & 15];
                    bool pend = true;
                    while (__any_sync(FULL, pend)) {  // warp-uniform loop
                        if (pend) {
                            wv = ld_relaxed_sys_u64(src);
                            if ((uint32_t)(wv >> 32) == tag) {
                                pend = false;
                            } else if (gtime_ns() > deadline) {  // a peer is missing
                                timed_out = true, pend = false;
                            }
                        }
                    }
                    const uint32_t lo = __shfl_sync(FULL, (uint32_t)wv, (lane & 7) * 2);
                    const uint32_t hi = __shfl_sync(FULL, (uint32_t)wv, (lane & 7) * 2 + 1);
                    tot += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
                }
                timed_out = __any_sync(FULL, timed_out);
                if (lane < 8) s_sum[lane] = tot;
                if (lane == 0) s_flag[1] = timed_out ? 1 : 0;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            t_red = gtime_ns();
            // a wait that gave up (upload flag on this GPU, a peer's words): keep the last pose, report, leave
            int ab = (int)ld_acquire_gpu_u32((const unsigned int *)&st->abort);
            if (!ab && multi && s_flag[1]) ab = KICP_ERR_NCCL;
            if (ab) {
                s_ps.status = ab, s_ps.done = 1;
                if (blockIdx.x == 0) st->result.status = ab;
            } else {
                double s[8];
                for (int k = 0; k < 8; ++k) s[k] = s_sum[k];
                solve_and_update(&s_ps, s, blockIdx.x == 0 ? &st->result : nullptr, blockIdx.x == 0 ? a.init.iters_out : nullptr);
            }
            if (blockIdx.x == 0 && it < KICP_MAX_ITERATIONS) {
                st->dbg[it][0] = (double)(t_win - t_iter0), st->dbg[it][1] = (double)(t_arr - t_win);
                st->dbg[it][2] = (double)(t_red - t_arr), st->dbg[it][3] = (double)(gtime_ns() - t_red);
            }
        }
        __syncthreads();
    }

    if (PERSISTENT) {
        if (a.collect_stats) stats_flush(st, n_probe, n_cand, n_line);
        // the last CTA to leave zeroes the counters for the next registration on this stream
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned left = atomicAdd(&st->exit_ctr, 1u);
            if (left == gridDim.x - 1) {
                st->win_ctr = 0, st->arrive = 0, st->abort = 0;
                __threadfence();
                st->exit_ctr = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------- host
// Per-context options (kicp_ctx_set_option): "persistent" 1 = one cooperative launch per registration (default), 0 = one
// launch per IRLS iteration; "stats" 1 = count probes / candidate points / lines on the device (kicp_debug_last_stats);
// "ctas_per_sm" caps the resident CTAs per SM the grid is sized for (0 = occupancy limit); "spin_timeout_ms" bounds every
// device-side wait (upload flags, peers of the fused exchange).
extern "C" int kicp_ctx_set_option(kicp_ctx *c, const char *name, int32_t value) {
    if (!c || !name) return KICP_ERR_INVALID;
    if (!strcmp(name, "persistent")) {
        if (value != 0 && value != 1) return KICP_ERR_INVALID;
        c->persistent = value;
    } else if (!strcmp(name, "stats")) {
        if (value != 0 && value != 1) return KICP_ERR_INVALID;
        c->collect_stats = value;
    } else if (!strcmp(name, "ctas_per_sm")) {
        if (value < 0 || value > 16) return KICP_ERR_INVALID;
        c->ctas_per_sm_cap = value;
    } else if (!strcmp(name, "spin_timeout_ms")) {
        if (value < 1) return KICP_ERR_INVALID;
        c->spin_timeout_ms = value;
    } else if (!strcmp(name, "overlap_upload")) {
        if (value != 0 && value != 1) return KICP_ERR_INVALID;
        c->overlap_upload = value;
    } else {
        kicp_set_error(std::string("kicp_ctx_set_option: unknown option ") + name);
        return KICP_ERR_INVALID;
    }
    return KICP_OK;
}

static int reg_reserve(kicp_ctx *c) {
    if (c->d_state) return KICP_OK;
    KICP_CUDA(cudaMalloc(&c->d_state, sizeof(RegState)));
    KICP_CUDA(cudaMemset(c->d_state, 0, sizeof(RegState)));
    int per_sm = 0;
    KICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_register<true>, KR_THREADS, 0));
    c->persistent_ctas_per_sm = std::max(per_sm, 1);
    KICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_register<false>, KR_THREADS, 0));
    c->pruned_ctas_per_sm = std::max(per_sm, 1);
    const int max_grid = c->sm_count * std::max(c->persistent_ctas_per_sm, c->pruned_ctas_per_sm);
    KICP_CUDA(cudaMalloc(&c->d_partials, (size_t)2 * max_grid * 8 * sizeof(double)));
    int coop = 0;
    KICP_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
    if (!coop) c->persistent = 0;
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

// Host side of the chunked upload: the frame's bytes go out in KICP_UPLOAD_CHUNKS pieces on the copy stream, each followed
// by a 4-byte flag copy; the persistent kernel's first pass waits per chunk on the flag, so the association starts while
// later chunks are still on the bus.
struct HostUpload {
    const unsigned char *src;
    unsigned char *dst;
    int64_t n, wpc;   // points, windows per chunk
    int64_t stride;   // bytes per point
};
static int issue_chunks(kicp_ctx *c, const HostUpload &hu) {
    for (int k = 0; k < KICP_UPLOAD_CHUNKS; ++k) {
        const int64_t lo = std::min<int64_t>(hu.n, k * hu.wpc * 32), hi = std::min<int64_t>(hu.n, (k + 1) * hu.wpc * 32);
        if (hi > lo)
            KICP_CUDA(cudaMemcpyAsync(hu.dst + lo * hu.stride, hu.src + lo * hu.stride, (size_t)((hi - lo) * hu.stride),
                                      cudaMemcpyHostToDevice, c->copy_stream));
        KICP_CUDA(cudaMemcpyAsync(c->d_chunk_flags + k, c->h_chunk_tags + k, sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream));
    }
    return KICP_OK;
}

static ScanView scan_view(const kicp_scan *s) {
    ScanView v;
    v.base = (const unsigned char *)s->d_data, v.n = (int)s->n, v.d_n = s->d_n;
    v.stride = s->stride, v.ox = s->ox, v.oy = s->oy, v.oz = s->oz, v.f32 = s->dtype == KICP_DTYPE_F32 ? 1 : 0;
    return v;
}

// Enqueue one full registration on the context stream.  `sharded`: this rank holds a contiguous index range of the frame;
// the 8 sums of every iteration are exchanged (fused peer-memory exchange when kicp_comm_p2p_init was called, else NCCL).
// Everything that can fail on arguments is checked before any device work is issued.
static int enqueue_registration(kicp_map *m, const kicp_scan *scan, const double last[7], const double odom[7], double tau,
                                const kicp_reg_params *p, kicp_reg_result *result, bool sharded, const HostUpload *host_upload = nullptr) {
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
    KICP_TRY(reg_reserve(c));
    KernelArgs ka{};
    ka.st = c->d_state;
    ka.scan = scan_view(scan);
    ka.map = m->view();
    ka.partials = c->d_partials;
    ka.px.nranks = 1;
    ka.up = UploadArgs{nullptr, 0u, 1};
    ka.init.last = Pose{last[0], last[1], last[2], last[3], last[4], last[5], last[6]};
    ka.init.odom = Pose{odom[0], odom[1], odom[2], odom[3], odom[4], odom[5], odom[6]};
    ka.init.tau = tau, ka.init.conv = p->convergence_criterion, ka.init.fixed_reg = p->fixed_regularization;
    ka.init.adaptive = p->use_adaptive_odometry_regularization ? 1 : 0;
    // an empty map returns the prediction (Registration.cpp:157): no association, no solve
    ka.init.max_iter = m->num_blocks == 0 ? 0 : p->max_num_iterations;
    ka.init.iters_out = nullptr;
    {
        int e = 0;
        ka.pow2_voxel = std::frexp(m->voxel_size, &e) == 0.5 ? 1 : 0;
    }
    ka.collect_stats = c->collect_stats;
    ka.timeout_ns = (unsigned long long)c->spin_timeout_ms * 1000000ull;
    const int n = (int)scan->n;
    const bool p2p = sharded && c->p2p_ready;
    const bool persistent = c->persistent && (!sharded || p2p);

    kicp_ctx::ProfReg *pr = nullptr;
    if (c->profiling && (int64_t)c->prof.size() < c->prof_cap) {
        c->prof.emplace_back();
        pr = &c->prof.back();
        pr->d_iters = c->d_prof_iters + (c->prof.size() - 1);
        ka.init.iters_out = pr->d_iters;
        KICP_CUDA(cudaEventCreate(&pr->prep0));
        KICP_CUDA(cudaEventCreate(&pr->prep1));
        KICP_CUDA(cudaEventRecord(pr->prep0, c->stream));
    }
    if (c->collect_stats)
        KICP_CUDA(cudaMemsetAsync((char *)c->d_state + offsetof(RegState, stats), 0, sizeof(((RegState *)0)->stats), c->stream));

    const bool dbg = getenv("KICP_DEBUG_SYNC") != nullptr;
    if (ka.init.max_iter <= 0 || !persistent) {
        k_reg_init<<<1, 32, 0, c->stream>>>(c->d_state, ka.init);
        KICP_CHECK_LAUNCH(c);
    }
    if (pr) KICP_CUDA(cudaEventRecord(pr->prep1, c->stream));
    if (ka.init.max_iter > 0) {
        // every CTA is resident and pulls 32-point windows from a device-side counter; a small scan is spread one window per
        // CTA over the whole machine (a window is a chain of dependent memory round trips: latency, not throughput)
        const int num_windows = (n + 31) / 32;
        int per_sm = persistent ? c->persistent_ctas_per_sm : c->pruned_ctas_per_sm;
        if (c->ctas_per_sm_cap > 0) per_sm = std::min(per_sm, c->ctas_per_sm_cap);
        const int grid = std::max(1, std::min(num_windows, c->sm_count * per_sm));
        if (persistent) {
            if (host_upload && c->overlap_upload) {
                const uint32_t seq = ++c->upload_seq ? c->upload_seq : ++c->upload_seq;  // never 0
                for (int k = 0; k < KICP_UPLOAD_CHUNKS; ++k) c->h_chunk_tags[k] = seq;
                ka.up = UploadArgs{c->d_chunk_flags, seq, (int)host_upload->wpc};
                KICP_TRY(issue_chunks(c, *host_upload));
            }
            if (p2p) {
                for (int r = 0; r < c->nranks; ++r) ka.px.peer[r] = c->p2p_peer[r];
                ka.px.nranks = c->nranks, ka.px.rank = c->rank;
                ka.px.parity = (int)(c->p2p_seq & 1ull);
                ka.px.tag_base = (uint32_t)((c->p2p_seq * KICP_MAX_ITERATIONS + 1ull) & 0xFFFFFFFFull);
                if (ka.px.tag_base > 0xFFFFFF00u) ka.px.tag_base = 1u, c->p2p_seq = 0;  // wrap (tags stay non-zero)
                c->p2p_seq++;
            }
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (pr) {
                KICP_CUDA(cudaEventCreate(&e0));
                KICP_CUDA(cudaEventCreate(&e1));
                pr->it.push_back(e0), pr->it.push_back(e1);
                pr->persistent = true;
                KICP_CUDA(cudaEventRecord(e0, c->stream));
            }
            void *args[] = {&ka};
            KICP_CUDA(cudaLaunchCooperativeKernel((const void *)k_register<true>, dim3(grid), dim3(KR_THREADS), args, 0, c->stream));
            c->launches++;
            if (pr) KICP_CUDA(cudaEventRecord(e1, c->stream));
            if (dbg) {
                cudaError_t e = cudaStreamSynchronize(c->stream);
                fprintf(stderr, "[kicp] persistent launch (grid %d, n %d): %s\n", grid, n, cudaGetErrorString(e));
            }
        } else {
            for (int j = 0; j < ka.init.max_iter; ++j) {
                cudaEvent_t e0 = nullptr, e1 = nullptr;
                if (pr) {
                    KICP_CUDA(cudaEventCreate(&e0));
                    KICP_CUDA(cudaEventCreate(&e1));
                    pr->it.push_back(e0), pr->it.push_back(e1);
                    KICP_CUDA(cudaEventRecord(e0, c->stream));
                }
                k_register<false><<<grid, KR_THREADS, 0, c->stream>>>(ka);
                KICP_CHECK_LAUNCH(c);
                if (pr) KICP_CUDA(cudaEventRecord(e1, c->stream));
                if (sharded) KICP_TRY(kicp_comm_allreduce8(c, c->d_state->acc));
                k_solve<<<1, 32, 0, c->stream>>>(c->d_state);
                KICP_CHECK_LAUNCH(c);
                if (dbg) {
                    cudaError_t e = cudaStreamSynchronize(c->stream);
                    fprintf(stderr, "[kicp] pass %d (grid %d, n %d): %s\n", j, grid, n, cudaGetErrorString(e));
                }
            }
        }
    }
    if (result)
        KICP_CUDA(cudaMemcpyAsync(result, &c->d_state->result, sizeof(kicp_reg_result), cudaMemcpyDeviceToHost, c->stream));
    return KICP_OK;
}

// debugging aids (not part of the public header): per-pass device timings and work counters of the last registration
extern "C" int kicp_debug_last_timing(kicp_ctx *c, double *out /* [KICP_MAX_ITERATIONS][4] */) {
    if (!c || !c->d_state || !out) return KICP_ERR_INVALID;
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    KICP_CUDA(cudaMemcpy(out, (const char *)c->d_state + offsetof(RegState, dbg), sizeof(double) * KICP_MAX_ITERATIONS * 4,
                         cudaMemcpyDeviceToHost));
    return KICP_OK;
}
extern "C" int kicp_debug_last_stats(kicp_ctx *c, uint64_t out[4] /* probes, candidate points, 128-byte lines, 0 */) {
    if (!c || !c->d_state || !out) return KICP_ERR_INVALID;
    KICP_CUDA(cudaStreamSynchronize(c->stream));
    KICP_CUDA(cudaMemcpy(out, (const char *)c->d_state + offsetof(RegState, stats), sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost));
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

int kicp_enqueue_registration_device(kicp_map *m, const double *d_xyz, int64_t n_max, const int *d_n, const double last[7],
                                     const double odom[7], double tau, const kicp_reg_params *p) {
    if (!m || n_max < 0 || (n_max > 0 && !d_xyz)) return KICP_ERR_INVALID;
    kicp_scan view;  // non-owning alias of the caller's device buffer (packed xyz doubles)
    view.ctx = m->ctx, view.d_data = const_cast<double *>(d_xyz), view.cap_bytes = n_max * 24, view.n = n_max, view.d_n = d_n;
    const int st = enqueue_registration(m, &view, last, odom, tau, p, m->ctx->h_result, false);
    view.d_data = nullptr;  // not ours
    return st;
}

// Host-pointer entry points: validate, upload (chunked, overlapped with the first pass), register, read the result back.
static int register_host(kicp_map *map, const void *data, int64_t n, int32_t dtype, int32_t point_step, int32_t ox, int32_t oy,
                         int32_t oz, const double last[7], const double odom[7], double tau, const kicp_reg_params *params,
                         double out_pose[7], kicp_reg_result *result, bool sharded) {
    if (!map || n < 0 || (n > 0 && !data) || !out_pose || !last || !odom) return KICP_ERR_INVALID;
    KICP_TRY(check_params(params));
    kicp_ctx *c = map->ctx;
    if (sharded && !c->nccl_comm && !c->p2p_ready) {
        kicp_set_error("kicp_register_sharded: neither kicp_comm_p2p_init nor kicp_comm_init has been called on this context");
        return KICP_ERR_INVALID;
    }
    KICP_CUDA(cudaSetDevice(c->device));
    if (!c->upload_scan) KICP_TRY(kicp_scan_create(c, 0, &c->upload_scan));
    kicp_scan *s = c->upload_scan;
    KICP_TRY(kicp_scan_set_layout(s, dtype, point_step, ox, oy, oz));
    KICP_TRY(kicp_scan_reserve_bytes(s, n * (int64_t)s->stride));
    s->n = n, s->d_n = nullptr;
    const bool overlap = n > 0 && c->overlap_upload && c->persistent && (!sharded || c->p2p_ready) && map->num_blocks != 0 &&
                         params->max_num_iterations > 0;
    HostUpload hu{(const unsigned char *)data, (unsigned char *)s->d_data, n, 1, s->stride};
    if (n > 0) {
        const int64_t windows = (n + 31) / 32;
        hu.wpc = (windows + KICP_UPLOAD_CHUNKS - 1) / KICP_UPLOAD_CHUNKS;
        if (!overlap) {
            KICP_CUDA(cudaMemcpyAsync(s->d_data, data, (size_t)(n * s->stride), cudaMemcpyHostToDevice, c->stream));
        }
    }
    int st = enqueue_registration(map, s, last, odom, tau, params, c->h_result, sharded, overlap ? &hu : nullptr);
    cudaError_t e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->copy_stream);  // drain on every path
    if (st != KICP_OK) return st;
    if (e1 != cudaSuccess) return kicp_cuda_fail(e1, "cudaStreamSynchronize(stream)", __FILE__, __LINE__);
    if (e2 != cudaSuccess) return kicp_cuda_fail(e2, "cudaStreamSynchronize(copy_stream)", __FILE__, __LINE__);
    for (int k = 0; k < 7; ++k) out_pose[k] = c->h_result->pose[k];
    if (result) *result = *c->h_result;
    if (c->h_result->status == KICP_ERR_CUDA || c->h_result->status == KICP_ERR_NCCL)
        kicp_set_error("a device-side wait of the registration kernel timed out (upload flag or a peer of the fused exchange)");
    return c->h_result->status;
}

extern "C" int kicp_register(kicp_map *map, const double *frame_xyz, int64_t n, const double last[7], const double odom[7],
                             double tau, const kicp_reg_params *params, double out_pose[7], kicp_reg_result *result) {
    return register_host(map, frame_xyz, n, KICP_DTYPE_F64, 0, 0, 0, 0, last, odom, tau, params, out_pose, result, false);
}
extern "C" int kicp_register_points(kicp_map *map, const void *data, int64_t n, int32_t dtype, int32_t point_step, int32_t offset_x,
                                    int32_t offset_y, int32_t offset_z, const double last[7], const double odom[7], double tau,
                                    const kicp_reg_params *params, double out_pose[7], kicp_reg_result *result) {
    return register_host(map, data, n, dtype, point_step, offset_x, offset_y, offset_z, last, odom, tau, params, out_pose, result, false);
}
extern "C" int kicp_register_sharded(kicp_map *map, const double *frame_xyz, int64_t n_local, const double last[7],
                                     const double odom[7], double tau, const kicp_reg_params *params, double out_pose[7],
                                     kicp_reg_result *result) {
    return register_host(map, frame_xyz, n_local, KICP_DTYPE_F64, 0, 0, 0, 0, last, odom, tau, params, out_pose, result, true);
}
extern "C" int kicp_register_points_sharded(kicp_map *map, const void *data, int64_t n_local, int32_t dtype, int32_t point_step,
                                            int32_t offset_x, int32_t offset_y, int32_t offset_z, const double last[7],
                                            const double odom[7], double tau, const kicp_reg_params *params, double out_pose[7],
                                            kicp_reg_result *result) {
    return register_host(map, data, n_local, dtype, point_step, offset_x, offset_y, offset_z, last, odom, tau, params, out_pose, result,
                         true);
}
