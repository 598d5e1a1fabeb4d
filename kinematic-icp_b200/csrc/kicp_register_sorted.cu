// kinematic_icp::KinematicRegistration::ComputeRobotMotion on the device, second engine: VOXEL-SORTED LANES
// (reference: cpp/kinematic_icp/registration/Registration.cpp:48-190 + kiss_icp::VoxelHashMap::GetClosestNeighbor).
//
// The pooled engine of kicp_register.cu fights the divergence of a thread-per-point search — the 32 points of a warp, taken
// in scan order, sit in ~20 different voxels — by flattening a window's work into streams.  This engine removes the cause
// instead: ONCE per registration the frame is grouped by the voxel of q = T0 p (a counting sort on a scratch hash table, three
// light grid-wide sweeps), after which the 32 lanes of a warp hold points of the same few voxels (4.9 distinct voxels per warp
// at BASELINE config 4, 19.7 in scan order).  Every lane then runs the plain exact search for its own point — own voxel, 6
// faces, 12 edges + 8 corners, each stage pruned with the best distance so far — and its neighbours in the warp ask for the
// same hash slots and the same candidate points at about the same time: one L1 transaction serves them all, the loops have
// similar trip counts, and no work buffers, prefix sums or shared-memory hand-offs are needed.
//
// The order is only a matter of locality: every point is searched around ITS OWN current voxel in every pass, so the result
// does not depend on the grouping (the pose moves by centimetres between passes; the order of pass 0 stays coherent).
//
//   sort (once)   S1  q = T0 p, voxel -> scratch table slot (CAS on the key), rank = atomicAdd(count)
//                 S2  every CTA turns the counts of its slice of the table into offsets (block scan + one atomic per CTA)
//                 S3  sorted[offset + rank] = p
//   pass          lane t: p = sorted[t], q = T p, search (seeded with the distance to the previous pass's neighbour when that
//                 point is still one of the 27 voxels' points), gate, residual, Jacobian, seven sums   Registration.cpp:62-118
//                 grid barrier, every CTA: fixed-order sum of the per-CTA partials, solve, motion model  Registration.cpp:119-125,159-187
//
// Exactness: candidates are visited in the reference's order (shift table order, insertion order inside a voxel) with the
// reference's comparison (strict < on norms, first minimum wins); a voxel is skipped only when its cube is provably farther than
// min(best so far, tau, seed) — see kicp_register.cu for the argument, which is identical.
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "kicp_register_sorted.cuh"
namespace {  // this translation unit's own copy of the pose / solve helpers (kicp_register.cu has the other)
#include "kicp_solve.cuh"
}

using namespace kicp_dev;

#ifndef KS_THREADS
#define KS_THREADS 256                // threads per CTA
#endif
#ifndef KS_MINB
#define KS_MINB 3                     // resident CTAs per SM the kernel is compiled for
#endif
#define KS_WARPS (KS_THREADS / 32)
#ifndef KS_COUNT
#define KS_COUNT(counter, value)      // tests/emu: warp-level work model (shift-loop steps, candidate-loop steps); nothing on the device
#endif
#define KS_NONE 0xFFFFFFFFu
#define KS_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

struct SortedState {
    unsigned int win_ctr;   // chunk tickets handed out so far (monotonic inside a registration)
    unsigned int arrive;    // grid-barrier arrivals so far (monotonic inside a registration)
    unsigned int exit_ctr;  // CTAs that have left the kernel; the last one zeroes the counters
    unsigned int cursor;    // sorted positions handed to table slices so far
    int abort;              // a device-side wait gave up (status code); every CTA leaves after the current pass
};

// one stored point {x, y, z, pad} of the map or of the sorted frame: 32 bytes, one sector
struct __align__(32) KsPoint {
    double x, y, z, w;
};
#ifndef KS_EMU  // (tests/emu/ compiles this file for the host with its own versions of these few PTX helpers)
__device__ __forceinline__ unsigned long long ks_gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int ks_ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ks_ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// (a single 256-bit load: the map is read-only during a registration)
__device__ __forceinline__ KsPoint ks_ld_map_point(const double *p) {
    KsPoint r;
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
// a sorted frame point: written by this launch (S3), so it is read through L2, not through the read-only path
__device__ __forceinline__ KsPoint ks_ld_frame_point(const double *p) {
    KsPoint r;
    asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void ks_st_frame_point(double *dst, double x, double y, double z) {
    asm volatile("st.global.cg.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(x), "d"(y), "d"(z), "d"(0.0) : "memory");
}
#endif

// "norm(d2) < norm(best)" as the reference evaluates it (strict < on the rounded square roots; see kicp_register.cu)
__device__ __noinline__ bool ks_same_norm(double a, double b) { return sqrt(a) == sqrt(b); }
__device__ __forceinline__ bool ks_closer(double d2, double best) {
    if (!(d2 < best)) return false;
    if (d2 >= best * (1.0 - 4e-16)) return !ks_same_norm(d2, best);
    return true;
}
__device__ __forceinline__ double ks_dist2(double cx, double cy, double cz, double qx, double qy, double qz) {
    const double dx = cx - qx, dy = cy - qy, dz = cz - qz;
    return __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
}
__device__ __forceinline__ void ks_load_scan_point(const ScanView &sv, int i, double &x, double &y, double &z) {
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
__device__ __forceinline__ int ks_voxel_of(double x, double vs, double inv_vs, int pow2) {
    // PointToVoxel: floor(x / voxel_size); for a power-of-two voxel size the product with the (exact) reciprocal is the same double
    return pow2 ? (int)floor(x * inv_vs) : voxel_coord(x, vs);
}

// The shifts (bit k <-> voxel_shifts[k]) of one search stage whose cube can hold a point within `bound` (squared) of q, q in
// voxel (vx, vy, vz).  Stage 0: the own voxel; stage 1: the 6 faces; stage 2: the 12 edges and 8 corners.
__device__ __forceinline__ unsigned ks_stage_mask(int stage, bool valid, double bound, double qx, double qy, double qz, int vx, int vy, int vz,
                                                  double vs) {
    if (!valid) return 0u;
    if (stage == 0) return 1u;
    double t;
    t = (double)(vx + 1) * vs - qx; const double gxp = t * t;
    t = qx - (double)vx * vs;       const double gxm = t * t;
    t = (double)(vy + 1) * vs - qy; const double gyp = t * t;
    t = qy - (double)vy * vs;       const double gym = t * t;
    t = (double)(vz + 1) * vs - qz; const double gzp = t * t;
    t = qz - (double)vz * vs;       const double gzm = t * t;
    unsigned mask = 0u;
    if (stage == 1) {
        if (gxp <= bound) mask |= 1u << 1;
        if (gxm <= bound) mask |= 1u << 2;
        if (gyp <= bound) mask |= 1u << 3;
        if (gym <= bound) mask |= 1u << 4;
        if (gzp <= bound) mask |= 1u << 5;
        if (gzm <= bound) mask |= 1u << 6;
        return mask;
    }
#pragma unroll
    for (int kk = 7; kk < 27; ++kk) {  // edges and corners: the summed gap decides
        const double lb2 = (shift_x(kk) > 0 ? gxp : (shift_x(kk) < 0 ? gxm : 0.0)) + (shift_y(kk) > 0 ? gyp : (shift_y(kk) < 0 ? gym : 0.0)) +
                           (shift_z(kk) > 0 ? gzp : (shift_z(kk) < 0 ? gzm : 0.0));
        if (lb2 <= bound) mask |= 1u << kk;
    }
    return mask;
}

// Grid barrier number `index` (1, 2, ...) of this launch: every CTA of the cooperative grid arrives, thread 0 waits for all.
__device__ __forceinline__ void ks_grid_barrier(SortedState *st, unsigned index, unsigned long long timeout_ns) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&st->arrive, 1u);
        const unsigned target = index * gridDim.x;
        const unsigned long long deadline = ks_gtime_ns() + timeout_ns;
        while (ks_ld_acquire_gpu(&st->arrive) < target) {
            __nanosleep(40);
            if (ks_gtime_ns() > deadline) {  // a CTA of this grid never arrived: give up instead of hanging
                atomicExch(&st->abort, KICP_ERR_CUDA);
                break;
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// Where a voxel lives in the scratch table of the sort: blocks of 8 x-adjacent voxels stay adjacent (the lanes of a warp then
// hold points of neighbouring voxels, which share most of their 27-voxel neighbourhoods), the blocks themselves are hashed.
__device__ __forceinline__ uint32_t ks_home(int vx, int vy, int vz, uint32_t mask) {
    return ((voxel_hash(vx >> 3, vy, vz) << 3) | ((uint32_t)vx & 7u)) & mask;
}
__device__ __forceinline__ unsigned long long ks_key(int vx, int vy, int vz) {
    // 21 bits per axis; coordinates beyond +-2^20 voxels wrap around: the grouping is a locality heuristic, nothing depends on it
    return ((unsigned long long)((uint32_t)vx & 0x1FFFFFu)) | ((unsigned long long)((uint32_t)vy & 0x1FFFFFu) << 21) |
           ((unsigned long long)((uint32_t)vz & 0x1FFFFFu) << 42);
}

template <int DUMMY>
__global__ void __launch_bounds__(KS_THREADS, KS_MINB) k_register_sorted(const SortedArgs a) {
    __shared__ PoseState s_ps;
    __shared__ double s_part[KS_WARPS][8];
    __shared__ double s_sum[8];
    __shared__ unsigned s_scan[KS_WARPS];
    __shared__ unsigned s_base;

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xFFFFFFFFu;
    SortedState *const st = a.st;
    const int n = a.scan.d_n ? min(__ldg(a.scan.d_n), a.scan.n) : a.scan.n;
    const int num_chunks = (n + 31) >> 5;  // 32 consecutive points: one warp step
    const unsigned total_warps = gridDim.x * KS_WARPS;
    const unsigned gwarp = blockIdx.x * KS_WARPS + (unsigned)wid;
    const double vs = a.map.voxel_size, inv_vs = 1.0 / a.map.voxel_size;
    unsigned nbar = 0;  // grid barriers passed so far (identical in every thread of the grid)

    if (threadIdx.x == 0) {
        pose_init(&s_ps, a.init);
        if (blockIdx.x == 0) {
            result_init(a.result, &s_ps);
            if (a.init.iters_out) *a.init.iters_out = 0;
        }
    }
    __syncthreads();
    const unsigned long long t_start = ks_gtime_ns();

    // ----------------------------------------------------------------------------------------------------------------
    // S1: every point finds (or opens) the table slot of its voxel at the initial estimate and draws its rank in it
    // ----------------------------------------------------------------------------------------------------------------
    for (int w = (int)gwarp; w < num_chunks; w += (int)total_warps) {
        const int i = w * 32 + lane;
        const bool valid = i < n;
        if (a.up.flags != nullptr) {
            // a frame that is still being uploaded: wait until this chunk's piece has landed (the host raises one flag per piece)
            const uint32_t *f = a.up.flags + min(w / a.up.windows_per_chunk, KICP_UPLOAD_CHUNKS - 1);
            const unsigned long long deadline = ks_gtime_ns() + a.timeout_ns;
            bool pend = true;
            while (__any_sync(FULL, pend)) {  // warp-uniform: every lane polls the same word
                if (pend) {
                    if (ks_ld_acquire_sys(f) == a.up.seq) {
                        pend = false;
                    } else if (ks_gtime_ns() > deadline) {  // the copy never arrived
                        if (lane == 0) atomicExch(&st->abort, KICP_ERR_CUDA);
                        pend = false;
                    }
                }
            }
        }
        double px = 0, py = 0, pz = 0;
        if (valid) ks_load_scan_point(a.scan, i, px, py, pz);
        const double qx = s_ps.R[0] * px + s_ps.R[1] * py + s_ps.R[2] * pz + s_ps.t[0];
        const double qy = s_ps.R[3] * px + s_ps.R[4] * py + s_ps.R[5] * pz + s_ps.t[1];
        const double qz = s_ps.R[6] * px + s_ps.R[7] * py + s_ps.R[8] * pz + s_ps.t[2];
        // (a NaN or far-away coordinate gives some voxel or other: only the grouping depends on it)
        const int vx = ks_voxel_of(qx, vs, inv_vs, a.pow2_voxel), vy = ks_voxel_of(qy, vs, inv_vs, a.pow2_voxel),
                  vz = ks_voxel_of(qz, vs, inv_vs, a.pow2_voxel);
        const unsigned long long key = ks_key(vx, vy, vz);
        uint32_t idx = ks_home(vx, vy, vz, a.bin_mask);
        bool pend = valid;
        while (__any_sync(FULL, pend)) {  // warp-uniform loop (the table has at least 2 slots per point: it cannot fill up)
            if (pend) {
                const unsigned long long old = atomicCAS(&a.bin_key[idx], KS_EMPTY_KEY, key);
                if (old == KS_EMPTY_KEY || old == key) pend = false;
                else idx = (idx + 1u) & a.bin_mask;
            }
        }
        if (valid) {
            const unsigned rank = atomicAdd(&a.bin_cnt[idx], 1u);
            a.pslot[i] = make_uint2(idx, rank);
        }
        __syncwarp();
    }
    ks_grid_barrier(st, ++nbar, a.timeout_ns);

    // ----------------------------------------------------------------------------------------------------------------
    // S2: counts -> offsets.  Every CTA owns a contiguous slice of the table: it sums the slice, reserves that many sorted
    // positions with ONE atomic, and scans the slice tile by tile.  Occupied slots are reset to "empty" on the way (the table
    // is clean again for the next registration; the offsets are zeroed at the end of the launch).
    // ----------------------------------------------------------------------------------------------------------------
    {
        const uint32_t nslots = a.bin_mask + 1u;
        const uint32_t per = (nslots + gridDim.x - 1u) / gridDim.x;
        const uint32_t lo = min(nslots, blockIdx.x * per), hi = min(nslots, lo + per);
        unsigned sum = 0;
        for (uint32_t s = lo + threadIdx.x; s < hi; s += KS_THREADS) sum += __ldcg(&a.bin_cnt[s]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(FULL, sum, d);
        if (lane == 0) s_scan[wid] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned total = 0;
            for (int k = 0; k < KS_WARPS; ++k) total += s_scan[k];
            s_base = total ? atomicAdd(&st->cursor, total) : 0u;
        }
        __syncthreads();
        unsigned running = s_base;
        __syncthreads();
        for (uint32_t tile = lo; tile < hi; tile += KS_THREADS) {  // CTA-uniform trip count
            const uint32_t s = tile + threadIdx.x;
            const unsigned c = s < hi ? __ldcg(&a.bin_cnt[s]) : 0u;
            unsigned x = c;  // inclusive scan inside the warp
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned y = __shfl_up_sync(FULL, x, d);
                if (lane >= d) x += y;
            }
            if (lane == 31) s_scan[wid] = x;
            __syncthreads();
            unsigned woff = 0, tile_total = 0;
#pragma unroll
            for (int k = 0; k < KS_WARPS; ++k) {
                const unsigned v = s_scan[k];
                if (k < wid) woff += v;
                tile_total += v;
            }
            if (c) {
                __stcg(&a.bin_cnt[s], running + woff + x - c);
                a.bin_key[s] = KS_EMPTY_KEY;
            }
            running += tile_total;
            __syncthreads();
        }
    }
    ks_grid_barrier(st, ++nbar, a.timeout_ns);

    // ----------------------------------------------------------------------------------------------------------------
    // S3: scatter the frame into voxel order (32 bytes per point: one 256-bit load per pass and lane)
    // ----------------------------------------------------------------------------------------------------------------
    for (int w = (int)gwarp; w < num_chunks; w += (int)total_warps) {
        const int i = w * 32 + lane;
        if (i < n) {
            double px, py, pz;
            ks_load_scan_point(a.scan, i, px, py, pz);
            const uint2 sr = __ldcg(&a.pslot[i]);
            const unsigned pos = __ldcg(&a.bin_cnt[sr.x]) + sr.y;
            if (pos < (unsigned)n) {  // (always, unless the frame changed under the kernel: stay inside the buffer whatever happens)
                double *dst = a.sorted + (size_t)pos * 4;
                ks_st_frame_point(dst, px, py, pz);
            }
        }
        __syncwarp();  // (divergence safety, kicp_device.cuh: every loop with a divergent body ends in a hard convergence point)
    }
    ks_grid_barrier(st, ++nbar, a.timeout_ns);
    const unsigned long long t_sorted = ks_gtime_ns();

    unsigned long long n_probe = 0, n_cand = 0, n_line = 0;
    unsigned tbase = 0;  // first ticket of the current pass (identical in every warp of the grid)
    const double *const mpts = a.map.pts;
    const uint32_t mmask = a.map.mask;
    const int mcap = a.map.cap;

    for (unsigned it = 0; !s_ps.done; ++it) {
        const unsigned long long t_iter0 = ks_gtime_ns();
        double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0, acc6 = 0;
        const double tau = s_ps.tau, tau2 = s_ps.tau * s_ps.tau;
        // the first chunk of every warp is its own index, the rest are handed out through one monotonic ticket counter
        // (chunks differ in cost: dense voxels bring many candidates)
        int w = min((int)gwarp, num_chunks);
        const unsigned dyn = (unsigned)max(num_chunks - (int)total_warps, 0);  // chunks behind tickets
        unsigned tk = 0;
        while (w < num_chunks) {
            if (lane == 0) tk = atomicAdd(&st->win_ctr, 1u);
            const int t = w * 32 + lane;
            const bool valid = t < n;
            KsPoint p;
            p.x = 0, p.y = 0, p.z = 0, p.w = 0;
            if (valid) p = ks_ld_frame_point(a.sorted + (size_t)t * 4);
            const double qx = s_ps.R[0] * p.x + s_ps.R[1] * p.y + s_ps.R[2] * p.z + s_ps.t[0];
            const double qy = s_ps.R[3] * p.x + s_ps.R[4] * p.y + s_ps.R[5] * p.z + s_ps.t[1];
            const double qz = s_ps.R[6] * p.x + s_ps.R[7] * p.y + s_ps.R[8] * p.z + s_ps.t[2];
            const int vx = ks_voxel_of(qx, vs, inv_vs, a.pow2_voxel), vy = ks_voxel_of(qy, vs, inv_vs, a.pow2_voxel),
                      vz = ks_voxel_of(qz, vs, inv_vs, a.pow2_voxel);
            // The previous pass's neighbour, if it is still a point of the 27 voxels around the new position, bounds the search
            // from the start: nothing farther than it can be the answer (ties are still resolved by the visiting order below,
            // because everything at that distance or closer is evaluated).
            double seed2 = DBL_MAX;
            if (it > 0u && valid) {
                const unsigned g = __ldcg(&a.nn_g[t]);
                if (g != KS_NONE) {
                    const KsPoint c = ks_ld_map_point(mpts + (size_t)g * KICP_PSTRIDE);
                    const bool inside = abs(ks_voxel_of(c.x, vs, inv_vs, a.pow2_voxel) - vx) <= 1 &&
                                        abs(ks_voxel_of(c.y, vs, inv_vs, a.pow2_voxel) - vy) <= 1 &&
                                        abs(ks_voxel_of(c.z, vs, inv_vs, a.pow2_voxel) - vz) <= 1;
                    if (inside) seed2 = ks_dist2(c.x, c.y, c.z, qx, qy, qz) * (1.0 + 1e-6);
                }
            }
            double best = DBL_MAX;    // squared distance to the nearest candidate so far ...
            unsigned bidx = KS_NONE;  // ... and its index in the map's point array
#pragma unroll 1
            for (int stage = 0; stage < 3; ++stage) {
                // exact pruning bound: the best squared distance so far, never more than the gate (a neighbour at tau or beyond is
                // rejected anyway, Registration.cpp:75) or the seed
                const double bsq = fmin(fmin(tau2, best), seed2);
                const double bound = bsq * (1.0 + 1e-6) + 1e-10;
                unsigned mask = ks_stage_mask(stage, valid, bound, qx, qy, qz, vx, vy, vz, vs);
                // Every lane walks ITS OWN surviving shifts in visiting order; the loop runs as long as any lane has one left
                // (warp-uniform trip count: the longest list of the warp, 1 for the own voxel, 2-3 for faces, 1-2 for edges and corners).
                // Lanes of the same voxel mostly hold the same list and ask for the same slots and points together.
                while (__any_sync(FULL, mask != 0u)) {
                    const bool act = mask != 0u;
                    const int k = act ? __ffs(mask) - 1 : 0;
                    mask &= mask - 1u;
                    KS_COUNT(0 + (it ? 4 : 0), lane == 0 ? 1 : 0)
                    KS_COUNT(2 + (it ? 4 : 0), act ? 1 : 0)
                    const int kx = vx + shift_x(k), ky = vy + shift_y(k), kz = vz + shift_z(k);
                    uint32_t h = voxel_hash(kx, ky, kz) & mmask;
                    uint32_t meta = KICP_SLOT_EMPTY;
                    {
                        // the home slot and the next one travel together: with a load factor <= 0.25 a longer chain is rare
                        const int4 s0 = __ldg(&a.map.slots[h]);
                        const int4 s1 = __ldg(&a.map.slots[(h + 1u) & mmask]);
                        bool pend = false;
                        if (act && (uint32_t)s0.w != KICP_SLOT_EMPTY) {
                            if (s0.x == kx && s0.y == ky && s0.z == kz) {
                                meta = (uint32_t)s0.w;
                            } else if ((uint32_t)s1.w != KICP_SLOT_EMPTY) {
                                if (s1.x == kx && s1.y == ky && s1.z == kz) meta = (uint32_t)s1.w;
                                else pend = true, h = (h + 2u) & mmask;
                            }
                        }
                        while (__any_sync(FULL, pend)) {  // warp-uniform loop
                            if (pend) {
                                const int4 sl = __ldg(&a.map.slots[h]);
                                if ((uint32_t)sl.w == KICP_SLOT_EMPTY) pend = false;
                                else if (sl.x == kx && sl.y == ky && sl.z == kz) meta = (uint32_t)sl.w, pend = false;
                                else h = (h + 1u) & mmask;
                            }
                        }
                    }
                    const int cnt = meta == KICP_SLOT_EMPTY ? 0 : (int)(meta & 0xFFu);
                    const unsigned first = meta == KICP_SLOT_EMPTY ? 0u : (meta >> 8) * (unsigned)mcap;  // the run's first point
                    if (a.collect_stats) n_probe += act ? 1 : 0, n_cand += cnt, n_line += (cnt + 3) >> 2;
                    const int maxcnt = __reduce_max_sync(FULL, cnt);
                    KS_COUNT(1 + (it ? 4 : 0), lane == 0 ? (maxcnt + 1) / 2 : 0)
                    KS_COUNT(3 + (it ? 4 : 0), cnt)
                    for (int j = 0; j < maxcnt; j += 2) {  // warp-uniform trip count; two independent 256-bit loads in flight
                        const bool h0 = j < cnt, h1 = j + 1 < cnt;
                        const KsPoint c0 = ks_ld_map_point(mpts + (size_t)(first + (h0 ? (unsigned)j : 0u)) * KICP_PSTRIDE);
                        const KsPoint c1 = ks_ld_map_point(mpts + (size_t)(first + (h1 ? (unsigned)j + 1u : 0u)) * KICP_PSTRIDE);
                        const double d0 = ks_dist2(c0.x, c0.y, c0.z, qx, qy, qz);
                        const double d1 = ks_dist2(c1.x, c1.y, c1.z, qx, qy, qz);
                        if (h0 && ks_closer(d0, best)) best = d0, bidx = first + (unsigned)j;
                        if (h1 && ks_closer(d1, best)) best = d1, bidx = first + (unsigned)j + 1u;
                    }
                }
            }
            // ---------------------------------------------------------------- gate, residual, Jacobian, sums
            {
                const bool have = valid && bidx != KS_NONE;
                const KsPoint c = ks_ld_map_point(mpts + (size_t)(have ? bidx : 0u) * KICP_PSTRIDE);
                const double rx = qx - c.x, ry = qy - c.y, rz = qz - c.z;  // r = T p - n
                const double rr = rx * rx + ry * ry + rz * rz;
                if (have && sqrt(rr) < tau) {  // distance < max_correspondance_distance   (Registration.cpp:75)
                    // J = [R e_x | R (-p_y, p_x, 0)]      (Registration.cpp:89-91)
                    const double c0x = s_ps.R[0], c0y = s_ps.R[3], c0z = s_ps.R[6];
                    const double c1x = s_ps.R[1] * p.x - s_ps.R[0] * p.y, c1y = s_ps.R[4] * p.x - s_ps.R[3] * p.y,
                                 c1z = s_ps.R[7] * p.x - s_ps.R[6] * p.y;
                    acc0 += c0x * c0x + c0y * c0y + c0z * c0z;
                    acc1 += c0x * c1x + c0y * c1y + c0z * c1z;
                    acc2 += c1x * c1x + c1y * c1y + c1z * c1z;
                    acc3 += c0x * rx + c0y * ry + c0z * rz;
                    acc4 += c1x * rx + c1y * ry + c1z * rz;
                    acc5 += 1.0;
                    acc6 += rr;
                }
                if (valid) __stcg(&a.nn_g[t], bidx);
            }
            __syncwarp();
            w = (int)(total_warps + min(__shfl_sync(FULL, tk, 0) - tbase, dyn));  // >= num_chunks once the tickets are used up
        }
        // every processed chunk drew exactly one ticket
        tbase += (unsigned)num_chunks;

        const unsigned long long t_win = ks_gtime_ns();
        // ---------------------------------------------------------------- lane -> warp -> CTA partial (plain stores, fixed order)
        double v[7] = {acc0, acc1, acc2, acc3, acc4, acc5, acc6};
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
        double *const part = a.partials + (size_t)(it & 1u) * gridDim.x * 8;
        if (threadIdx.x < 8) {
            double s = 0.0;
            for (int k = 0; k < KS_WARPS; ++k) s += s_part[k][threadIdx.x];
            __stcg(&part[(size_t)blockIdx.x * 8 + threadIdx.x], s);
        }
        ks_grid_barrier(st, ++nbar, a.timeout_ns);
        const unsigned long long t_arr = ks_gtime_ns();
        {
            // every CTA: column k = thread & 7, rows strided by KS_THREADS / 8 — a fixed summation tree, identical in every CTA
            const int col = threadIdx.x & 7, row0 = threadIdx.x >> 3;
            double s = 0.0;
            const unsigned nrow = (gridDim.x + KS_THREADS / 8 - 1) / (KS_THREADS / 8);  // uniform trip count, tail predicated
            for (unsigned kr = 0; kr < nrow; ++kr) {
                const unsigned bb = kr * (KS_THREADS / 8) + (unsigned)row0;
                if (bb < gridDim.x) s += __ldcg(&part[(size_t)bb * 8 + col]);
            }
            s += __shfl_xor_sync(FULL, s, 8);
            s += __shfl_xor_sync(FULL, s, 16);
            if (lane < 8) s_part[wid][lane] = s;
            __syncthreads();
            if (threadIdx.x < 8) {
                double tsum = 0.0;
                for (int k = 0; k < KS_WARPS; ++k) tsum += s_part[k][threadIdx.x];
                s_sum[threadIdx.x] = tsum;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const unsigned long long t_red = ks_gtime_ns();
            const int ab = (int)ks_ld_acquire_gpu((const unsigned int *)&st->abort);
            if (ab) {  // a wait gave up: keep the last pose, report, leave
                s_ps.status = ab, s_ps.done = 1;
                if (blockIdx.x == 0) a.result->status = ab;
            } else {
                double s[8];
                for (int k = 0; k < 8; ++k) s[k] = s_sum[k];
                solve_and_update(&s_ps, s, blockIdx.x == 0 ? a.result : nullptr, blockIdx.x == 0 ? a.init.iters_out : nullptr);
            }
            if (blockIdx.x == 0 && it < KICP_MAX_ITERATIONS && a.dbg != nullptr) {
                // per pass, ns: [sort (pass 0 only), -, search (CTA 0), barrier wait, reduce, solve]
                double *d = a.dbg + (size_t)it * 6;
                d[0] = it == 0u ? (double)(t_sorted - t_start) : 0.0, d[1] = 0.0, d[2] = (double)(t_win - t_iter0);
                d[3] = (double)(t_arr - t_win), d[4] = (double)(t_red - t_arr), d[5] = (double)(ks_gtime_ns() - t_red);
            }
        }
        __syncthreads();
    }

    if (a.collect_stats) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            n_probe += __shfl_xor_sync(FULL, n_probe, d);
            n_cand += __shfl_xor_sync(FULL, n_cand, d);
            n_line += __shfl_xor_sync(FULL, n_line, d);
        }
        if (lane == 0) atomicAdd(&a.stats[0], n_probe), atomicAdd(&a.stats[1], n_cand), atomicAdd(&a.stats[2], n_line);
    }
    // the offsets of this registration go back to zero (nobody reads them after S3; every CTA is past several barriers since)
    {
        const uint32_t nslots = a.bin_mask + 1u;
        for (uint32_t s = blockIdx.x * KS_THREADS + threadIdx.x; s < nslots; s += gridDim.x * KS_THREADS)
            if (__ldcg(&a.bin_cnt[s]) != 0u) a.bin_cnt[s] = 0u;
    }
    __syncthreads();
    // the result block goes straight to the caller's page-locked host memory (no copy-engine operation after the kernel)
    if (blockIdx.x == 0 && a.result_host != nullptr) {
        const double *src = reinterpret_cast<const double *>(a.result);
        double *dst = reinterpret_cast<double *>(a.result_host);
        for (unsigned i = threadIdx.x; i < sizeof(kicp_reg_result) / sizeof(double); i += KS_THREADS) dst[i] = __ldcg(src + i);
        __threadfence_system();
    }
    // the last CTA to leave zeroes the counters for the next registration on this stream
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned left = atomicAdd(&st->exit_ctr, 1u);
        if (left == gridDim.x - 1) {
            st->win_ctr = 0, st->arrive = 0, st->abort = 0, st->cursor = 0;
            __threadfence();
            st->exit_ctr = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------- entry points for the API file
#ifndef KS_EMU
size_t ks_state_bytes() { return sizeof(SortedState); }
int ks_threads() { return KS_THREADS; }
cudaError_t ks_prepare(int *ctas_per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_register_sorted<0>, KS_THREADS, 0);
}
cudaError_t ks_launch(int grid, SortedArgs &ka, cudaStream_t stream) {
    void *args[] = {&ka};  // cooperative: every CTA resident (the grid barriers inside the kernel rely on it)
    return cudaLaunchCooperativeKernel((const void *)k_register_sorted<0>, dim3(grid), dim3(KS_THREADS), args, 0, stream);
}
#endif
