// A small SIMT emulator for host-side tests of a CUDA kernel's LOGIC (test infrastructure only; never linked by the product).
//
// The kernel source is compiled unchanged by g++ against this header.  Every CUDA thread is a fiber (ucontext); the 32 fibers of
// a warp meet at every warp collective (__shfl*_sync, __any_sync, __reduce_*_sync, __syncwarp), the fibers of a CTA meet at
// __syncthreads().  One OS thread runs one CTA (so `__shared__` maps to `static thread_local`); CTAs of a grid run concurrently
// and communicate through real atomics, which is what a cooperative launch guarantees on the device.  What this checks:
// indexing, phase/barrier structure, reductions, the search and tie rules — with the kernel's own code.  What it cannot check:
// the GPU memory model, launch limits, register-level hazards.
#pragma once
#include <sched.h>
#include <time.h>
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <memory>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
// (after EVERY standard header a harness needs — include them above, not in the harness: cuda's host_defines.h turns __noinline__
// into a macro, which libstdc++ uses as an attribute name)
#include <cuda_runtime.h>

#undef __shared__
#define __shared__ static thread_local
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __host__
#define __host__
#undef __launch_bounds__
#define __launch_bounds__(...)
#ifndef __forceinline__
#define __forceinline__ inline __attribute__((always_inline))
#endif
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

using std::max;
using std::min;

namespace emu {
constexpr int kMaxWarps = 32;
struct Idx3 {
    unsigned x, y, z;
};
struct Warp {
    uint64_t vals[2][32];
    long count[2];
};
struct Cta {
    Warp warps[kMaxWarps];
    long bar_count;
    int nthreads;
};
struct Fiber {
    ucontext_t ctx;
    Idx3 tid;
    int lane, wid;
    long wcoll[2];
    int wphase;
    long bars;
    bool done;
    std::vector<char> stack;
};
inline thread_local Cta *cta = nullptr;
inline thread_local Fiber *cur = nullptr;
inline thread_local ucontext_t sched_ctx;
inline thread_local Idx3 block_idx = {0, 0, 0};
inline Idx3 grid_dim = {1, 1, 1};
inline Idx3 block_dim = {1, 1, 1};

inline void yield() { swapcontext(&cur->ctx, &sched_ctx); }

// Interleaving stress (environment KICP_EMU_CHAOS = seed, 0 / unset = off): every global-memory access that can order threads — the
// atomics, __ldcg / __stcg, the fences — additionally gives the fiber's turn away with probability 1/4, so that the lanes of a warp and
// the warps of a CTA run in many more orders than the plain round robin produces.  Collectives keep their meaning (all 32 lanes meet).
// A harness opts in with `#define EMU_CHAOS 1` before this header — and must not, if its kernels hold spin locks: the fibers of a CTA share
// one OS thread, so a lock holder that gives its turn away inside the critical section starves the spinners for ever (on the device the
// holder simply keeps running).  The registration kernel has no such locks: all of its waits are collectives or sleep.
inline unsigned chaos_seed() {
    static const unsigned seed = []() {
        const char *e = getenv("KICP_EMU_CHAOS");
        return e ? (unsigned)strtoul(e, nullptr, 10) : 0u;
    }();
    return seed;
}
inline thread_local uint64_t chaos_state = 0;
inline void chaos_point() {
#ifndef EMU_CHAOS
    return;
#endif
    if (!chaos_seed() || !cur) return;
    if (!chaos_state) chaos_state = 0x9E3779B97F4A7C15ull * (chaos_seed() + 1) + (uint64_t)(uintptr_t)cta;
    chaos_state ^= chaos_state << 13, chaos_state ^= chaos_state >> 7, chaos_state ^= chaos_state << 17;
    if ((chaos_state & 3u) == 0u) yield();
}

// all 32 lanes of the calling fiber's warp deposit a value and leave together; returns the buffer holding the 32 values
inline const uint64_t *exchange(uint64_t v) {
    Fiber *f = cur;
    Warp &w = cta->warps[f->wid];
    const int p = f->wphase;
    w.vals[p][f->lane] = v;
    w.count[p]++;
    const long target = 32 * (f->wcoll[p] + 1);
    while (w.count[p] < target) yield();
    f->wcoll[p]++;
    f->wphase ^= 1;
    return w.vals[p];  // valid until this lane enters its next-but-one collective (the caller reads it at once)
}
inline void syncthreads() {
    Fiber *f = cur;
    cta->bar_count++;
    const long target = (long)cta->nthreads * (f->bars + 1);
    while (cta->bar_count < target) yield();
    f->bars++;
}

template <class T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

// one CTA of `threads` threads on the calling OS thread (fibers and their stacks are reused from CTA to CTA)
template <class F>
inline void run_cta(int b, int threads, F &body, std::vector<Fiber> &fibers) {
    Cta c;
    memset(&c, 0, sizeof(c));
    c.nthreads = threads;
    cta = &c;
    block_idx = {(unsigned)b, 0, 0};
    struct Tramp {
        static void run(unsigned lo, unsigned hi) {
            F *fn = reinterpret_cast<F *>(((uintptr_t)hi << 32) | lo);
            (*fn)();
            cur->done = true;
            swapcontext(&cur->ctx, &sched_ctx);
        }
    };
    if ((int)fibers.size() != threads) fibers.resize(threads);
    for (int t = 0; t < threads; ++t) {
        Fiber &f = fibers[t];
        f.tid = {(unsigned)t, 0, 0}, f.lane = t & 31, f.wid = t >> 5;
        f.wcoll[0] = f.wcoll[1] = 0, f.wphase = 0, f.bars = 0, f.done = false;
        if (f.stack.empty()) f.stack.resize(256 << 10);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data(), f.ctx.uc_stack.ss_size = f.stack.size(), f.ctx.uc_link = nullptr;
        const uintptr_t pf = (uintptr_t)&body;
        makecontext(&f.ctx, (void (*)())Tramp::run, 2, (unsigned)(pf & 0xFFFFFFFFu), (unsigned)(pf >> 32));
    }
    int alive = threads;
    while (alive > 0) {
        alive = 0;
        for (int t = 0; t < threads; ++t) {
            if (fibers[t].done) continue;
            cur = &fibers[t];
            swapcontext(&sched_ctx, &fibers[t].ctx);
            if (!fibers[t].done) alive++;
        }
    }
    cta = nullptr;
}

// run `kernel_body` as a grid of `grid` CTAs x `threads` threads (threads a multiple of 32), ALL CTAs resident at once — one OS
// thread each — which is what a cooperative launch guarantees and what a kernel with grid-wide barriers needs
template <class F>
inline void launch(int grid, int threads, F kernel_body) {
    grid_dim = {(unsigned)grid, 1, 1};
    block_dim = {(unsigned)threads, 1, 1};
    std::vector<std::thread> pool;
    for (int b = 0; b < grid; ++b) {
        pool.emplace_back([=]() {
            F body = kernel_body;
            std::vector<Fiber> fibers;
            run_cta(b, threads, body, fibers);
        });
    }
    for (auto &t : pool) t.join();
}

// an ordinary launch: CTAs are handed to `workers` OS threads in index order, so only some are resident at a time — fine for
// kernels whose CTAs meet through atomics and short critical sections only (no grid-wide barrier)
// (environment KICP_EMU_WORKERS overrides the number of OS threads: 1 = CTAs strictly one after the other, 16+ = many CTAs in flight)
inline int wave_workers(int dflt) {
    static const int env = []() {
        const char *e = getenv("KICP_EMU_WORKERS");
        return e ? atoi(e) : 0;
    }();
    return env > 0 ? env : dflt;
}
template <class F>
inline void launch_waves(int grid, int threads, F kernel_body, int workers = 6) {
    workers = wave_workers(workers);
    grid_dim = {(unsigned)grid, 1, 1};
    block_dim = {(unsigned)threads, 1, 1};
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    for (int w = 0; w < std::min(workers, grid); ++w) {
        pool.emplace_back([&next, grid, threads, kernel_body]() {
            F body = kernel_body;
            std::vector<Fiber> fibers;
            for (int b = next.fetch_add(1); b < grid; b = next.fetch_add(1)) run_cta(b, threads, body, fibers);
        });
    }
    for (auto &t : pool) t.join();
}
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::block_idx)
#define gridDim (emu::grid_dim)
#define blockDim (emu::block_dim)

// ---- warp collectives (full mask only: the kernels under test use nothing else)
#define EMU_FULL_ONLY(m) \
    if ((m) != 0xFFFFFFFFu) { fprintf(stderr, "emu: partial-mask collective\n"); abort(); }
template <class T>
inline T __shfl_sync(unsigned m, T v, int src) {
    EMU_FULL_ONLY(m)
    return emu::from_bits<T>(emu::exchange(emu::to_bits(v))[src & 31]);
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int d) {
    EMU_FULL_ONLY(m)
    const int lane = emu::cur->lane;
    return emu::from_bits<T>(emu::exchange(emu::to_bits(v))[(lane ^ d) & 31]);
}
template <class T>
inline T __shfl_up_sync(unsigned m, T v, int d) {
    EMU_FULL_ONLY(m)
    const int lane = emu::cur->lane;
    const uint64_t *b = emu::exchange(emu::to_bits(v));
    return lane >= d ? emu::from_bits<T>(b[lane - d]) : v;
}
inline unsigned __ballot_sync(unsigned m, bool p) {
    EMU_FULL_ONLY(m)
    const uint64_t *b = emu::exchange(p ? 1 : 0);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(b[i] & 1) << i;
    return r;
}
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xFFFFFFFFu; }
inline unsigned __reduce_or_sync(unsigned m, unsigned v) {
    EMU_FULL_ONLY(m)
    const uint64_t *b = emu::exchange(v);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)b[i];
    return r;
}
inline int __reduce_max_sync(unsigned m, int v) {
    EMU_FULL_ONLY(m)
    const uint64_t *b = emu::exchange(emu::to_bits(v));
    int r = INT32_MIN;
    for (int i = 0; i < 32; ++i) r = std::max(r, emu::from_bits<int>(b[i]));
    return r;
}
inline void __syncwarp(unsigned m = 0xFFFFFFFFu) {
    EMU_FULL_ONLY(m)
    emu::exchange(0);
}
inline void __syncthreads() { emu::syncthreads(); }

// ---- memory
template <class T>
inline T __ldg(const T *p) { return *(const volatile T *)p; }
template <class T>
inline T __ldcg(const T *p) {
    emu::chaos_point();
    return *(const volatile T *)p;
}
inline int4 __ldg(const int4 *p) { int4 v; memcpy(&v, p, sizeof(v)); return v; }
#ifndef __restrict__
#define __restrict__ __restrict
#endif
inline uint2 __ldcg(const uint2 *p) { uint2 v; memcpy(&v, p, sizeof(v)); return v; }
template <class T>
inline void __stcg(T *p, T v) {
    emu::chaos_point();
    *(volatile T *)p = v;
}
inline void __threadfence() {
    emu::chaos_point();
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned *p, unsigned v) {
    emu::chaos_point();
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int *p, int v) {
    emu::chaos_point();
    return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST);
}
inline int atomicMin(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {
    }
    return old;
}
inline unsigned atomicExch(unsigned *p, unsigned v) {
    emu::chaos_point();
    return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST);
}
inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) {
    emu::chaos_point();
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
inline int atomicExch(volatile int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
inline void __trap() {
    fprintf(stderr, "emu: __trap()\n");
    abort();
}
inline void __nanosleep(unsigned) {
    emu::yield();
    sched_yield();
}

// ---- arithmetic intrinsics
inline float __double2float_rz(double x) {  // round towards zero
    float f = (float)x;
    if (std::fabs((double)f) > std::fabs(x)) f = std::nextafterf(f, 0.0f);
    return f;
}
inline float __double2float_ru(double x) {  // round up
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, INFINITY);
    return f;
}
inline long long __double_as_longlong(double x) { long long v; memcpy(&v, &x, 8); return v; }
inline double __longlong_as_double(long long x) { double v; memcpy(&v, &x, 8); return v; }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline double __fma_rn(double a, double b, double c) { return __builtin_fma(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
