// Single-pass prefix sums across the CTAs of ONE ordinary launch (chained scan with decoupled look-back), the building block of the
// library's own order-preserving stream compaction and exclusive sum: the front end's stable selects (Preprocess, VoxelDownsample —
// fused with the kernels that used to write a flag array for a library select) and the block renumbering of the map's eviction and
// export.  No library kernels, no temporary flag arrays, no init launch.
//
// A CTA owns a *tile* of kScanTile consecutive items.  Tiles are handed out by a ticket counter in launch order, so a tile only
// ever waits for tiles that are already running (no assumption about the order in which the hardware starts CTAs).  Per tile one
// 64-bit status word  { launch number : 32 | state : 2 | value : 30 }  is written with a single aligned 8-byte store, which arrives
// whole: a word that carries this launch's number is valid, whatever else is in flight — no fence, no separate flag, and nothing to
// clear between launches (words of earlier launches simply read as "not written yet").  A tile publishes its own total (state
// AGGREGATE) as soon as it knows it, then warp 0 looks back over its predecessors, 32 tiles per step, adding totals until it meets a
// tile that already knows its inclusive prefix (state PREFIX), and publishes its own inclusive prefix in turn.
//
// Inside a tile the items are dealt to the lanes warp-striped (item k of lane l of warp w = tile*1024 + w*128 + k*32 + l): global
// loads are coalesced, and the rank of a selected item among the tile's selected items in INDEX order falls out of one ballot per k.
// Every loop has a warp-uniform trip count (DIVERGENCE SAFETY, kicp_device.cuh).  Values must stay below 2^30 (checked by the hosts).
#pragma once
#include <stdint.h>

#include "kicp_internal.h"

namespace kicp_dev {
constexpr int kScanThreads = 256, kScanItems = 4, kScanWarps = kScanThreads / 32;
constexpr int kScanTile = kScanThreads * kScanItems;  // items per CTA
constexpr unsigned long long kScanAggregate = 1ull, kScanPrefix = 2ull;
constexpr unsigned kScanSpinLimit = 1u << 26;  // polls of one look-back step before the kernel traps instead of hanging the GPU

__device__ __forceinline__ unsigned long long scan_word(uint32_t launch, unsigned long long state, uint32_t value) {
    return ((unsigned long long)launch << 32) | (state << 30) | (unsigned long long)(value & 0x3FFFFFFFu);
}

// index of the first item that (warp, lane) handles in `tile`; its k-th item is 32 * k further
__device__ __forceinline__ int64_t scan_first_item(uint32_t tile) {
    return (int64_t)tile * kScanTile + (int64_t)(threadIdx.x >> 5) * (32 * kScanItems) + (threadIdx.x & 31);
}

// The tile of this CTA (all threads call; s_word: one shared word).  The CTA that draws the last ticket puts the counter back to zero:
// every other CTA of the launch has drawn by then, and the next launch on the stream starts after this one has finished.
__device__ __forceinline__ uint32_t scan_take_tile(const kicp_scan_args &a, uint32_t *s_word) {
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(a.ticket, 1u);
        if (t == gridDim.x - 1) atomicExch(a.ticket, 0u);
        *s_word = t;
    }
    __syncthreads();
    const uint32_t t = *(volatile uint32_t *)s_word;
    __syncthreads();
    return t;
}

// Ranks of the flagged items of one warp in index order: rank[k] = flagged items of the warp before item k of this lane
__device__ __forceinline__ uint32_t scan_warp_ranks(const bool (&flag)[kScanItems], uint32_t (&rank)[kScanItems]) {
    const unsigned below = (1u << (threadIdx.x & 31)) - 1u;
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const unsigned b = __ballot_sync(0xFFFFFFFFu, flag[k]);
        rank[k] = run + (uint32_t)__popc(b & below);
        run += (uint32_t)__popc(b);
    }
    return run;  // the warp's total
}

// Exclusive prefix of one warp's values in index order (value k of lane l before value k of lane l+1 before value k+1 of lane 0)
__device__ __forceinline__ uint32_t scan_warp_values(const uint32_t (&v)[kScanItems], uint32_t (&excl)[kScanItems]) {
    const int lane = threadIdx.x & 31;
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint32_t inc = v[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += up;
        }
        excl[k] = run + inc - v[k];
        run += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    return run;
}

// All threads of the CTA call this with their warp's total.  Returns the number of flagged items (the sum of the values) in front of
// this WARP over the whole launch; `tile_inclusive` receives the count up to and including this tile.  s_warp: kScanWarps + 1 shared
// words.
__device__ __forceinline__ uint32_t scan_offset(const kicp_scan_args &a, uint32_t tile, uint32_t warp_total, uint32_t *s_warp,
                                                uint32_t &tile_inclusive) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = warp_total;
    __syncthreads();
    uint32_t before = 0, aggregate = 0;
#pragma unroll
    for (int w = 0; w < kScanWarps; ++w) {
        const uint32_t t = ((volatile uint32_t *)s_warp)[w];
        before += w < warp ? t : 0u;
        aggregate += t;
    }
    if (warp == 0) {
        volatile unsigned long long *status = a.status;
        uint32_t exclusive = 0;
        if (tile == 0) {
            if (lane == 0) status[0] = scan_word(a.launch, kScanPrefix, aggregate);
        } else {
            if (lane == 0) status[tile] = scan_word(a.launch, kScanAggregate, aggregate);
            int64_t first = (int64_t)tile - 1;  // lane l looks at tile first - l
            bool met_prefix = false;
            while (!met_prefix) {  // warp-uniform: decided by a ballot
                const int64_t p = first - lane;
                unsigned long long w = 0;
                unsigned polls = 0;
                for (;;) {  // left by all lanes together
                    bool ready = true;
                    if (p >= 0) {
                        w = status[p];
                        ready = (uint32_t)(w >> 32) == a.launch;
                    }
                    if (__all_sync(0xFFFFFFFFu, ready)) break;
                    if (++polls > kScanSpinLimit) __trap();
                    __nanosleep(20);
                }
                const bool is_prefix = p < 0 || ((w >> 30) & 3ull) == kScanPrefix;  // in front of tile 0 there is nothing: prefix 0
                const uint32_t value = p < 0 ? 0u : (uint32_t)(w & 0x3FFFFFFFull);
                const unsigned prefixes = __ballot_sync(0xFFFFFFFFu, is_prefix);
                const int nearest = prefixes ? __ffs((int)prefixes) - 1 : 31;  // every tile up to the nearest prefix counts
                uint32_t part = lane <= nearest ? value : 0u;
#pragma unroll
                for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, d);
                exclusive += part;
                met_prefix = prefixes != 0u;
                first -= 32;
            }
            if (lane == 0) status[tile] = scan_word(a.launch, kScanPrefix, exclusive + aggregate);
        }
        if (lane == 0) s_warp[kScanWarps] = exclusive;
    }
    __syncthreads();
    const uint32_t exclusive = ((volatile uint32_t *)s_warp)[kScanWarps];
    tile_inclusive = exclusive + aggregate;
    return exclusive + before;
}

}  // namespace kicp_dev
