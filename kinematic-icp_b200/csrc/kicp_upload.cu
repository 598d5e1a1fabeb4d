// Host-to-device upload of a caller's buffer.  Page-locked memory (kicp_host_alloc, cudaHostRegister) goes to the copy engine
// directly; ordinary (pageable) memory — what the reference's callers hold: std::vector<Eigen::Vector3d>, the byte vector of a
// PointCloud2 message (RosUtils.cpp:30-39) — is staged into the context's own page-locked area by a few helper threads
// (kicp_stager.hpp) while the calling thread hands every finished piece to the copy engine, so staging and DMA overlap.
// Option "upload_threads" (environment KICP_UPLOAD_THREADS): helper threads, 0 = leave pageable copies to the driver.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "kicp_internal.h"
#include "kicp_stager.hpp"

namespace {
constexpr size_t kGranule = 128 << 10;     // bytes one helper copies at a time
constexpr size_t kPiece = 512 << 10;       // bytes per device copy of a plain upload
constexpr size_t kMinStaged = 256 << 10;   // smaller uploads are not worth waking anybody

struct UploadScratch {
    kicp::Stager *stager = nullptr;
    int helpers = -1;
    unsigned char *h_stage[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool pending[2] = {false, false};  // a device copy out of the slot's staging area may still be in flight
};

void upload_free(kicp_ctx *c) {
    UploadScratch *u = static_cast<UploadScratch *>(c->upload);
    if (!u) return;
    delete u->stager;
    for (int k = 0; k < 2; ++k) {
        if (u->h_stage[k]) cudaFreeHost(u->h_stage[k]);
        if (u->ev[k]) cudaEventDestroy(u->ev[k]);
    }
    delete u;
    c->upload = nullptr;
}

bool is_pageable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}
}  // namespace

// Everything a staged upload of `total` bytes from `src` needs before its first byte moves: the helper threads, a staging area of
// that size, the previous upload out of that area finished.  Called by kicp_h2d_groups itself; a caller that launches a kernel
// which WAITS for the upload calls it before the launch (allocating page-locked memory may wait for the device).
int kicp_h2d_prepare(kicp_ctx *c, int slot, const void *src, size_t total) {
    if (!c || slot < 0 || slot > 1) return KICP_ERR_INVALID;
    if (!(c->upload_threads > 0 && total >= kMinStaged && is_pageable(src))) return KICP_OK;
    UploadScratch *u = static_cast<UploadScratch *>(c->upload);
    if (!u) {
        u = new UploadScratch();
        c->upload = u, c->upload_free = upload_free;
    }
    if (!u->stager || u->helpers != c->upload_threads) {
        delete u->stager;
        u->stager = new kicp::Stager(c->upload_threads);
        u->helpers = c->upload_threads;
    }
    if (u->pending[slot]) {  // the previous upload out of this staging area must have left it
        KICP_CUDA(cudaEventSynchronize(u->ev[slot]));
        u->pending[slot] = false;
    }
    if (total > u->cap[slot]) {
        if (u->h_stage[slot]) cudaFreeHost(u->h_stage[slot]);
        u->h_stage[slot] = nullptr, u->cap[slot] = 0;
        const size_t cap = total + total / 4 + (1 << 20);
        KICP_CUDA(cudaMallocHost((void **)&u->h_stage[slot], cap));
        u->cap[slot] = cap;
    }
    if (!u->ev[slot]) KICP_CUDA(cudaEventCreateWithFlags(&u->ev[slot], cudaEventDisableTiming));
    return KICP_OK;
}

// Copy src[bounds[0], bounds[ngroups]) to the same offsets of dst, group by group in order ([bounds[g], bounds[g+1]) is one
// device copy on `stream`); after_group(user, g), if given, runs right after group g's copy has been issued (the registration
// path raises the chunk's flag there).  `slot` (0 or 1) names the staging area: two uploads of one call use different slots.
int kicp_h2d_groups(kicp_ctx *c, int slot, void *dst, const void *src, const size_t *bounds, int ngroups, cudaStream_t stream,
                    int (*after_group)(void *, int), void *user) {
    if (!c || ngroups < 0 || slot < 0 || slot > 1) return KICP_ERR_INVALID;
    if (ngroups == 0) return KICP_OK;
    unsigned char *d = static_cast<unsigned char *>(dst);
    const unsigned char *s = static_cast<const unsigned char *>(src);
    const size_t lo0 = bounds[0], total = bounds[ngroups] - bounds[0];
    const bool staged = c->upload_threads > 0 && total >= kMinStaged && is_pageable(s + lo0);
    if (!staged) {
        for (int g = 0; g < ngroups; ++g) {
            if (bounds[g + 1] > bounds[g])
                KICP_CUDA(cudaMemcpyAsync(d + bounds[g], s + bounds[g], bounds[g + 1] - bounds[g], cudaMemcpyHostToDevice, stream));
            if (after_group) KICP_TRY(after_group(user, g));
        }
        return KICP_OK;
    }
    KICP_TRY(kicp_h2d_prepare(c, slot, s + lo0, total));
    UploadScratch *u = static_cast<UploadScratch *>(c->upload);
    unsigned char *h = u->h_stage[slot];
    auto job = u->stager->start(s + lo0, h, total, kGranule);
    int status = KICP_OK;
    for (int g = 0; g < ngroups && status == KICP_OK; ++g) {
        const size_t a = bounds[g] - lo0, b = bounds[g + 1] - lo0;
        u->stager->wait_prefix(*job, b);
        if (b > a) {
            const cudaError_t e = cudaMemcpyAsync(d + bounds[g], h + a, b - a, cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) status = kicp_cuda_fail(e, "cudaMemcpyAsync (staged upload)", __FILE__, __LINE__);
        }
        if (status == KICP_OK && after_group) status = after_group(user, g);
    }
    u->stager->wait_prefix(*job, total);  // the helpers are done with the caller's buffer before the call returns
    if (cudaEventRecord(u->ev[slot], stream) == cudaSuccess) u->pending[slot] = true;
    return status;
}

// Plain upload of `bytes` bytes, cut into pieces so that the device copy of one piece overlaps the staging of the next.
int kicp_h2d(kicp_ctx *c, int slot, void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return KICP_OK;
    std::vector<size_t> bounds;
    for (size_t off = 0; off < bytes; off += kPiece) bounds.push_back(off);
    bounds.push_back(bytes);
    return kicp_h2d_groups(c, slot, dst, src, bounds.data(), (int)bounds.size() - 1, stream, nullptr, nullptr);
}
