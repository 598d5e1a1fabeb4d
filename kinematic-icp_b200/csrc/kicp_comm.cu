// Multi-GPU plumbing: one process per GPU, one NCCL communicator per context, one sum-allreduce of the 8
// accumulated doubles per IRLS iteration.  NCCL is resolved lazily with dlopen so that a
// single-GPU user of libkicp_b200.so does not need libnccl at all; under PyTorch the already-loaded
// libnccl.so.2 is the one that gets picked up.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>

#include "kicp_internal.h"

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclFloat64 = 8 };

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    bool ok = false;
};

NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
        }
    }
    return api;
}

int nccl_fail(ncclResult_t r, const char *what) {
    char buf[256];
    snprintf(buf, sizeof(buf), "NCCL error %d (%s) in %s", (int)r, nccl().ok ? nccl().GetErrorString(r) : "?", what);
    kicp_set_error(buf);
    return KICP_ERR_NCCL;
}
}  // namespace

extern "C" int kicp_comm_unique_id(uint8_t id[KICP_UNIQUE_ID_BYTES]) {
    if (!id) return KICP_ERR_INVALID;
    if (!nccl().ok) {
        kicp_set_error("libnccl.so.2 could not be loaded");
        return KICP_ERR_NCCL;
    }
    ncclUniqueId uid;
    ncclResult_t r = nccl().GetUniqueId(&uid);
    if (r != 0) return nccl_fail(r, "ncclGetUniqueId");
    static_assert(sizeof(uid) == KICP_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &uid, sizeof(uid));
    return KICP_OK;
}

// NCCL and the peer-memory mailboxes are torn down separately: (re-)initialising one never invalidates the other.
static void nccl_teardown(kicp_ctx *ctx) {
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
}
static void p2p_close_peers(kicp_ctx *ctx) {
    if (!ctx->p2p_ready) return;
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < KICP_MAX_RANKS; ++r) {
        if (r != ctx->rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
        ctx->p2p_peer[r] = nullptr;
    }
    ctx->p2p_ready = false;
}

extern "C" int kicp_comm_init(kicp_ctx *ctx, const uint8_t id[KICP_UNIQUE_ID_BYTES], int32_t nranks, int32_t rank) {
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return KICP_ERR_INVALID;
    if (ctx->p2p_ready && (ctx->nranks != nranks || ctx->rank != rank)) {
        kicp_set_error("kicp_comm_init: rank layout differs from the one given to kicp_comm_p2p_init");
        return KICP_ERR_INVALID;
    }
    if (!nccl().ok) {
        kicp_set_error("libnccl.so.2 could not be loaded");
        return KICP_ERR_NCCL;
    }
    KICP_CUDA(cudaSetDevice(ctx->device));
    nccl_teardown(ctx);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    ncclResult_t r = nccl().CommInitRank(&comm, nranks, uid, rank);
    if (r != 0) return nccl_fail(r, "ncclCommInitRank");
    ctx->nccl_comm = comm;
    ctx->nranks = nranks, ctx->rank = rank;
    return KICP_OK;
}

extern "C" int kicp_comm_destroy(kicp_ctx *ctx) {
    if (!ctx) return KICP_ERR_INVALID;
    nccl_teardown(ctx);
    p2p_close_peers(ctx);
    if (ctx->p2p_local) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->p2p_local);
        ctx->p2p_local = nullptr;
    }
    ctx->nranks = 1, ctx->rank = 0;
    return KICP_OK;
}

// in-place sum of 8 doubles across the communicator, enqueued on the context stream
int kicp_comm_allreduce8(kicp_ctx *ctx, double *d_buf) {
    if (!ctx->nccl_comm) return KICP_ERR_INVALID;
    ncclResult_t r = nccl().AllReduce(d_buf, d_buf, 8, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != 0) return nccl_fail(r, "ncclAllReduce");
    ctx->launches++;
    return KICP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused peer-memory exchange: mailbox allocation, CUDA-IPC export / import.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int kicp_comm_p2p_handle(kicp_ctx *ctx, uint8_t handle[KICP_IPC_HANDLE_BYTES]) {
    if (!ctx || !handle) return KICP_ERR_INVALID;
    KICP_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->p2p_local) {
        KICP_CUDA(cudaMalloc(&ctx->p2p_local, sizeof(P2PMailbox)));
        KICP_CUDA(cudaMemset(ctx->p2p_local, 0, sizeof(P2PMailbox)));
        KICP_CUDA(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    KICP_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p_local));
    static_assert(sizeof(h) == KICP_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(handle, &h, sizeof(h));
    return KICP_OK;
}

extern "C" int kicp_comm_p2p_init(kicp_ctx *ctx, const uint8_t *handles, int32_t nranks, int32_t rank) {
    if (!ctx || !handles || nranks < 1 || nranks > KICP_MAX_RANKS || rank < 0 || rank >= nranks || !ctx->p2p_local)
        return KICP_ERR_INVALID;
    if (ctx->nccl_comm && (ctx->nranks != nranks || ctx->rank != rank)) {
        kicp_set_error("kicp_comm_p2p_init: rank layout differs from the one given to kicp_comm_init");
        return KICP_ERR_INVALID;
    }
    KICP_CUDA(cudaSetDevice(ctx->device));
    p2p_close_peers(ctx);  // a second init must not leak the handles of the first
    ctx->rank = rank;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            ctx->p2p_peer[r] = ctx->p2p_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * KICP_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        KICP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->p2p_peer[r] = (P2PMailbox *)p;
    }
    ctx->nranks = nranks;
    // p2p_seq is NOT reset: the words left in the mailboxes carry tags of earlier registrations, and tags only grow
    ctx->p2p_ready = true;
    return KICP_OK;
}
