// Multi-GPU plumbing for the sharded MSM: one process (one context) per GPU, NCCL over NVLink for the one exchange the
// path has -- an all-gather of each rank's W window sums (W x 192 B) after its local bucket method.  NCCL cannot reduce
// group elements, so "allreduce of partials" is all-gather + local fold.  NCCL is bound at run time
// (dlopen of libnccl.so.2: the copy the host process already loaded, e.g. torch's, or the system one), so the library
// has no link-time dependency and single-GPU hosts never touch it.
#include <dlfcn.h>

#include "comm.cuh"

namespace zk {

struct NcclUniqueId {
    char internal[128];
};
using ncclComm_t = void*;
struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static NcclApi g_nccl;

static bool load_nccl(std::string& err) {
    if (g_nccl.ok) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return false;
    }
    g_nccl.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclBroadcast");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.Broadcast || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd ||
        !g_nccl.CommDestroy) {
        err = "libnccl.so.2 lacks a required symbol";
        return false;
    }
    g_nccl.ok = true;
    return true;
}
static std::string nccl_err(int rc) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : std::to_string(rc); }

int comm_unique_id(uint8_t out[128], std::string& err) {
    if (!load_nccl(err)) return ZK_ERR_STATE;
    NcclUniqueId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) {
        err = "ncclGetUniqueId: " + nccl_err(rc);
        return ZK_ERR_STATE;
    }
    memcpy(out, id.internal, 128);
    return ZK_OK;
}

int comm_init(zkaes_ctx* ctx, int rank, int nranks, const uint8_t unique_id[128]) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, ZK_ERR_ARG, "comm_init: bad rank / nranks");
    if (ctx->nccl_comm) return fail(ctx, ZK_ERR_STATE, "comm_init: communicator already initialised");
    ctx->rank = rank;
    ctx->nranks = nranks;
    if (nranks == 1) return ZK_OK;
    std::string err;
    if (!load_nccl(err)) return fail(ctx, ZK_ERR_STATE, err);
    NcclUniqueId id;
    memcpy(id.internal, unique_id, 128);
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclComm_t comm = nullptr;
    int rc = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (rc != 0) return fail(ctx, ZK_ERR_STATE, "ncclCommInitRank: " + nccl_err(rc));
    ctx->nccl_comm = comm;
    return ZK_OK;
}

void comm_destroy(zkaes_ctx* ctx) {
    if (ctx->nccl_comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

int comm_all_gather(zkaes_ctx* ctx, const void* send_dev, void* recv_dev, size_t bytes_per_rank) {
    if (ctx->nranks == 1) {
        ZK_CUDA(ctx, cudaMemcpyAsync(recv_dev, send_dev, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return ZK_OK;
    }
    if (!ctx->nccl_comm) return fail(ctx, ZK_ERR_STATE, "all_gather: communicator not initialised");
    int rc = g_nccl.AllGather(send_dev, recv_dev, bytes_per_rank, /*ncclInt8*/ 0, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (rc != 0) return fail(ctx, ZK_ERR_STATE, "ncclAllGather: " + nccl_err(rc));
    return ZK_OK;
}

// In-place broadcast of `bytes` at `buf` from `root` to every rank.
int comm_broadcast(zkaes_ctx* ctx, void* buf, size_t bytes, int root) {
    if (ctx->nranks == 1) return ZK_OK;
    if (!ctx->nccl_comm) return fail(ctx, ZK_ERR_STATE, "broadcast: communicator not initialised");
    int rc = g_nccl.Broadcast(buf, buf, bytes, /*ncclInt8*/ 0, root, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (rc != 0) return fail(ctx, ZK_ERR_STATE, "ncclBroadcast: " + nccl_err(rc));
    return ZK_OK;
}

int comm_group_start(zkaes_ctx* ctx) {
    if (!ctx->nccl_comm) return fail(ctx, ZK_ERR_STATE, "group: communicator not initialised");
    int rc = g_nccl.GroupStart();
    if (rc != 0) return fail(ctx, ZK_ERR_STATE, "ncclGroupStart: " + nccl_err(rc));
    return ZK_OK;
}
int comm_group_end(zkaes_ctx* ctx) {
    int rc = g_nccl.GroupEnd();
    if (rc != 0) return fail(ctx, ZK_ERR_STATE, "ncclGroupEnd: " + nccl_err(rc));
    return ZK_OK;
}
int comm_send(zkaes_ctx* ctx, const void* buf, size_t bytes, int peer) {
    if (!ctx->nccl_comm) return fail(ctx, ZK_ERR_STATE, "send: communicator not initialised");
    int rc = g_nccl.Send(buf, bytes, /*ncclInt8*/ 0, peer, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (rc != 0) return fail(ctx, ZK_ERR_STATE, "ncclSend: " + nccl_err(rc));
    return ZK_OK;
}
int comm_recv(zkaes_ctx* ctx, void* buf, size_t bytes, int peer) {
    if (!ctx->nccl_comm) return fail(ctx, ZK_ERR_STATE, "recv: communicator not initialised");
    int rc = g_nccl.Recv(buf, bytes, /*ncclInt8*/ 0, peer, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (rc != 0) return fail(ctx, ZK_ERR_STATE, "ncclRecv: " + nccl_err(rc));
    return ZK_OK;
}

// Contiguous point range of `rank`: the first (n mod nranks) ranks get one extra element.
void shard_range(size_t n, int rank, int nranks, size_t* start, size_t* count) {
    size_t base = n / (size_t)nranks, extra = n % (size_t)nranks;
    *start = base * (size_t)rank + ((size_t)rank < extra ? (size_t)rank : extra);
    *count = base + ((size_t)rank < extra ? 1 : 0);
}

}  // namespace zk
