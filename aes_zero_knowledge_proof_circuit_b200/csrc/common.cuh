// Shared host-side plumbing for libzkaes_b200: context, error reporting, stream-ordered scratch memory.
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#define ZK_OK 0
#define ZK_ERR_ARG (-1)
#define ZK_ERR_CUDA (-2)
#define ZK_ERR_STATE (-3)
#define ZK_ERR_UNSUPPORTED (-4)

struct zkaes_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;  // kernels launched by this library on this context (bench.py's gpu_launches)
    // cached device tables keyed by (curve, kind, log size)
    std::map<uint64_t, void*> tables;
    // tuning knobs (0 = automatic)
    int msm_window_bits = 0;
    int msm_pair_round = 0;  // 1 = batched-affine pair round before the XYZZ accumulation (msm_core.cuh); off by default: the two extra gather passes cost what the cheaper additions save (profiles/r1_launches_msm_2p26_pair_round.txt)
    int msm_acc_blocks = 3;  // resident blocks per SM of the bucket accumulation kernel (3 or 4)
    int msm_window_max = 22;  // cap of the automatic window choice: bounds the bucket array (2^(c-1) W points of 192 B)
    // multi-GPU: this process' rank among the contexts that share one sharded MSM (comm.cu)
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;
    // optional per-kernel timing of the dominant kernel (bench.py's roofline): CUDA events around every bucket
    // accumulation launch, resolved by zkaes_ctx_profile_read
    bool prof = false;
    struct ProfSpan {
        cudaEvent_t e0, e1;
        uint64_t terms, madds;
    };
    std::vector<ProfSpan> prof_spans;
};

namespace zk {

inline int fail(zkaes_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define ZK_CUDA(ctx, call)                                                                              \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return zk::fail((ctx), ZK_ERR_CUDA,                                                         \
                            std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                                std::to_string(__LINE__) + ")");                                        \
    } while (0)

#define ZK_TRY(expr)             \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != ZK_OK) return rc__; \
    } while (0)

// stream-ordered temporary: freed (stream-ordered) when it leaves scope
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    // host time spent inside the stream-ordered allocator (ZKAES_TRACE prints it per phase: the pool remaps physical memory
    // when a large request does not fit a cached block, which is host-synchronous work)
    static double& alloc_seconds() {
        static double t = 0;
        return t;
    }
    cudaError_t alloc(size_t n, cudaStream_t stream) {
        release();
        s = stream;
        bytes = n;
        if (n == 0) return cudaSuccess;
        auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = cudaMallocAsync(&p, n, stream);
        alloc_seconds() += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return e;
    }
    void release() {
        if (p) {
            auto t0 = std::chrono::steady_clock::now();
            cudaFreeAsync(p, s);
            alloc_seconds() += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        p = nullptr;
    }
    ~DevBuf() { release(); }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace zk
