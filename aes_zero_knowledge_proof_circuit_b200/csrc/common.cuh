// Shared host-side plumbing for libzkaes_b200: context, error reporting, stream-ordered scratch memory.
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <iterator>
#include <cstdint>
#include <cstdio>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#define ZK_OK 0
#define ZK_ERR_ARG (-1)
#define ZK_ERR_CUDA (-2)
#define ZK_ERR_STATE (-3)
#define ZK_ERR_UNSUPPORTED (-4)

// Host-managed arena for the prover's large temporaries.  The stream-ordered pool (cudaMallocAsync) reuses cached
// physical memory by REMAPPING it when a request does not match a cached block; with the prover's mix of multi-GB buffers
// that is host-synchronous driver work whose amount depends on the pool's fragmentation state: identical runs of the 256-byte
// proof took 1.28 s or 2.12 s with identical kernel times (profiles/r1_alloc_bimodal_256B.txt).  After the first proof has
// shown the peak of live scratch, the context carves one cudaMalloc'ed arena of that size (+15 %) and DevBuf serves from it
// with a best-fit free list on the host -- no driver calls, the same addresses every proof.  Everything runs on the
// context's single stream, so a freed block may be handed out again at once (stream order = program order), exactly the
// guarantee cudaFreeAsync gives.  A request that does not fit falls back to the pool.
struct DevArena {
    char* base = nullptr;
    size_t size = 0, live = 0, high = 0;
    std::map<size_t, size_t> free_segs;  // offset -> length, coalesced
    static size_t round_up(size_t n) { return (n + 511) & ~(size_t)511; }
    void reset(char* b, size_t s) {
        base = b;
        size = s;
        live = high = 0;
        free_segs.clear();
        if (s) free_segs[0] = s;
    }
    void* alloc(size_t n) {
        n = round_up(n);
        auto best = free_segs.end();
        for (auto it = free_segs.begin(); it != free_segs.end(); ++it)
            if (it->second >= n && (best == free_segs.end() || it->second < best->second)) best = it;
        if (best == free_segs.end()) return nullptr;
        const size_t off = best->first, len = best->second;
        free_segs.erase(best);
        if (len > n) free_segs[off + n] = len - n;
        live += n;
        if (live > high) high = live;
        return base + off;
    }
    bool owns(const void* p) const { return base && (const char*)p >= base && (const char*)p < base + size; }
    void free(void* p, size_t n) {
        n = round_up(n);
        size_t off = (size_t)((char*)p - base);
        live -= n;
        auto next = free_segs.lower_bound(off);
        if (next != free_segs.begin()) {
            auto prev = std::prev(next);
            if (prev->first + prev->second == off) {
                off = prev->first;
                n += prev->second;
                free_segs.erase(prev);
            }
        }
        if (next != free_segs.end() && off + n == next->first) {
            n += next->second;
            free_segs.erase(next);
        }
        free_segs[off] = n;
    }
};

// Worker of a single-process multi-GPU context (zkaes_ctx_create_multi): one host thread per peer rank, bound to that rank's
// device, executing the jobs the leader posts (key synthesis, encrypt(), key files) in lock step with the leader's own call.
struct ZkWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
};

struct zkaes_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // the context's OWN stream-ordered memory pool: every DevBuf of this context comes from it (release threshold, trimming and the
    // high-water mark that sizes the scratch arena then concern this context only, not other cudaMallocAsync users of the process)
    cudaMemPool_t pool = nullptr;
    bool pool_owned = false;
    std::string err;
    uint64_t launches = 0;  // kernels launched by this library on this context (bench.py's gpu_launches)
    // cached device tables keyed by (curve, kind, log size)
    std::map<uint64_t, void*> tables;
    // tuning knobs (0 = automatic)
    int msm_window_bits = 0;
    int msm_pair_round = 0;  // R = number of batched-affine pair rounds before the XYZZ accumulation (msm_core.cuh); 0 = plain accumulation.
                             // Off by default: one round measured break-even (profiles/r1_launches_msm_2p26_pair_round.txt)
    int msm_acc_blocks = 3;  // resident blocks per SM of the bucket accumulation kernel (3 or 4)
    int msm_plan_ranks = 1;  // stand-alone sharded MSM entry points (zkaes_msm_g1_windows / _fold): ranks sharing the MSM, so the window plan fits the per-rank share
    int msm_prefetch = 0;    // 1 / 2: stage the next entry's point in shared memory (cp.async / cp.async.bulk + mbarrier) while the current one is added
    int msm_madd_call = 1;   // 1: the mixed addition issues its ten products through one out-of-line multiplier (XYZZ::madd_call)
    int r1_lagrange = 1;     // 1: encrypt() commits to w, z_A, z_B in the Lagrange basis when the key holds those points (prover.cu, pk_build_lagrange)
    int msm_window_max = 23;  // cap of the automatic window choice: bounds the bucket array (2^(c-1) W points of 192 B: 8.9 GB at c = 23, W = 11)
    // multi-GPU: this context's rank among the contexts that share one sharded MSM (comm.cu) -- one process per GPU
    // (zkaes_ctx_comm_init), or one process driving all GPUs (zkaes_ctx_create_multi: the leader, rank 0, owns the peers)
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;
    std::vector<zkaes_ctx*> peers;                  // leader only: ranks 1..nranks-1
    std::vector<std::unique_ptr<ZkWorker>> workers;  // leader only: workers[i] drives peers[i]
    // optional per-kernel timing of the dominant kernel (bench.py's roofline): CUDA events around every bucket
    // accumulation launch, resolved by zkaes_ctx_profile_read
    bool prof = false;
    struct ProfSpan {
        cudaEvent_t e0, e1;
        uint64_t terms, madds;
    };
    std::vector<ProfSpan> prof_spans;
    // scratch arena (see DevArena): created by the second encrypt() on this context from the first one's measured peak
    DevArena arena;
    uint64_t scratch_peak = 0;   // pool high-water mark of the last encrypt() that ran without the arena
    int arena_state = 0;         // 0 = not tried yet, 1 = active, -1 = disabled / allocation failed
    uint64_t arena_domain = 0;   // |H| of the key the peak was measured with: a key of another size releases the arena and measures again
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
        static thread_local double t = 0;
        return t;
    }
    // the pool of the context that is working on THIS host thread (set at every C-ABI entry point); null = the device's default pool
    static cudaMemPool_t& pool() {
        static thread_local cudaMemPool_t p = nullptr;
        return p;
    }
    // the arena of the context that is proving on THIS host thread (one context per thread); null = stream-ordered pool only
    static DevArena*& arena() {
        static thread_local DevArena* a = nullptr;
        return a;
    }
    cudaError_t alloc(size_t n, cudaStream_t stream) {
        release();
        s = stream;
        bytes = n;
        if (n == 0) return cudaSuccess;
        if (DevArena* a = arena()) {
            if ((p = a->alloc(n)) != nullptr) return cudaSuccess;
        }
        auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = pool() ? cudaMallocFromPoolAsync(&p, n, pool(), stream) : cudaMallocAsync(&p, n, stream);
        alloc_seconds() += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return e;
    }
    void release() {
        if (p) {
            DevArena* a = arena();
            if (a && a->owns(p)) {
                a->free(p, bytes);
            } else {
                auto t0 = std::chrono::steady_clock::now();
                cudaFreeAsync(p, s);
                alloc_seconds() += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            }
        }
        p = nullptr;
    }
    ~DevBuf() { release(); }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace zk
