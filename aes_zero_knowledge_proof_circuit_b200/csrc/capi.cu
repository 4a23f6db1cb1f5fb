// C ABI of libzkaes_b200 (declared in include/zkaes_b200.h).  Thin: argument checking, curve dispatch,
// host<->device staging.  No CPU fallback anywhere: without a CUDA device zkaes_ctx_create fails.
#include "../../include/zkaes_b200.h"
#include "common.cuh"
#include "msm.cuh"
#include "ntt.cuh"
#include "srs.cuh"
#include "circuit.h"
#include "witness.cuh"
#include "prover.cuh"
#include "comm.cuh"
#include "verifier.h"
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace zk {
template <class F> int selftest_field_asm(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template <class F> int selftest_field_portable(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template <class C> int selftest_g1(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template <class F> int selftest_field_call(zkaes_ctx*, int, const void*, const void*, void*, size_t);
int selftest_fq52(zkaes_ctx*, const void*, const void*, void*, size_t);
}
using namespace zk;

// context-free entry points (the host verifier, the proof wire format) and calls made without a context report through a
// per-thread string, read with zkaes_last_error(NULL)
static thread_local std::string g_host_err;
#define NEED_CTX(ctx)                   \
    if (!(ctx)) {                       \
        g_host_err = "null context";    \
        return ZK_ERR_ARG;              \
    }                                   \
    zk::DevBuf::pool() = (ctx)->pool
#define CURVE_DISPATCH(ctx, curve_id, expr377, expr381)                    \
    ((curve_id) == 377 ? (expr377) : (curve_id) == 381 ? (expr381) : fail((ctx), ZK_ERR_ARG, "unknown curve_id (use 377 or 381)"))

// ---- MSM ---------------------------------------------------------------------------------------------------
template <class C>
static int msm_windows_impl(zkaes_ctx* ctx, const void* bases, const void* scalars, size_t n_local, size_t n_total, int mont,
                            void* windows_dev) {
    MsmPlan p = msm_make_plan(n_total ? n_total : 1, C::FrP::BITS, ctx->msm_window_bits, ctx->msm_plan_ranks, ctx->msm_window_max);
    return msm_window_sums<C>(ctx, bases, scalars, n_local, mont & 1, p, windows_dev, (mont >> 1) & 1);
}
template <class C>
static int msm_fold_impl(zkaes_ctx* ctx, const void* gathered, int n_ranks, size_t n_total, void* out96) {
    MsmPlan p = msm_make_plan(n_total ? n_total : 1, C::FrP::BITS, ctx->msm_window_bits, ctx->msm_plan_ranks, ctx->msm_window_max);
    std::vector<XYZZ<C>> h((size_t)n_ranks * p.W);
    ZK_CUDA(ctx, cudaMemcpyAsync(h.data(), gathered, sizeof(XYZZ<C>) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    Affine<C> r = msm_fold_windows_host<C>(h.data(), n_ranks, p);
    memcpy(out96, &r, 96);
    return ZK_OK;
}
template <class C>
static int msm_device_impl(zkaes_ctx* ctx, const void* bases, const void* scalars, size_t n, int mont, void* out96) {
    MsmPlan p = msm_make_plan(n ? n : 1, C::FrP::BITS, ctx->msm_window_bits, 1, ctx->msm_window_max);
    DevBuf win;
    ZK_CUDA(ctx, win.alloc(sizeof(XYZZ<C>) * p.W, ctx->stream));
    ZK_TRY(msm_window_sums<C>(ctx, bases, scalars, n, mont & 1, p, win.p, (mont >> 1) & 1));
    return msm_fold_impl<C>(ctx, win.p, 1, n, out96);
}
template <class C>
static int msm_host_impl(zkaes_ctx* ctx, const void* bases, const void* scalars, size_t n, void* out96) {
    DevBuf db, ds;
    ZK_CUDA(ctx, db.alloc(96 * n, ctx->stream));
    ZK_CUDA(ctx, ds.alloc(32 * n, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(db.p, bases, 96 * n, cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(ds.p, scalars, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
    return msm_device_impl<C>(ctx, db.p, ds.p, n, 0, out96);
}

// small signed scalars (host buffers): the Lagrange-basis commitments' kernel path (msm.cu, "small-scalar MSM") on its own
static constexpr int MSM_SMALL_WINDOWS = 32;
template <class C>
static int msm_small_host_impl(zkaes_ctx* ctx, const void* bases, const int32_t* vals, size_t n, int c, void* out96) {
    DevBuf db, dv, win;
    cudaStream_t st = ctx->stream;
    ZK_CUDA(ctx, db.alloc(96 * n, st));
    ZK_CUDA(ctx, dv.alloc(4 * n, st));
    ZK_CUDA(ctx, win.alloc(sizeof(XYZZ<C>) * MSM_SMALL_WINDOWS, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(db.p, bases, 96 * n, cudaMemcpyHostToDevice, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(dv.p, vals, 4 * n, cudaMemcpyHostToDevice, st));
    ZK_TRY(msm_small_window_sums<C>(ctx, db.p, dv.as<int32_t>(), n, 0, 1, c, MSM_SMALL_WINDOWS, win.p));
    std::vector<XYZZ<C>> h(MSM_SMALL_WINDOWS);
    ZK_CUDA(ctx, cudaMemcpyAsync(h.data(), win.p, sizeof(XYZZ<C>) * MSM_SMALL_WINDOWS, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    Affine<C> r = msm_sum_windows_host<C>(h.data(), h.size());
    memcpy(out96, &r, 96);
    return ZK_OK;
}

// ---- single-process multi-GPU (zkaes_ctx_create_multi) ------------------------------------------------------------------------
// The opaque key handle: the leader's key plus, for a multi-GPU context, one key per peer rank (each holds its rank's SRS share).
struct zkaes_pk {
    zk::zkaes_pk_impl* impl = nullptr;
    std::vector<zk::zkaes_pk_impl*> peer_impls;
};

static void worker_loop(ZkWorker* w, int device) {
    cudaSetDevice(device);
    std::unique_lock<std::mutex> lk(w->m);
    for (;;) {
        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
        if (w->quit) return;
        std::function<int()> job = std::move(w->job);
        w->has_job = false;
        lk.unlock();
        int rc;
        try {
            rc = job();
        } catch (const std::exception&) {
            rc = ZK_ERR_STATE;
        }
        lk.lock();
        w->rc = rc;
        w->done = true;
        w->cv.notify_all();
    }
}
// fn(context of rank i, i) on every rank at once: peers on their worker threads, rank 0 on the calling thread.  Returns the first
// failure (the failing rank's message is copied into the leader's last_error).
static int run_on_all(zkaes_ctx* ctx, const std::function<int(zkaes_ctx*, int)>& fn) {
    cudaSetDevice(ctx->device);
    for (size_t i = 0; i < ctx->peers.size(); ++i) {
        ZkWorker* w = ctx->workers[i].get();
        zkaes_ctx* pc = ctx->peers[i];
        std::lock_guard<std::mutex> g(w->m);
        w->job = [&fn, pc, i] {
            zk::DevBuf::pool() = pc->pool;
            return fn(pc, (int)i + 1);
        };
        w->has_job = true;
        w->done = false;
        w->cv.notify_all();
    }
    int rc;
    zk::DevBuf::pool() = ctx->pool;
    try {
        rc = fn(ctx, 0);
    } catch (const std::exception& e) {
        rc = fail(ctx, ZK_ERR_STATE, e.what());
    }
    for (size_t i = 0; i < ctx->peers.size(); ++i) {
        ZkWorker* w = ctx->workers[i].get();
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc != ZK_OK && rc == ZK_OK) rc = fail(ctx, w->rc, "rank " + std::to_string(i + 1) + ": " + ctx->peers[i]->err);
    }
    return rc;
}
static std::string rank_path(const zkaes_ctx* ctx, const char* path, int rank) {
    return ctx->peers.empty() ? std::string(path) : std::string(path) + ".r" + std::to_string(rank);
}

extern "C" {

int zkaes_ctx_create(int device_id, zkaes_ctx** out) {
    if (!out) return ZK_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device_id < 0 || device_id >= count) return ZK_ERR_CUDA;
    if (cudaSetDevice(device_id) != cudaSuccess) return ZK_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return ZK_ERR_CUDA;
    if (prop.major != 10) return ZK_ERR_UNSUPPORTED;  // sm_100a cubins only
    zkaes_ctx* c = new zkaes_ctx();
    c->device = device_id;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return ZK_ERR_CUDA;
    }
    if (const char* e = getenv("ZKAES_MSM_WINDOW_MAX")) {  // experiments: cap of the automatic MSM window choice (same as the tuning knob)
        const int v = atoi(e);
        if (v >= 3 && v <= 24) c->msm_window_max = v;
    }
    // a private stream-ordered pool that keeps freed scratch cached instead of returning it to the driver every call (falls back to the
    // device's default pool, untouched, if the driver refuses to create one)
    {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device_id;
        if (cudaMemPoolCreate(&c->pool, &props) == cudaSuccess) {
            c->pool_owned = true;
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr);
        } else {
            cudaGetLastError();
            c->pool = nullptr;
        }
    }
    *out = c;
    return ZK_OK;
}
int zkaes_ctx_create_multi(const int* device_ids, int n_devices, zkaes_ctx** out) {
    if (!out) return ZK_ERR_ARG;
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 64) return ZK_ERR_ARG;
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return ZK_ERR_ARG;
    zkaes_ctx* leader = nullptr;
    int rc = zkaes_ctx_create(device_ids[0], &leader);
    if (rc != ZK_OK) return rc;
    if (n_devices == 1) {
        *out = leader;
        return ZK_OK;
    }
    // peers are created ON their worker threads, so each thread's current device is its rank's
    for (int i = 1; i < n_devices && rc == ZK_OK; ++i) {
        leader->workers.emplace_back(new ZkWorker());
        leader->peers.push_back(nullptr);
        ZkWorker* w = leader->workers.back().get();
        w->th = std::thread(worker_loop, w, device_ids[i]);
    }
    uint8_t uid[128];
    std::string err;
    rc = zk::comm_unique_id(uid, err);
    if (rc != ZK_OK) fail(leader, rc, err);
    if (rc == ZK_OK) {
        std::vector<int> devs(device_ids, device_ids + n_devices);
        // create the peer contexts and join the communicator: ncclCommInitRank blocks until every rank has called it
        std::vector<zkaes_ctx*>& peers = leader->peers;
        for (size_t i = 0; i < peers.size(); ++i) {
            ZkWorker* w = leader->workers[i].get();
            std::lock_guard<std::mutex> g(w->m);
            w->job = [&peers, &devs, &uid, i, n_devices] {
                int r = zkaes_ctx_create(devs[i + 1], &peers[i]);
                if (r != ZK_OK) return r;
                return zk::comm_init(peers[i], (int)i + 1, n_devices, uid);
            };
            w->has_job = true;
            w->done = false;
            w->cv.notify_all();
        }
        cudaSetDevice(leader->device);
        rc = zk::comm_init(leader, 0, n_devices, uid);
        for (size_t i = 0; i < peers.size(); ++i) {
            ZkWorker* w = leader->workers[i].get();
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->done; });
            if (w->rc != ZK_OK && rc == ZK_OK) rc = w->rc;
        }
    }
    if (rc != ZK_OK) {
        zkaes_ctx_destroy(leader);
        return rc;
    }
    *out = leader;
    return ZK_OK;
}
int zkaes_ctx_devices(const zkaes_ctx* ctx) { return ctx ? ctx->nranks : 0; }
void zkaes_ctx_destroy(zkaes_ctx* ctx) {
    if (!ctx) return;
    for (size_t i = 0; i < ctx->workers.size(); ++i) {  // peers first, each on its own thread
        ZkWorker* w = ctx->workers[i].get();
        zkaes_ctx* pc = ctx->peers[i];
        {
            std::lock_guard<std::mutex> g(w->m);
            w->job = [pc] {
                if (pc) zkaes_ctx_destroy(pc);
                return ZK_OK;
            };
            w->has_job = true;
            w->done = false;
            w->cv.notify_all();
        }
        {
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->done; });
            w->quit = true;
            w->cv.notify_all();
        }
        if (w->th.joinable()) w->th.join();
    }
    ctx->workers.clear();
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    zk::comm_destroy(ctx);
    for (auto& kv : ctx->tables) cudaFree(kv.second);
    if (ctx->arena.base) cudaFree(ctx->arena.base);
    if (ctx->pool_owned) cudaMemPoolDestroy(ctx->pool);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
const char* zkaes_last_error(const zkaes_ctx* ctx) { return ctx ? ctx->err.c_str() : g_host_err.empty() ? "null context" : g_host_err.c_str(); }
void* zkaes_ctx_stream(zkaes_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t zkaes_ctx_launches(const zkaes_ctx* ctx) {
    if (!ctx) return 0;
    uint64_t n = ctx->launches;
    for (const zkaes_ctx* p : ctx->peers) n += p ? p->launches : 0;
    return n;
}
int zkaes_ctx_sync(zkaes_ctx* ctx) {
    NEED_CTX(ctx);
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
int zkaes_comm_unique_id(uint8_t out128[128]) {
    if (!out128) return ZK_ERR_ARG;
    std::string err;
    return zk::comm_unique_id(out128, err);
}
int zkaes_ctx_comm_init(zkaes_ctx* ctx, int rank, int nranks, const uint8_t unique_id128[128]) {
    NEED_CTX(ctx);
    if (nranks > 1 && !unique_id128) return fail(ctx, ZK_ERR_ARG, "comm_init: null unique id");
    return zk::comm_init(ctx, rank, nranks, unique_id128);
}
int zkaes_shard_range(size_t n, int rank, int nranks, size_t* start, size_t* count) {
    if (!start || !count || nranks < 1 || rank < 0 || rank >= nranks) return ZK_ERR_ARG;
    zk::shard_range(n, rank, nranks, start, count);
    return ZK_OK;
}
int zkaes_coset_plan(int nranks, int ncoset, int ntask, double own_extra, int* owner_out, int* exec_out) {
    if (nranks < 1 || ncoset < 1 || ntask < 1 || ncoset > 64 || ntask > 64 || !owner_out || !exec_out) return ZK_ERR_ARG;
    const zk::CosetPlan pl = zk::coset_plan(nranks, ncoset, ntask, own_extra);
    for (int j = 0; j < ncoset; ++j) owner_out[j] = pl.owner[j];
    for (int i = 0; i < ncoset * ntask; ++i) exec_out[i] = pl.exec[i];
    return ZK_OK;
}
int zkaes_ctx_profile(zkaes_ctx* ctx, int enable) {
    NEED_CTX(ctx);
    ctx->prof = enable != 0;
    return ZK_OK;
}
int zkaes_ctx_profile_read(zkaes_ctx* ctx, double out[4]) {
    NEED_CTX(ctx);
    if (!out) return fail(ctx, ZK_ERR_ARG, "profile_read: null pointer");
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0, terms = 0, madds = 0;
    for (auto& sp : ctx->prof_spans) {
        float t = 0;
        cudaEventElapsedTime(&t, sp.e0, sp.e1);
        ms += t;
        terms += (double)sp.terms;
        madds += (double)sp.madds;
        cudaEventDestroy(sp.e0);
        cudaEventDestroy(sp.e1);
    }
    out[0] = (double)ctx->prof_spans.size();
    out[1] = ms;
    out[2] = terms;
    out[3] = madds;
    ctx->prof_spans.clear();
    return ZK_OK;
}
int zkaes_ctx_set_tuning(zkaes_ctx* ctx, const char* key, int value) {
    NEED_CTX(ctx);
    if (!key) return fail(ctx, ZK_ERR_ARG, "tuning: null key");
    const std::string k(key);
    if (k == "msm_window_max") {
        if (value < 3 || value > 24) return fail(ctx, ZK_ERR_ARG, "msm_window_max must be in 3..24");
        ctx->msm_window_max = value;
    } else if (k == "msm_pair_round") {
        if (value < 0 || value > 4) return fail(ctx, ZK_ERR_ARG, "msm_pair_round must be in 0..4");
        ctx->msm_pair_round = value;
    } else if (k == "msm_acc_blocks") {
        if (value != 3 && value != 4) return fail(ctx, ZK_ERR_ARG, "msm_acc_blocks must be 3 or 4");
        ctx->msm_acc_blocks = value;
    } else if (k == "msm_plan_ranks") {
        if (value < 1 || value > 1024) return fail(ctx, ZK_ERR_ARG, "msm_plan_ranks must be in 1..1024");
        ctx->msm_plan_ranks = value;
    } else if (k == "msm_prefetch") {
        if (value < 0 || value > 2) return fail(ctx, ZK_ERR_ARG, "msm_prefetch must be 0, 1 or 2");
        ctx->msm_prefetch = value;
    } else if (k == "r1_lagrange") {
        if (value != 0 && value != 1) return fail(ctx, ZK_ERR_ARG, "r1_lagrange must be 0 or 1");
        ctx->r1_lagrange = value;
    } else if (k == "msm_madd_call") {
        if (value != 0 && value != 1) return fail(ctx, ZK_ERR_ARG, "msm_madd_call must be 0 or 1");
        ctx->msm_madd_call = value;
    } else {
        return fail(ctx, ZK_ERR_ARG, "tuning: unknown key " + k);
    }
    for (zkaes_ctx* p : ctx->peers)  // every rank must derive the same window plan
        if (p) zkaes_ctx_set_tuning(p, key, value);
    return ZK_OK;
}
int zkaes_ctx_set_msm_window(zkaes_ctx* ctx, int window_bits) {
    NEED_CTX(ctx);
    if (window_bits < 0 || window_bits > 24 || window_bits == 1 || window_bits == 2) return fail(ctx, ZK_ERR_ARG, "window bits must be 0 or 3..24");
    ctx->msm_window_bits = window_bits;
    for (zkaes_ctx* p : ctx->peers)
        if (p) p->msm_window_bits = window_bits;
    return ZK_OK;
}

int zkaes_dev_alloc(zkaes_ctx* ctx, size_t bytes, void** out_dev) {
    NEED_CTX(ctx);
    if (!out_dev) return fail(ctx, ZK_ERR_ARG, "null out pointer");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    ZK_CUDA(ctx, cudaMalloc(out_dev, bytes ? bytes : 1));
    return ZK_OK;
}
int zkaes_dev_free(zkaes_ctx* ctx, void* dev) {
    NEED_CTX(ctx);
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZK_CUDA(ctx, cudaFree(dev));
    return ZK_OK;
}
int zkaes_dev_upload(zkaes_ctx* ctx, void* dev, const void* host, size_t bytes) {
    NEED_CTX(ctx);
    ZK_CUDA(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
int zkaes_dev_download(zkaes_ctx* ctx, void* host, const void* dev, size_t bytes) {
    NEED_CTX(ctx);
    ZK_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

// ---- MSM (entry points; templates are defined above the extern block) ----
int zkaes_msm_g1(zkaes_ctx* ctx, int curve_id, const void* bases, const void* scalars, size_t n, void* out96) {
    NEED_CTX(ctx);
    if ((n && (!bases || !scalars)) || !out96) return fail(ctx, ZK_ERR_ARG, "msm: null pointer");
    if (n >= ((size_t)1 << 31)) return fail(ctx, ZK_ERR_ARG, "msm: n must be < 2^31");
    return CURVE_DISPATCH(ctx, curve_id, msm_host_impl<G1_377Params>(ctx, bases, scalars, n, out96),
                          msm_host_impl<G1_381Params>(ctx, bases, scalars, n, out96));
}
int zkaes_msm_g1_device(zkaes_ctx* ctx, int curve_id, const void* bases, const void* scalars, size_t n, int mont, void* out96) {
    NEED_CTX(ctx);
    if ((n && (!bases || !scalars)) || !out96) return fail(ctx, ZK_ERR_ARG, "msm: null pointer");
    if (n >= ((size_t)1 << 31)) return fail(ctx, ZK_ERR_ARG, "msm: n must be < 2^31");
    return CURVE_DISPATCH(ctx, curve_id, msm_device_impl<G1_377Params>(ctx, bases, scalars, n, mont, out96),
                          msm_device_impl<G1_381Params>(ctx, bases, scalars, n, mont, out96));
}
int zkaes_msm_g1_small(zkaes_ctx* ctx, int curve_id, const void* bases, const int32_t* values, size_t n, int value_bits, void* out96) {
    NEED_CTX(ctx);
    if ((n && (!bases || !values)) || !out96) return fail(ctx, ZK_ERR_ARG, "small msm: null pointer");
    if (n >= ((size_t)1 << 31)) return fail(ctx, ZK_ERR_ARG, "small msm: n must be < 2^31");
    if (value_bits < 1 || value_bits > 13) return fail(ctx, ZK_ERR_ARG, "small msm: value_bits must be 1..13");
    const int64_t lim = (int64_t)1 << (value_bits - 1);
    for (size_t i = 0; i < n; ++i)
        if (values[i] > lim || values[i] < -lim) return fail(ctx, ZK_ERR_ARG, "small msm: |value| exceeds 2^(value_bits - 1)");
    return CURVE_DISPATCH(ctx, curve_id, msm_small_host_impl<G1_377Params>(ctx, bases, values, n, value_bits, out96),
                          msm_small_host_impl<G1_381Params>(ctx, bases, values, n, value_bits, out96));
}
int zkaes_msm_g1_prepare_bases(zkaes_ctx* ctx, int curve_id, void* bases_dev, size_t n) {
    NEED_CTX(ctx);
    if (n && !bases_dev) return fail(ctx, ZK_ERR_ARG, "prepare_bases: null pointer");
    return CURVE_DISPATCH(ctx, curve_id, msm_bases_to_internal<G1_377Params>(ctx, bases_dev, bases_dev, n),
                          msm_bases_to_internal<G1_381Params>(ctx, bases_dev, bases_dev, n));
}
size_t zkaes_msm_g1_windows_bytes(zkaes_ctx* ctx, int curve_id, size_t n_total) {
    if (!ctx) return 0;
    int bits = curve_id == 377 ? Fr377Params::BITS : Fr381Params::BITS;
    MsmPlan p = msm_make_plan(n_total ? n_total : 1, bits, ctx->msm_window_bits, ctx->msm_plan_ranks, ctx->msm_window_max);
    return (size_t)p.W * 192;
}
int zkaes_msm_g1_windows(zkaes_ctx* ctx, int curve_id, const void* bases, const void* scalars, size_t n_local, size_t n_total,
                         int mont, void* windows_dev) {
    NEED_CTX(ctx);
    if ((n_local && (!bases || !scalars)) || !windows_dev) return fail(ctx, ZK_ERR_ARG, "msm: null pointer");
    if (n_local >= ((size_t)1 << 31)) return fail(ctx, ZK_ERR_ARG, "msm: n must be < 2^31");
    return CURVE_DISPATCH(ctx, curve_id, msm_windows_impl<G1_377Params>(ctx, bases, scalars, n_local, n_total, mont, windows_dev),
                          msm_windows_impl<G1_381Params>(ctx, bases, scalars, n_local, n_total, mont, windows_dev));
}
int zkaes_msm_g1_fold(zkaes_ctx* ctx, int curve_id, const void* gathered, int n_ranks, size_t n_total, void* out96) {
    NEED_CTX(ctx);
    if (!gathered || !out96 || n_ranks < 1) return fail(ctx, ZK_ERR_ARG, "msm fold: bad arguments");
    return CURVE_DISPATCH(ctx, curve_id, msm_fold_impl<G1_377Params>(ctx, gathered, n_ranks, n_total, out96),
                          msm_fold_impl<G1_381Params>(ctx, gathered, n_ranks, n_total, out96));
}

// ---- NTT ---------------------------------------------------------------------------------------------------
int zkaes_ntt_fr_device(zkaes_ctx* ctx, int curve_id, void* data, uint32_t log_n, int inverse, int coset) {
    NEED_CTX(ctx);
    if (!data) return fail(ctx, ZK_ERR_ARG, "ntt: null pointer");
    return CURVE_DISPATCH(ctx, curve_id, ntt_device<Fr377Params>(ctx, 377, data, (int)log_n, inverse, coset),
                          ntt_device<Fr381Params>(ctx, 381, data, (int)log_n, inverse, coset));
}
int zkaes_ntt_fr(zkaes_ctx* ctx, int curve_id, void* data_host, uint32_t log_n, int inverse, int coset) {
    NEED_CTX(ctx);
    if (!data_host) return fail(ctx, ZK_ERR_ARG, "ntt: null pointer");
    if (log_n > 30) return fail(ctx, ZK_ERR_ARG, "ntt: log_n out of range");
    size_t bytes = (size_t)32 << log_n;
    DevBuf d;
    ZK_CUDA(ctx, d.alloc(bytes, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(d.p, data_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ZK_TRY(zkaes_ntt_fr_device(ctx, curve_id, d.p, log_n, inverse, coset));
    ZK_CUDA(ctx, cudaMemcpyAsync(data_host, d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

// ---- SRS ---------------------------------------------------------------------------------------------------
int zkaes_srs_powers_device(zkaes_ctx* ctx, int curve_id, const uint8_t seed32[32], size_t n, void* out_bases_dev) {
    NEED_CTX(ctx);
    if (!seed32 || (n && !out_bases_dev)) return fail(ctx, ZK_ERR_ARG, "srs: null pointer");
    return CURVE_DISPATCH(ctx, curve_id, srs_powers_device<G1_377Params>(ctx, seed32, n, out_bases_dev),
                          srs_powers_device<G1_381Params>(ctx, seed32, n, out_bases_dev));
}

// ---- self tests ----------------------------------------------------------------------------------------------
int zkaes_selftest_field(zkaes_ctx* ctx, int curve_id, int field, int op, int variant, const void* a, const void* b, void* out,
                         size_t count) {
    NEED_CTX(ctx);
    if (!a || !b || !out || op < 0 || op > 3) return fail(ctx, ZK_ERR_ARG, "selftest: bad arguments");
    if (curve_id != 377 && curve_id != 381) return fail(ctx, ZK_ERR_ARG, "unknown curve_id");
    if (variant == 3) {  // FP64-limb product (csrc/fq52.cuh), BLS12-377 Fq, multiplication only; result in Montgomery radix 2^416
        if (curve_id != 377 || field != 1 || op != 2) return fail(ctx, ZK_ERR_ARG, "selftest: variant 3 is the BLS12-377 Fq product only");
        return selftest_fq52(ctx, a, b, out, count);
    }
    if (variant == 2) {  // products through the out-of-line multiplier of the MSM inner loop (Fp::mul_call)
        if (curve_id == 377) return field ? selftest_field_call<Fq377>(ctx, op, a, b, out, count) : selftest_field_call<Fr377>(ctx, op, a, b, out, count);
        return field ? selftest_field_call<Fq381>(ctx, op, a, b, out, count) : selftest_field_call<Fr381>(ctx, op, a, b, out, count);
    }
    int sel = (curve_id == 381 ? 4 : 0) | (field ? 2 : 0) | (variant ? 1 : 0);
    switch (sel) {
        case 0: return selftest_field_asm<Fr377>(ctx, op, a, b, out, count);
        case 1: return selftest_field_portable<Fr377>(ctx, op, a, b, out, count);
        case 2: return selftest_field_asm<Fq377>(ctx, op, a, b, out, count);
        case 3: return selftest_field_portable<Fq377>(ctx, op, a, b, out, count);
        case 4: return selftest_field_asm<Fr381>(ctx, op, a, b, out, count);
        case 5: return selftest_field_portable<Fr381>(ctx, op, a, b, out, count);
        case 6: return selftest_field_asm<Fq381>(ctx, op, a, b, out, count);
        default: return selftest_field_portable<Fq381>(ctx, op, a, b, out, count);
    }
}
int zkaes_selftest_g1(zkaes_ctx* ctx, int curve_id, int op, const void* a, const void* b, void* out, size_t count) {
    NEED_CTX(ctx);
    if (!a || !b || !out || op < 0 || op > 2) return fail(ctx, ZK_ERR_ARG, "selftest: bad arguments");
    return CURVE_DISPATCH(ctx, curve_id, selftest_g1<G1_377Params>(ctx, op, a, b, out, count),
                          selftest_g1<G1_381Params>(ctx, op, a, b, out, count));
}

// Host-side (CPU) execution of the same templates: validates the portable arithmetic that the host uses for
// the final window fold / affine normalisation.  No context or GPU needed.
int zkaes_selftest_host_field(int curve_id, int field, int op, const void* a, const void* b, void* out, size_t count) {
    if (!a || !b || !out) return ZK_ERR_ARG;
    auto run = [&](auto tag) {
        using F = decltype(tag);
        const F* x = reinterpret_cast<const F*>(a);
        const F* y = reinterpret_cast<const F*>(b);
        F* o = reinterpret_cast<F*>(out);
        for (size_t i = 0; i < count; ++i) {
            switch (op) {
                case 0: o[i] = x[i] + y[i]; break;
                case 1: o[i] = x[i] - y[i]; break;
                case 2: o[i] = x[i] * y[i]; break;
                case 3: o[i] = x[i].inverse(); break;
                default: o[i] = x[i].neg(); break;
            }
        }
        return ZK_OK;
    };
    if (curve_id == 377) return field ? run(Fq377()) : run(Fr377());
    if (curve_id == 381) return field ? run(Fq381()) : run(Fr381());
    return ZK_ERR_ARG;
}
int zkaes_selftest_host_g1(int curve_id, int op, const void* a, const void* b, void* out, size_t count) {
    if (!a || !b || !out) return ZK_ERR_ARG;
    auto run = [&](auto tag) {
        using C = decltype(tag);
        const Affine<C>* x = reinterpret_cast<const Affine<C>*>(a);
        const Affine<C>* y = reinterpret_cast<const Affine<C>*>(b);
        Affine<C>* o = reinterpret_cast<Affine<C>*>(out);
        for (size_t i = 0; i < count; ++i) {
            XYZZ<C> acc = XYZZ<C>::from_affine(x[i]);
            if (op == 0) {
                acc.madd(y[i]);
            } else if (op == 1) {
                XYZZ<C> q = XYZZ<C>::from_affine(y[i]).dbl();
                q.add(XYZZ<C>::from_affine(y[i]).neg());
                acc.add(q);
            } else {
                acc = acc.dbl();
            }
            o[i] = acc.to_affine();
        }
        return ZK_OK;
    };
    if (curve_id == 377) return run(G1_377Params());
    if (curve_id == 381) return run(G1_381Params());
    return ZK_ERR_ARG;
}

// ---- circuit shape (host only) -----------------------------------------------------------------------------------
struct zkaes_circuit {
    zk::AesCircuit c;
};
int zkaes_circuit_build(size_t msg_len, zkaes_circuit** out) {
    if (!out) return ZK_ERR_ARG;
    *out = nullptr;
    zkaes_circuit* h = new zkaes_circuit();
    try {
        zk::build_aes_circuit(msg_len, h->c);
    } catch (const std::exception&) {
        delete h;
        return ZK_ERR_ARG;
    }
    *out = h;
    return ZK_OK;
}
void zkaes_circuit_free(zkaes_circuit* c) { delete c; }
int zkaes_circuit_info(const zkaes_circuit* h, uint64_t info[ZKAES_CIRCUIT_INFO_WORDS]) {
    if (!h || !info) return ZK_ERR_ARG;
    const zk::AesCircuit& c = h->c;
    uint64_t v[ZKAES_CIRCUIT_INFO_WORDS] = {c.msg_len, c.n_blocks, c.num_instance, c.num_instance_used, c.num_witness, c.num_witness_real,
                                            c.num_constraints, c.a.nnz(), c.b.nnz(), c.c.nnz(), c.wit_key0, c.wit_fixed0, c.wit_block0,
                                            c.wit_block_stride, c.fixed_prog.instrs.size(), c.block_prog.instrs.size(),
                                            c.fixed_prog.level_start.size() - 1, c.block_prog.level_start.size() - 1};
    memcpy(info, v, sizeof(v));
    return ZK_OK;
}
int zkaes_circuit_matrix(const zkaes_circuit* h, int which, uint32_t* row_ptr, uint32_t* col, int8_t* coeff) {
    if (!h || which < 0 || which > 2 || !row_ptr || !col || !coeff) return ZK_ERR_ARG;
    const zk::CsrMatrix& m = which == 0 ? h->c.a : which == 1 ? h->c.b : h->c.c;
    memcpy(row_ptr, m.row_ptr.data(), m.row_ptr.size() * sizeof(uint32_t));
    memcpy(col, m.col.data(), m.col.size() * sizeof(uint32_t));
    memcpy(coeff, m.coeff.data(), m.coeff.size());
    return ZK_OK;
}

int zkaes_witness_aes128_ecb(zkaes_ctx* ctx, const zkaes_circuit* h, const uint8_t* msg, size_t msg_len, const uint8_t key[16],
                             uint8_t* ct_out, uint8_t* assignment_out) {
    NEED_CTX(ctx);
    if (!h || !msg || !key || !ct_out) return fail(ctx, ZK_ERR_ARG, "witness: null pointer");
    const zk::AesCircuit& c = h->c;
    if (msg_len != c.msg_len) return fail(ctx, ZK_ERR_ARG, "witness: message length differs from the circuit's");
    cudaStream_t st = ctx->stream;
    WitnessDev w;
    int rc = witness_upload(ctx, c, w);
    if (rc != ZK_OK) {
        witness_free(w);
        return rc;
    }
    size_t nvar = (size_t)c.num_instance + c.num_witness;
    DevBuf dmsg, dkey, dz, dct;
    auto body = [&]() -> int {
        ZK_CUDA(ctx, dmsg.alloc(msg_len, st));
        ZK_CUDA(ctx, dkey.alloc(16, st));
        ZK_CUDA(ctx, dz.alloc(nvar, st));
        ZK_CUDA(ctx, dct.alloc(msg_len, st));
        ZK_CUDA(ctx, cudaMemcpyAsync(dmsg.p, msg, msg_len, cudaMemcpyHostToDevice, st));
        ZK_CUDA(ctx, cudaMemcpyAsync(dkey.p, key, 16, cudaMemcpyHostToDevice, st));
        ZK_TRY(witness_generate(ctx, c, w, dmsg.as<uint8_t>(), dkey.as<uint8_t>(), dz.as<uint8_t>(), dct.as<uint8_t>()));
        ZK_CUDA(ctx, cudaMemcpyAsync(ct_out, dct.p, msg_len, cudaMemcpyDeviceToHost, st));
        if (assignment_out) ZK_CUDA(ctx, cudaMemcpyAsync(assignment_out, dz.p, nvar, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(ctx, cudaStreamSynchronize(st));
        return ZK_OK;
    };
    rc = body();
    witness_free(w);
    return rc;
}

// ---- keys + encrypt ------------------------------------------------------------------------------------------------
static void pk_handle_free(zkaes_pk* pk) {
    if (!pk) return;
    zk::pk_free(pk->impl);  // cudaFree finds the owning device through the pointer; no device switch needed
    for (zk::zkaes_pk_impl* p : pk->peer_impls) zk::pk_free(p);
    delete pk;
}
int zkaes_synthesize_keys(zkaes_ctx* ctx, size_t plaintext_len, const uint8_t tau_seed32[32], const uint8_t gamma_seed32[32], zkaes_pk** out) {
    NEED_CTX(ctx);
    if (!tau_seed32 || !gamma_seed32 || !out) return fail(ctx, ZK_ERR_ARG, "synthesize_keys: null pointer");
    *out = nullptr;
    zkaes_pk* pk = new zkaes_pk();
    pk->peer_impls.assign(ctx->peers.size(), nullptr);
    int rc = run_on_all(ctx, [&](zkaes_ctx* c, int r) {
        try {
            return zk::pk_synthesize(c, plaintext_len, tau_seed32, gamma_seed32, r == 0 ? &pk->impl : &pk->peer_impls[r - 1]);
        } catch (const std::exception& e) {
            return fail(c, ZK_ERR_STATE, std::string("synthesize_keys: ") + e.what());
        }
    });
    if (rc != ZK_OK) {
        pk_handle_free(pk);
        return rc;
    }
    *out = pk;
    return ZK_OK;
}
void zkaes_pk_free(zkaes_pk* pk) { pk_handle_free(pk); }
int zkaes_pk_save(zkaes_ctx* ctx, const zkaes_pk* pk, const char* path, int flags) {
    NEED_CTX(ctx);
    if (!pk || !path) return fail(ctx, ZK_ERR_ARG, "pk_save: null pointer");
    if (pk->peer_impls.size() != ctx->peers.size()) return fail(ctx, ZK_ERR_STATE, "pk_save: key belongs to another context");
    return run_on_all(ctx, [&](zkaes_ctx* c, int r) {
        try {
            return zk::pk_save(c, r == 0 ? pk->impl : pk->peer_impls[r - 1], rank_path(ctx, path, r).c_str(), flags);
        } catch (const std::exception& e) {
            return fail(c, ZK_ERR_STATE, std::string("pk_save: ") + e.what());
        }
    });
}
int zkaes_pk_load(zkaes_ctx* ctx, const char* path, zkaes_pk** out) {
    NEED_CTX(ctx);
    if (!path || !out) return fail(ctx, ZK_ERR_ARG, "pk_load: null pointer");
    *out = nullptr;
    zkaes_pk* pk = new zkaes_pk();
    pk->peer_impls.assign(ctx->peers.size(), nullptr);
    int rc = run_on_all(ctx, [&](zkaes_ctx* c, int r) {
        try {
            return zk::pk_load(c, rank_path(ctx, path, r).c_str(), r == 0 ? &pk->impl : &pk->peer_impls[r - 1]);
        } catch (const std::exception& e) {
            return fail(c, ZK_ERR_STATE, std::string("pk_load: ") + e.what());
        }
    });
    if (rc != ZK_OK) {
        pk_handle_free(pk);
        return rc;
    }
    *out = pk;
    return ZK_OK;
}
int zkaes_pk_info(const zkaes_pk* pk, uint64_t info[ZKAES_PK_INFO_WORDS]) {
    if (!pk || !info) return ZK_ERR_ARG;
    zk::pk_info(pk->impl, info);
    return ZK_OK;
}
int zkaes_pk_vk_bytes(const zkaes_pk* pk, uint8_t* out, size_t* len) {
    if (!pk || !len) return ZK_ERR_ARG;
    const std::vector<uint8_t>& v = zk::pk_vk_bytes(pk->impl);
    if (out) {
        if (*len < v.size()) return ZK_ERR_ARG;
        memcpy(out, v.data(), v.size());
    }
    *len = v.size();
    return ZK_OK;
}
int zkaes_pk_verifying_key(const zkaes_pk* pk, uint8_t* out, size_t* len) {
    if (!pk || !len) return ZK_ERR_ARG;
    const std::vector<uint8_t>& v = zk::pk_verifying_key(pk->impl);
    if (out) {
        if (*len < v.size()) return ZK_ERR_ARG;
        memcpy(out, v.data(), v.size());
    }
    *len = v.size();
    return ZK_OK;
}
int zkaes_verify_encryption(const uint8_t* vk, size_t vk_len, const uint8_t* proof, size_t proof_len, const uint8_t* ciphertext, size_t ct_len,
                            int* accepted) {
    if (!vk || !proof || (!ciphertext && ct_len) || !accepted) {
        g_host_err = "verify_encryption: null pointer";
        return ZK_ERR_ARG;
    }
    g_host_err.clear();
    int rc = zk::verify_encryption_host(vk, vk_len, proof, proof_len, ciphertext, ct_len, accepted, &g_host_err);
    return rc == 0 ? ZK_OK : ZK_ERR_ARG;
}
int zkaes_proof_deserialize(const uint8_t* proof, size_t proof_len, zkaes_proof_fields* out) {
    if (!proof || !out) {
        g_host_err = "proof_deserialize: null pointer";
        return ZK_ERR_ARG;
    }
    g_host_err.clear();
    return zk::proof_deserialize_host(proof, proof_len, out, &g_host_err) == 0 ? ZK_OK : ZK_ERR_ARG;
}
int zkaes_proof_serialize(const zkaes_proof_fields* in, uint8_t* out, size_t* len) {
    if (!in || !len) {
        g_host_err = "proof_serialize: null pointer";
        return ZK_ERR_ARG;
    }
    g_host_err.clear();
    std::vector<uint8_t> bytes;
    if (zk::proof_serialize_host(in, bytes, &g_host_err) != 0) return ZK_ERR_ARG;
    if (out) {
        if (*len < bytes.size()) {
            g_host_err = "proof_serialize: buffer too small";
            return ZK_ERR_ARG;
        }
        memcpy(out, bytes.data(), bytes.size());
    }
    *len = bytes.size();
    return ZK_OK;
}
int zkaes_selftest_pairing(const uint8_t a32[32], const uint8_t b32[32], uint8_t out576[576]) {
    if (!a32 || !b32 || !out576) return ZK_ERR_ARG;
    zk::pairing_selftest(a32, b32, out576);
    return ZK_OK;
}
int zkaes_encrypt(zkaes_ctx* ctx, const zkaes_pk* pk, const uint8_t* msg, size_t msg_len, const uint8_t key[16], const uint8_t zk_seed32[32],
                  uint8_t* ct_out, uint8_t* proof_out, size_t* proof_len) {
    NEED_CTX(ctx);
    if (!pk || !msg || !key || !zk_seed32 || !ct_out || !proof_len) return fail(ctx, ZK_ERR_ARG, "encrypt: null pointer");
    // 3 rounds of commitments (4 + 3 + 2, two of them with a shifted part), 7 evaluations, 3 empty messages, 2 opening proofs
    const size_t need = 8 + (8 + 4 * 49) + (8 + 3 * 49 + 48) + (8 + 2 * 49 + 48) + 8 + 7 * 32 + 8 + 3 + 8 + (48 + 33) + (48 + 1) + 1;
    if (!proof_out) {
        *proof_len = need;
        return ZK_OK;
    }
    if (*proof_len < need) return fail(ctx, ZK_ERR_ARG, "encrypt: proof buffer too small");
    if (pk->peer_impls.size() != ctx->peers.size()) return fail(ctx, ZK_ERR_STATE, "encrypt: key belongs to another context");
    // every rank runs the prover in lock step (their transcripts are identical; the MSMs meet in the NCCL all-gather); rank 0's
    // outputs go to the caller, the peers' are compared with them
    const size_t nr = ctx->peers.size() + 1;
    std::vector<std::vector<uint8_t>> proofs(nr), cts(nr);
    int rc = run_on_all(ctx, [&](zkaes_ctx* c, int r) {
        try {
            cts[r].resize(msg_len ? msg_len : 1);
            return zk::pk_encrypt(c, r == 0 ? pk->impl : pk->peer_impls[r - 1], msg, msg_len, key, zk_seed32, r == 0 ? ct_out : cts[r].data(), proofs[r]);
        } catch (const std::exception& e) {
            return fail(c, ZK_ERR_STATE, std::string("encrypt: ") + e.what());
        }
    });
    if (rc != ZK_OK) return rc;
    for (size_t r = 1; r < nr; ++r)
        if (proofs[r] != proofs[0] || memcmp(cts[r].data(), ct_out, msg_len) != 0) return fail(ctx, ZK_ERR_STATE, "encrypt: rank " + std::to_string(r) + " diverged from rank 0");
    std::vector<uint8_t>& proof = proofs[0];
    if (proof.size() > *proof_len) return fail(ctx, ZK_ERR_STATE, "encrypt: proof larger than its bound");
    memcpy(proof_out, proof.data(), proof.size());
    *proof_len = proof.size();
    return ZK_OK;
}

}  // extern "C"
