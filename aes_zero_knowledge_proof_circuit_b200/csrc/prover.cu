// The Marlin prover for the AES-128-ECB circuit, device-resident: key synthesis (indexer) and encrypt() (prove).
//
// Stands in for what the reference reaches through simpleworks::marlin (third-party, un-vendored):
//   synthesize_keys  src/lib.rs:138-174 -> generate_universal_srs + generate_proving_and_verifying_keys
//                    (ark-poly-commit 0.3.0 KZG10::setup / MarlinKZG10::trim, ark-marlin 0.3.0 ahp/indexer.rs)
//   encrypt          src/lib.rs:60-114  -> generate_proof (ark-marlin 0.3.0 lib.rs::prove, ahp/prover.rs three rounds,
//                    ark-poly-commit 0.3.0 marlin_pc commit / open_combinations)
// Orchestration, the Fiat-Shamir transcript and the O(1)-size group / field scalars run on the host; every vector of
// size |H| or |K| lives in HBM and is only touched by kernels: K1 (witness.cu), K2 (ntt.cu), K3 (msm.cu) and the
// elementwise passes of polyops.cu.  The proving key (matrices, index polynomials, SRS powers) is uploaded once by
// zkaes_synthesize_keys and stays resident; encrypt() moves msg_len + 48 bytes down and ~1 KB of proof up.
#include "prover.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <thread>

#include "comm.cuh"
#include "msm.cuh"
#include "ntt.cuh"
#include "srs.cuh"
#include "transcript.h"
#include "verifier.h"

namespace zk {

namespace {
using C = G1_377Params;
using Fr = FrS;
using Fq = Fp<Fq377Params>;
using Aff = Affine<C>;
using XY = XYZZ<C>;
constexpr int CURVE = 377;

// ---- host scalars ------------------------------------------------------------------------------------------------------
Fr fr_from_seed(const uint8_t seed32[32]) {  // csrc/srs.cu convention: LE integer, top 4 bits cleared
    Fr t = Fr::zero();
    for (int i = 0; i < 32; ++i) t.v[i >> 2] |= (uint32_t)seed32[i] << (8 * (i & 3));
    t.v[7] &= 0x0fffffffu;
    return t.to_mont();
}
Fr fr_pow_u64(Fr b, uint64_t e) {
    Fr r = Fr::one();
    while (e) {
        if (e & 1) r = r * b;
        b = b.sqr();
        e >>= 1;
    }
    return r;
}
Fr domain_gen(int log_n) {
    Fr g;
    for (int i = 0; i < 8; ++i) g.v[i] = Fr377Params::ROOT(i);
    for (int i = log_n; i < Fr377Params::TWO_ADICITY; ++i) g = g.sqr();
    return g;
}
Fr coset_gen() {
    Fr g;
    for (int i = 0; i < 8; ++i) g.v[i] = Fr377Params::GEN(i);
    return g;
}
Fr vanishing(const Fr& x, size_t n) { return fr_pow_u64(x, n) - Fr::one(); }
// u_H(x, y) = (v_H(x) - v_H(y)) / (x - y), x != y
Fr bivariate_u(const Fr& x, const Fr& y, size_t n) { return (vanishing(x, n) - vanishing(y, n)) * (x - y).inverse(); }
void fr_canonical_bytes(const Fr& x, uint8_t out[32]) {
    Fr c = x.from_mont();
    memcpy(out, c.v, 32);
}
Fr fr_rand(ChaCha20Rng& rng) {
    uint64_t w[4];
    fr_rand_raw<Fr377Params>(rng, w);
    Fr r;
    memcpy(r.v, w, 32);
    return r;
}
Fr fr_from_u128(uint64_t lo, uint64_t hi) {
    Fr r = Fr::zero();
    r.v[0] = (uint32_t)lo; r.v[1] = (uint32_t)(lo >> 32); r.v[2] = (uint32_t)hi; r.v[3] = (uint32_t)(hi >> 32);
    return r.to_mont();
}
Fr sample_outside_domain(ChaCha20Rng& rng, size_t n) {
    for (;;) {
        Fr t = fr_rand(rng);
        if (!vanishing(t, n).is_zero()) return t;
    }
}

// ---- host group arithmetic (O(1) points per proof) ---------------------------------------------------------------------
Aff g1_add(const Aff& a, const Aff& b) {
    XY t = XY::from_affine(a);
    t.madd(b);
    return t.to_affine();
}
Aff g1_mul(const Aff& p, const Fr& s_mont) {
    Fr s = s_mont.from_mont();
    XY acc = XY::inf();
    for (int i = 255; i >= 0; --i) {
        acc = acc.dbl();
        if ((s.v[i >> 5] >> (i & 31)) & 1) acc.madd(p);
    }
    return acc.to_affine();
}
// ark-ff ToBytes for GroupAffine: x || y canonical LE || infinity flag
void g1_to_bytes(const Aff& p, std::vector<uint8_t>& out) {
    uint8_t b[97];
    if (p.is_inf()) {
        memset(b, 0, 97);
        b[48] = 1;  // y = 1
        b[96] = 1;
    } else {
        Fq x = p.x.from_mont(), y = p.y.from_mont();
        memcpy(b, x.v, 48);
        memcpy(b + 48, y.v, 48);
        b[96] = 0;
    }
    out.insert(out.end(), b, b + 97);
}
// ark-serialize 0.3.0 compressed GroupAffine: x canonical LE, bit 7 of the last byte = "y > -y", bit 6 = infinity
void g1_serialize(const Aff& p, std::vector<uint8_t>& out) {
    uint8_t b[48];
    if (p.is_inf()) {
        memset(b, 0, 48);
        b[47] |= 1 << 6;
    } else {
        Fq x = p.x.from_mont(), y = p.y.from_mont(), ny = p.y.neg().from_mont();
        memcpy(b, x.v, 48);
        if (y.canonical_gt(ny)) b[47] |= 1 << 7;
    }
    out.insert(out.end(), b, b + 48);
}
struct Comm {
    Aff comm;
    bool has_shifted = false;
    Aff shifted;
};
void comm_to_bytes(const Comm& c, std::vector<uint8_t>& out) {  // ToBytes of marlin_pc::Commitment
    g1_to_bytes(c.comm, out);
    out.push_back(c.has_shifted ? 1 : 0);
    g1_to_bytes(c.has_shifted ? c.shifted : Aff::inf(), out);
}
void put_u64(std::vector<uint8_t>& out, uint64_t v) {
    for (int i = 0; i < 8; ++i) out.push_back((uint8_t)(v >> (8 * i)));
}
void put_fr(std::vector<uint8_t>& out, const Fr& x) {
    uint8_t b[32];
    fr_canonical_bytes(x, b);
    out.insert(out.end(), b, b + 32);
}

size_t next_pow2(size_t v) {
    size_t n = 1;
    while (n < v) n <<= 1;
    return n;
}
int log2_exact(size_t n) {
    int l = 0;
    while (((size_t)1 << l) < n) ++l;
    return l;
}

// small blinding polynomials live on the host (degree <= 2)
struct Blind {
    std::vector<Fr> c;
    Fr eval(const Fr& x) const {
        Fr acc = Fr::zero();
        for (size_t i = c.size(); i-- > 0;) acc = acc * x + c[i];
        return acc;
    }
    void add_scaled(const Fr& s, const Blind& o) {
        if (o.c.size() > c.size()) c.resize(o.c.size(), Fr::zero());
        for (size_t i = 0; i < o.c.size(); ++i) c[i] = c[i] + s * o.c[i];
    }
    bool is_zero() const {
        for (const Fr& x : c)
            if (!x.is_zero()) return false;
        return true;
    }
    // quotient by (X - z)
    Blind div_linear(const Fr& z) const {
        Blind q;
        if (c.size() < 2) return q;
        q.c.resize(c.size() - 1);
        Fr acc = Fr::zero();
        for (size_t i = c.size(); i-- > 1;) {
            acc = acc * z + c[i];
            q.c[i - 1] = acc;
        }
        return q;
    }
};

// ZKAES_TRACE=1: wall-clock per prover phase on stderr (development aid; synchronises the stream at phase boundaries)
struct PhaseTrace {
    // the largest per-phase pool peak seen since ScratchScope last cleared it: mark() resets the pool's high-water mark at every phase
    // boundary, so under ZKAES_TRACE the scope's own reading at the end of encrypt() would only cover the last phase
    static uint64_t& high_seen() {
        static thread_local uint64_t v = 0;
        return v;
    }
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseTrace(cudaStream_t s) : on(getenv("ZKAES_TRACE") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        // device memory: in use overall (includes what the stream-ordered pool keeps cached), and the pool's own peak of live
        // allocations during this phase (the high-water mark is reset at every phase boundary)
        size_t mfree = 0, mtotal = 0;
        cudaMemGetInfo(&mfree, &mtotal);
        uint64_t used_high = 0, reserved = 0, zero = 0;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPool_t pool = DevBuf::pool();  // the working context's pool (set at the ABI entry)
        if (pool || cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &used_high);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &zero);
            if (used_high > high_seen()) high_seen() = used_high;
        }
        fprintf(stderr, "[zkaes] %-28s %9.2f ms (allocator %7.2f ms)   device %6.1f GB in use | pool: phase peak %6.1f GB live, %6.1f GB reserved\n",
                what, std::chrono::duration<double, std::milli>(t1 - t0).count(), DevBuf::alloc_seconds() * 1e3, (double)(mtotal - mfree) / 1e9,
                (double)used_high / 1e9, (double)reserved / 1e9);
        DevBuf::alloc_seconds() = 0;
        t0 = t1;
    }
};

// Scratch policy of one encrypt() (DevArena, common.cuh): the first call on a context runs on the stream-ordered pool and
// records its peak of live scratch; the next call returns the pool's cache to the driver, carves an arena of that peak + 15 %
// and every later call allocates from it.  ZKAES_ARENA=0 keeps the pool.
struct ScratchScope {
    zkaes_ctx* ctx;
    cudaMemPool_t pool = nullptr;
    bool measuring = false;
    ScratchScope(zkaes_ctx* c, uint64_t domain_h) : ctx(c) {
        static const bool enabled = !(getenv("ZKAES_ARENA") && getenv("ZKAES_ARENA")[0] == '0');
        pool = ctx->pool;
        if (!pool && cudaDeviceGetDefaultMemPool(&pool, ctx->device) != cudaSuccess) pool = nullptr;
        if (ctx->arena_domain != domain_h) {
            // a key of another size: the arena was carved for the previous key's peak -- too small for a larger circuit, tens of
            // GB held for nothing after a smaller one.  Give it back and measure this key's peak on the pool, as on the first call.
            if (ctx->arena.base) {
                cudaStreamSynchronize(ctx->stream);
                cudaFree(ctx->arena.base);
                ctx->arena.reset(nullptr, 0);
            }
            ctx->arena_state = 0;
            ctx->scratch_peak = 0;
            ctx->arena_domain = domain_h;
        }
        if (enabled && ctx->arena_state == 0 && ctx->scratch_peak && pool) {
            cudaStreamSynchronize(ctx->stream);
            cudaMemPoolTrimTo(pool, 0);
            const size_t want = DevArena::round_up((size_t)(ctx->scratch_peak * 1.15) + ((size_t)256 << 20));
            void* base = nullptr;
            if (cudaMalloc(&base, want) == cudaSuccess) {
                ctx->arena.reset((char*)base, want);
                ctx->arena_state = 1;
            } else {
                cudaGetLastError();
                ctx->arena_state = -1;
            }
        }
        if (ctx->arena_state == 1) {
            DevBuf::arena() = &ctx->arena;
        } else if (pool) {
            uint64_t zero = 0;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &zero);
            PhaseTrace::high_seen() = 0;
            measuring = true;
        }
    }
    ~ScratchScope() {
        DevBuf::arena() = nullptr;
        if (getenv("ZKAES_ALLOC_STATS")) {  // no synchronisation: safe to leave on in timed runs
            fprintf(stderr, "[zkaes] encrypt: %.1f ms of host time in cudaMallocAsync/cudaFreeAsync; arena %s (%.1f GB, peak %.1f GB live)\n",
                    DevBuf::alloc_seconds() * 1e3, ctx->arena_state == 1 ? "active" : ctx->arena_state == 0 ? "not yet" : "off",
                    (double)ctx->arena.size / 1e9, (double)ctx->arena.high / 1e9);
            DevBuf::alloc_seconds() = 0;
        }
        if (measuring) {
            uint64_t high = 0;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &high) != cudaSuccess) high = 0;
            if (PhaseTrace::high_seen() > high) high = PhaseTrace::high_seen();  // traced run: the phases' peaks
            if (high > ctx->scratch_peak) ctx->scratch_peak = high;
        }
    }
};

template <class T>
cudaError_t dev_upload(T** dst, const std::vector<T>& src, cudaStream_t st) {
    cudaError_t e = cudaMalloc((void**)dst, std::max<size_t>(src.size() * sizeof(T), 16));
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

}  // namespace

struct zkaes_pk_impl {
    AesCircuit circ;
    size_t h = 0, k = 0, x = 0, D = 0;
    int log_h = 0, log_k = 0, log_x = 0;
    size_t nnz[3] = {0, 0, 0};
    // device-resident
    uint32_t *csr_ptr[3] = {}, *csr_col[3] = {};
    int8_t* csr_cf[3] = {};
    uint32_t *csc_ptr[3] = {}, *csc_row[3] = {};
    int8_t* csc_cf[3] = {};
    uint32_t *krow[3] = {}, *kcol[3] = {};
    int8_t* kcoef[3] = {};
    uint8_t* heavy_flag = nullptr;  // per variable: column has > HEAVY_COL entries over A, B, C
    uint32_t* heavy_cols = nullptr;
    uint32_t* giant_cols = nullptr;  // the heavy columns with more than GIANT_COL entries (the constant one): not in heavy_cols
    size_t n_heavy = 0, n_giant = 0, giant_max_len = 0;  // giant_max_len: bound of a giant column's entries within one matrix
    Fr* elems_h = nullptr;
    Fr* idx_poly[12] = {};  // a_row a_col a_val a_row_col b_... (coefficients, k each)
    Aff* srs = nullptr;     // this rank's share of tau^i G, i <= D: the points i = rank (mod nranks), in the MSM kernels' internal form
    size_t srs_count = 0;
    // Lagrange-basis points for the round-1 commitments (pk_build_lagrange; absent = those commitments use the powers above):
    //   lag[j]  = L_k(tau) G,  lagw[j] = ((L_k(tau) - [k in X] l_k^X(tau)) / v_X(tau)) G   for k = rank + nranks * j < |H|
    Aff *lag = nullptr, *lagw = nullptr;
    size_t lag_count = 0;
    Aff v_h, v_hx;          // v_H(tau) G and (v_H(tau) / v_X(tau)) G (host): the blinding terms r v_H of z_A, z_B and of w
    int small_bits[2] = {0, 0};  // digit width of the row sums of A z and B z over bit assignments (|sum| <= 2^(bits-1)); 0 = too wide
    Aff gamma_g[3];         // gamma tau^i G (host)
    Aff index_comms[12];
    uint8_t tau_seed[32] = {}, gamma_seed[32] = {};  // the test SRS's trapdoor seeds (a key file can regenerate the SRS from them)
    std::vector<uint8_t> vk_bytes;   // IndexVerifierKey ToBytes (enters the Fiat-Shamir seed)
    std::vector<uint8_t> vk_full;    // the verifying key verify_encryption takes (verifier.h)
    WitnessDev wit;
    ~zkaes_pk_impl() {
        for (int m = 0; m < 3; ++m) {
            cudaFree(csr_ptr[m]); cudaFree(csr_col[m]); cudaFree(csr_cf[m]);
            cudaFree(csc_ptr[m]); cudaFree(csc_row[m]); cudaFree(csc_cf[m]);
            cudaFree(krow[m]); cudaFree(kcol[m]); cudaFree(kcoef[m]);
        }
        for (int i = 0; i < 12; ++i) cudaFree(idx_poly[i]);
        cudaFree(elems_h);
        cudaFree(heavy_flag);
        cudaFree(heavy_cols);
        cudaFree(giant_cols);
        cudaFree(srs);
        cudaFree(lag);
        cudaFree(lagw);
        witness_free(wit);
    }
};

int fr_rand_device(zkaes_ctx* ctx, ChaCha20Rng& rng, FrS* out, size_t count);  // rng.cu

namespace {

int ntt(zkaes_ctx* ctx, Fr* data, int log_n, bool inverse, bool coset) { return ntt_device<Fr377Params>(ctx, CURVE, data, log_n, inverse, coset); }

// this rank's contiguous slice of an n-element vector that every rank needs in full afterwards (all-gather in place): equal slices, or the
// whole vector when the ranks do not divide n evenly or the slices would be tiny
void shard_slice(const zkaes_ctx* ctx, size_t n, size_t* start, size_t* count) {
    const size_t N = (size_t)ctx->nranks;
    *start = 0;
    *count = n;
    if (N > 1 && n % N == 0 && n / N >= 4096) {
        *count = n / N;
        *start = (size_t)ctx->rank * *count;
    }
}

// commit(poly) through the device MSM: sum coeffs[i] * srs[offset + i].
// Multi-GPU: rank r keeps the SRS points i = r (mod N) (cyclic, so that polynomials of every length and the shifted
// commitments at the top of the SRS spread evenly), runs the bucket method on the matching every-N-th coefficients, the W
// window sums per rank are all-gathered (NCCL) and folded on every rank: all ranks obtain the same commitment and their
// transcripts stay in lock step.  The window plan comes from the GLOBAL n.
int msm_commit(zkaes_ctx* ctx, const zkaes_pk_impl& pk, const Fr* coeffs, size_t n, size_t offset, Aff* out) {
    if (offset + n > pk.D + 1) return fail(ctx, ZK_ERR_STATE, "commit: polynomial exceeds the SRS");
    const size_t N = (size_t)ctx->nranks, r = (size_t)ctx->rank;
    // local point j is global point r + N j: first j with r + N j >= offset
    const size_t j_lo = offset > r ? (offset - r + N - 1) / N : 0;
    const size_t i0 = r + N * j_lo;
    const size_t n_local = i0 < offset + n ? (offset + n - 1 - i0) / N + 1 : 0;
    if (j_lo + n_local > pk.srs_count) return fail(ctx, ZK_ERR_STATE, "commit: local point range exceeds this rank's SRS share");
    const Aff* bases = pk.srs + (n_local ? j_lo : 0);
    const Fr* scalars = coeffs + (n_local ? i0 - offset : 0);
    MsmPlan p = msm_make_plan(n ? n : 1, Fr377Params::BITS, ctx->msm_window_bits, ctx->nranks, ctx->msm_window_max);
    cudaStream_t st = ctx->stream;
    DevBuf win, all;
    const size_t wbytes = sizeof(XY) * p.W;
    ZK_CUDA(ctx, win.alloc(wbytes, st));
    ZK_CUDA(ctx, all.alloc(wbytes * N, st));
    ZK_TRY(msm_window_sums<C>(ctx, bases, scalars, n_local, /*scalars_mont=*/1, p, win.p, /*bases_internal=*/1, /*scalar_stride=*/N));
    ZK_TRY(comm_all_gather(ctx, win.p, all.p, wbytes));
    std::vector<XY> h((size_t)p.W * N);
    ZK_CUDA(ctx, cudaMemcpyAsync(h.data(), all.p, wbytes * N, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    *out = msm_fold_windows_host<C>(h.data(), (int)N, p);
    return ZK_OK;
}
// KZG10::commit with an optional hiding polynomial of degree hiding_bound + 1 (three draws for hiding_bound = 1)
int kzg_commit(zkaes_ctx* ctx, const zkaes_pk_impl& pk, const Fr* coeffs, size_t n, size_t offset, bool hiding, ChaCha20Rng& zk, Aff* out,
               Blind* blind) {
    ZK_TRY(msm_commit(ctx, pk, coeffs, n, offset, out));
    blind->c.clear();
    if (hiding) {
        for (int i = 0; i < 3; ++i) blind->c.push_back(fr_rand(zk));
        for (int i = 0; i < 3; ++i) *out = g1_add(*out, g1_mul(pk.gamma_g[i], blind->c[i]));
    }
    return ZK_OK;
}
struct Committed {
    Comm comm;
    Blind rand, shifted_rand;
};
// Lagrange-basis commitment (pk_build_lagrange): sum_k vals[k] * basis_k + r * v_point, vals = |H| small signed integers on the device
// (|v| <= 2^(bits-1)), this rank holding the basis points k = rank (mod nranks); then marlin_pc's hiding terms as in kzg_commit.
constexpr int SMALL_WINDOWS = 32;
int lagrange_commit(zkaes_ctx* ctx, const zkaes_pk_impl& pk, const Aff* basis, const int32_t* vals, int bits, const Fr& r_mask, const Aff& v_point,
                    ChaCha20Rng& zk, Committed* out) {
    const size_t N = (size_t)ctx->nranks;
    cudaStream_t st = ctx->stream;
    DevBuf win, all;
    const size_t wbytes = sizeof(XY) * SMALL_WINDOWS;
    ZK_CUDA(ctx, win.alloc(wbytes, st));
    ZK_CUDA(ctx, all.alloc(wbytes * N, st));
    ZK_TRY(msm_small_window_sums<C>(ctx, basis, vals, pk.lag_count, (size_t)ctx->rank, N, bits, SMALL_WINDOWS, win.p));
    ZK_TRY(comm_all_gather(ctx, win.p, all.p, wbytes));
    std::vector<XY> hsum((size_t)SMALL_WINDOWS * N);
    ZK_CUDA(ctx, cudaMemcpyAsync(hsum.data(), all.p, wbytes * N, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    Aff cm = g1_add(msm_sum_windows_host<C>(hsum.data(), hsum.size()), g1_mul(v_point, r_mask));
    out->rand.c.clear();
    for (int i = 0; i < 3; ++i) out->rand.c.push_back(fr_rand(zk));
    for (int i = 0; i < 3; ++i) cm = g1_add(cm, g1_mul(pk.gamma_g[i], out->rand.c[i]));
    out->comm.comm = cm;
    out->comm.has_shifted = false;
    return ZK_OK;
}
// marlin_pc commit of one labeled polynomial (degree bound -> extra commitment on the shifted powers)
int pc_commit(zkaes_ctx* ctx, const zkaes_pk_impl& pk, const Fr* coeffs, size_t n, long bound, bool hiding, ChaCha20Rng& zk, Committed* out) {
    ZK_TRY(kzg_commit(ctx, pk, coeffs, n, 0, hiding, zk, &out->comm.comm, &out->rand));
    out->comm.has_shifted = bound >= 0;
    if (bound >= 0) ZK_TRY(kzg_commit(ctx, pk, coeffs, n, pk.D - (size_t)bound, hiding, zk, &out->comm.shifted, &out->shifted_rand));
    return ZK_OK;
}

}  // namespace

// ======================================================================================================================
// synthesize_keys, in stages that zkaes_pk_load re-uses: shape (circuit, matrices, witness program), SRS share, index
// polynomials (+ their 12 commitments), verifying key
// ======================================================================================================================
namespace {

// circuit shape -> sizes, CSR / CSC / K-domain index arrays and the witness program in HBM, table of omega_H powers
int pk_build_shape(zkaes_ctx* ctx, zkaes_pk_impl& pk, size_t msg_len) {
    try {
        build_aes_circuit(msg_len, pk.circ);
    } catch (const std::exception& e) {
        return fail(ctx, ZK_ERR_ARG, std::string("synthesize_keys: ") + e.what());
    }
    const AesCircuit& c = pk.circ;
    cudaStream_t st = ctx->stream;
    const CsrMatrix* M[3] = {&c.a, &c.b, &c.c};
    for (int m = 0; m < 3; ++m) pk.nnz[m] = M[m]->nnz();
    for (int m = 0; m < 2; ++m) {  // every variable of this circuit is a bit: a row sum lies in [-(sum of negative coefficients), sum of positive ones]
        long bound = 1;
        const CsrMatrix& A = *M[m];
        for (size_t r = 0; r + 1 < A.row_ptr.size(); ++r) {
            long pos = 0, neg = 0;
            for (uint32_t e = A.row_ptr[r]; e < A.row_ptr[r + 1]; ++e) (A.coeff[e] > 0 ? pos : neg) += std::abs((long)A.coeff[e]);
            bound = std::max(bound, std::max(pos, neg));
        }
        int bits = 1;
        while (((long)1 << (bits - 1)) < bound) ++bits;
        pk.small_bits[m] = bits <= 13 ? bits : 0;
    }
    // ark-marlin balance_matrices swaps rows while A is the denser matrix; for this circuit nnz(A) < nnz(B) so it is the identity
    if (pk.nnz[0] >= pk.nnz[1]) return fail(ctx, ZK_ERR_UNSUPPORTED, "synthesize_keys: matrix A denser than B (balance_matrices not implemented)");
    size_t nnz_max = std::max(pk.nnz[0], std::max(pk.nnz[1], pk.nnz[2]));
    pk.h = next_pow2(c.num_constraints);
    pk.k = next_pow2(nnz_max);
    pk.x = c.num_instance;
    pk.log_h = log2_exact(pk.h); pk.log_k = log2_exact(pk.k); pk.log_x = log2_exact(pk.x);
    if (pk.log_k > 28) return fail(ctx, ZK_ERR_UNSUPPORTED, "synthesize_keys: |K| too large for this build (limit 2^28)");
    // AHPForR1CS::max_degree with zk_bound = 1
    pk.D = std::max(std::max(2 * pk.h - 1, 3 * pk.h - 1), std::max(pk.h, 3 * pk.k - 3));
    const size_t h = pk.h, k = pk.k, x = pk.x;
    {
        const size_t N = (size_t)ctx->nranks, r = (size_t)ctx->rank;
        pk.srs_count = pk.D + 1 > r ? (pk.D + 1 - r + N - 1) / N : 0;  // points i = r (mod N), i <= D
    }

    // ---- matrices: CSR, CSC (for t), K-domain arithmetisation indices -------------------------------------------------
    const size_t period = h / x;
    auto reindex = [&](size_t j) -> uint32_t {
        if (j < x) return (uint32_t)(j * period);
        size_t i = j - x;
        return (uint32_t)(i + i / (period - 1) + 1);
    };
    const size_t nvar = (size_t)c.num_instance + c.num_witness;
    std::vector<uint32_t> col_total(nvar, 0);
    for (int m = 0; m < 3; ++m) {
        const CsrMatrix& A = *M[m];
        for (uint32_t cc : A.col) col_total[cc]++;
        ZK_CUDA(ctx, dev_upload(&pk.csr_ptr[m], A.row_ptr, st));
        ZK_CUDA(ctx, dev_upload(&pk.csr_col[m], A.col, st));
        ZK_CUDA(ctx, dev_upload(&pk.csr_cf[m], A.coeff, st));
        std::vector<uint32_t> cptr(nvar + 1, 0), crow(A.nnz());
        std::vector<int8_t> ccf(A.nnz());
        for (uint32_t cc : A.col) cptr[cc + 1]++;
        for (size_t j = 0; j < nvar; ++j) cptr[j + 1] += cptr[j];
        std::vector<uint32_t> fill(cptr.begin(), cptr.end() - 1);
        std::vector<uint32_t> krow(k, 0), kcol(k, 0);
        std::vector<int8_t> kcf(k, 0);
        size_t nrows = A.row_ptr.size() - 1;
        for (size_t r = 0; r < nrows; ++r)
            for (uint32_t e = A.row_ptr[r]; e < A.row_ptr[r + 1]; ++e) {
                uint32_t p = fill[A.col[e]]++;
                crow[p] = (uint32_t)r;
                ccf[p] = A.coeff[e];
                krow[e] = reindex(A.col[e]);  // row(kappa): the VARIABLE's element (arithmetisation of M^*)
                kcol[e] = (uint32_t)r;        // col(kappa): the CONSTRAINT's element
                kcf[e] = A.coeff[e];
            }
        ZK_CUDA(ctx, dev_upload(&pk.csc_ptr[m], cptr, st));
        ZK_CUDA(ctx, dev_upload(&pk.csc_row[m], crow, st));
        ZK_CUDA(ctx, dev_upload(&pk.csc_cf[m], ccf, st));
        ZK_CUDA(ctx, dev_upload(&pk.krow[m], krow, st));
        ZK_CUDA(ctx, dev_upload(&pk.kcol[m], kcol, st));
        ZK_CUDA(ctx, dev_upload(&pk.kcoef[m], kcf, st));
        ZK_CUDA(ctx, cudaStreamSynchronize(st));  // host vectors go out of scope
    }
    {
        constexpr uint32_t HEAVY_COL = 128, GIANT_COL = 1u << 16;
        std::vector<uint8_t> flag(nvar, 0);
        std::vector<uint32_t> cols, giants;
        for (size_t j = 0; j < nvar; ++j)
            if (col_total[j] > GIANT_COL) {
                flag[j] = 1;
                giants.push_back((uint32_t)j);
                pk.giant_max_len = std::max<size_t>(pk.giant_max_len, col_total[j]);  // bound of the column's length in any one matrix
            } else if (col_total[j] > HEAVY_COL) {
                flag[j] = 1;
                cols.push_back((uint32_t)j);
            }
        pk.n_heavy = cols.size();
        pk.n_giant = giants.size();
        ZK_CUDA(ctx, dev_upload(&pk.heavy_flag, flag, st));
        ZK_CUDA(ctx, dev_upload(&pk.heavy_cols, cols, st));
        ZK_CUDA(ctx, dev_upload(&pk.giant_cols, giants, st));
        ZK_CUDA(ctx, cudaStreamSynchronize(st));
    }
    ZK_TRY(witness_upload(ctx, c, pk.wit));
    ZK_CUDA(ctx, cudaMalloc((void**)&pk.elems_h, sizeof(Fr) * h));
    ZK_TRY(po_powers(ctx, pk.elems_h, h, domain_gen(pk.log_h), Fr::one()));
    return ZK_OK;
}

// SRS: this rank's share of tau^i G on the device; gamma tau^i G (i < 3) on the host
int pk_build_srs(zkaes_ctx* ctx, zkaes_pk_impl& pk, bool generate_points) {
    ZK_CUDA(ctx, cudaMalloc((void**)&pk.srs, sizeof(Aff) * std::max<size_t>(pk.srs_count, 1)));
    if (generate_points)
        ZK_TRY(srs_powers_device<C>(ctx, pk.tau_seed, pk.srs_count, pk.srs, /*start=*/(size_t)ctx->rank, /*stride=*/(size_t)ctx->nranks));
    Fr tau = fr_from_seed(pk.tau_seed), gamma = fr_from_seed(pk.gamma_seed);
    Fr gt = gamma;
    for (int i = 0; i < 3; ++i) {
        pk.gamma_g[i] = g1_mul(Aff::generator(), gt);
        gt = gt * tau;
    }
    return ZK_OK;
}

// Lagrange-basis points over H for the round-1 commitments (msm_commit_small).  The evaluations of z_A, z_B over H are row sums of a few bits
// and the numerator of w is the bit assignment itself, so in the basis {L_k(tau) G} their commitments are MSMs with one small signed digit per
// term instead of W = 11 windows of a 253-bit coefficient: same group elements, same proof bytes.
//   z_M^(X) = sum_k (Mz)_k L_k(X) + r v_H(X)
//   w^(X)   = (sum_{k not in X} w_k L_k(X) - (x^(X) - sum_j x_j L_{k_j}(X)) + r v_H(X)) / v_X(X)      [x^ interpolates itself over H, k_j = j |H|/|X|]
//           = sum_{k in H} a_k (L_k(X) - [k = k_j] l_j^X(X)) / v_X(X) + r v_H(X) / v_X(X),             a = the full assignment in H order
// and every (L_k - [k in X] l_j^X) / v_X is a polynomial (L_k vanishes on X for k outside it; L_{k_j} - l_j^X vanishes on X), so the points
// are commitments an SRS holder can derive (group-element inverse NTT); with the test SRS's tau in hand they are fixed-base multiples:
//   L_k(tau) = v_H(tau)/|H| * w^k / (tau - w^k),   l_j^X(tau) = v_X(tau)/|X| * w^k / (tau - w^k)  (k = k_j).
// 2 x 96 B per element of H (12.9 GB at 4 KiB on one GPU, sharded cyclically like the SRS): built only if that leaves room for the prover's
// scratch (ZKAES_LAGRANGE=0 disables it).
int pk_build_lagrange(zkaes_ctx* ctx, zkaes_pk_impl& pk) {
    if (getenv("ZKAES_LAGRANGE") && getenv("ZKAES_LAGRANGE")[0] == '0') return ZK_OK;
    const size_t h = pk.h, x = pk.x, N = (size_t)ctx->nranks, r = (size_t)ctx->rank;
    cudaStream_t st = ctx->stream;
    const Fr tau = fr_from_seed(pk.tau_seed);
    const Fr vh = vanishing(tau, h), vx = vanishing(tau, x);
    if (vh.is_zero() || vx.is_zero()) return ZK_OK;  // tau inside H: no Lagrange form (the powers still work)
    const size_t count = h > r ? (h - r + N - 1) / N : 0;
    {   // room: the two point arrays for good, three |H| vectors of scalars while building, and afterwards the prover's scratch (about 480 B per
        // element of H at its round-3 / opening peak, measured 30.6 GB at |H| = 2^26) plus a margin
        ZK_CUDA(ctx, cudaStreamSynchronize(st));
        if (ctx->pool) cudaMemPoolTrimTo(ctx->pool, 0);
        size_t mfree = 0, mtotal = 0;
        ZK_CUDA(ctx, cudaMemGetInfo(&mfree, &mtotal));
        const size_t keep = 2 * sizeof(Aff) * count, build = 3 * sizeof(Fr) * h, later = 560 * h / N + ((size_t)2 << 30);
        // every rank must take the same path (the commitments' all-gathers pair up): all build, or none does
        const uint32_t mine = mfree >= keep + std::max(build, later) ? 1u : 0u;
        DevBuf flag, flags;
        ZK_CUDA(ctx, flag.alloc(sizeof(uint32_t), st));
        ZK_CUDA(ctx, flags.alloc(sizeof(uint32_t) * N, st));
        ZK_CUDA(ctx, cudaMemcpyAsync(flag.p, &mine, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        ZK_TRY(comm_all_gather(ctx, flag.p, flags.p, sizeof(uint32_t)));
        std::vector<uint32_t> all(N);
        ZK_CUDA(ctx, cudaMemcpyAsync(all.data(), flags.p, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(ctx, cudaStreamSynchronize(st));
        for (uint32_t f : all)
            if (!f) return ZK_OK;
    }
    ZK_CUDA(ctx, cudaMalloc((void**)&pk.lag, sizeof(Aff) * std::max<size_t>(count, 1)));
    ZK_CUDA(ctx, cudaMalloc((void**)&pk.lagw, sizeof(Aff) * std::max<size_t>(count, 1)));
    DevBuf d, u, sc;
    ZK_CUDA(ctx, d.alloc(sizeof(Fr) * h, st));
    ZK_CUDA(ctx, u.alloc(sizeof(Fr) * h, st));
    ZK_CUDA(ctx, sc.alloc(sizeof(Fr) * h, st));
    ZK_TRY(po_rsub_scalar(ctx, d.as<Fr>(), pk.elems_h, tau, h));       // tau - w^k
    ZK_TRY(po_batch_inverse(ctx, u.as<Fr>(), d.as<Fr>(), h));
    ZK_TRY(po_vec(ctx, 2, u.as<Fr>(), u.as<Fr>(), pk.elems_h, h));     // u_k = w^k / (tau - w^k)
    const Fr hinv = Fr::from_u64(h).inverse(), xinv = Fr::from_u64(x).inverse();
    const Fr c_l = vh * hinv;                  // L_k(tau) = c_l u_k
    const Fr c_w = c_l * vx.inverse();         // L_k(tau) / v_X(tau)
    ZK_TRY(po_scale(ctx, sc.as<Fr>(), u.as<Fr>(), c_l, h));
    ZK_TRY(fb_mul_scalars_device<C>(ctx, sc.p, count, pk.lag, r, N));
    ZK_TRY(po_scale(ctx, sc.as<Fr>(), u.as<Fr>(), c_w, h));
    ZK_TRY(po_scale_strided(ctx, sc.as<Fr>(), u.as<Fr>(), c_w - xinv, h / x, x));  // k in X: (L_k - l_j^X) / v_X = u_k (c_w - 1/|X|)
    ZK_TRY(fb_mul_scalars_device<C>(ctx, sc.p, count, pk.lagw, r, N));
    pk.lag_count = count;
    pk.v_h = g1_mul(Aff::generator(), vh);
    pk.v_hx = g1_mul(Aff::generator(), vh * vx.inverse());
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    return ZK_OK;
}

// index polynomials (ahp/indexer.rs arithmetize_matrix) and, unless they come from a key file, their 12 commitments
int pk_build_index_polys(zkaes_ctx* ctx, zkaes_pk_impl& pk, bool compute, bool commit) {
    const size_t h = pk.h, k = pk.k;
    Fr hinv = Fr::from_u64(h).inverse();
    for (int m = 0; m < 3; ++m) {
        Fr** P = &pk.idx_poly[4 * m];
        for (int j = 0; j < 4; ++j) ZK_CUDA(ctx, cudaMalloc((void**)&P[j], sizeof(Fr) * k));
        if (!compute) continue;
        ZK_TRY(po_gather(ctx, P[0], pk.elems_h, pk.krow[m], k));
        ZK_TRY(po_gather(ctx, P[1], pk.elems_h, pk.kcol[m], k));
        // val = M / u_H(row, row), u_H(y, y) = |H| y^(|H|-1) = |H| / y
        ZK_TRY(po_gather_scaled(ctx, P[2], pk.elems_h, pk.krow[m], pk.kcoef[m], hinv, k));
        ZK_TRY(po_vec(ctx, 2, P[3], P[0], P[1], k));
        for (int j = 0; j < 4; ++j) {
            ZK_TRY(ntt(ctx, P[j], pk.log_k, true, false));
            if (commit) ZK_TRY(msm_commit(ctx, pk, P[j], k, 0, &pk.index_comms[4 * m + j]));
        }
    }
    return ZK_OK;
}

// IndexVerifierKey ToBytes (index_info (3 x u64) || 12 commitments) and the VerifyingKey half of synthesize_keys' result
void pk_build_vk(zkaes_pk_impl& pk) {
    const AesCircuit& c = pk.circ;
    const size_t nvar = (size_t)c.num_instance + c.num_witness;
    const size_t nnz_max = std::max(pk.nnz[0], std::max(pk.nnz[1], pk.nnz[2]));
    pk.vk_bytes.clear();
    put_u64(pk.vk_bytes, nvar);
    put_u64(pk.vk_bytes, c.num_constraints);
    put_u64(pk.vk_bytes, nnz_max);
    for (int i = 0; i < 12; ++i) {
        Comm cm;
        cm.comm = pk.index_comms[i];
        comm_to_bytes(cm, pk.vk_bytes);
    }
    pk.vk_full = build_verifying_key(nvar, c.num_constraints, nnz_max, pk.x, pk.index_comms, pk.D, fr_from_seed(pk.tau_seed), fr_from_seed(pk.gamma_seed),
                                     {pk.h - 2, pk.k - 2});
}

}  // namespace

int pk_synthesize(zkaes_ctx* ctx, size_t msg_len, const uint8_t tau_seed[32], const uint8_t gamma_seed[32], zkaes_pk_impl** out) {
    *out = nullptr;
    std::unique_ptr<zkaes_pk_impl> pkp(new zkaes_pk_impl());
    zkaes_pk_impl& pk = *pkp;
    memcpy(pk.tau_seed, tau_seed, 32);
    memcpy(pk.gamma_seed, gamma_seed, 32);
    PhaseTrace tr(ctx->stream);
    ZK_TRY(pk_build_shape(ctx, pk, msg_len));
    tr.mark("keys: circuit + matrices");
    ZK_TRY(pk_build_srs(ctx, pk, true));
    tr.mark("keys: SRS share");
    ZK_TRY(pk_build_index_polys(ctx, pk, true, true));
    tr.mark("keys: index polys + 12 commitments");
    pk_build_vk(pk);
    tr.mark("keys: verifying key");
    ZK_TRY(pk_build_lagrange(ctx, pk));
    tr.mark("keys: Lagrange-basis points");
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = pkp.release();
    return ZK_OK;
}

// ======================================================================================================================
// key files (the reference regenerates SRS and keys in every process, src/lib.rs:138-174)
// ======================================================================================================================
// Layout (little-endian): "ZKAESPK1" | version u64 | msg_len | nranks | rank | flags | D | srs_count | |K| | tau_seed[32] | gamma_seed[32]
//   | 12 index commitments (96 B affine, arkworks Montgomery limbs) | vk_full length u64 | vk_full (ark-serialize VerifyingKey)
//   | [flags & 1: this rank's SRS share, srs_count x 96 B] | [flags & 2: 12 index polynomials, |K| x 32 B each]
// What is absent is recomputed by zkaes_pk_load: the circuit shape and matrices always (deterministic in msg_len), the SRS from the
// seeds (test SRS), the index polynomials from the matrices.  The 12 index commitments -- the expensive part of synthesize_keys,
// twelve |K|-term MSMs -- are never recomputed.  One file per rank: the SRS share is the rank's.
namespace {
constexpr char PK_MAGIC[8] = {'Z', 'K', 'A', 'E', 'S', 'P', 'K', '1'};
constexpr size_t IO_CHUNK = (size_t)64 << 20;
struct File {
    FILE* f = nullptr;
    ~File() {
        if (f) fclose(f);
    }
};
int dev_to_file(zkaes_ctx* ctx, FILE* f, const void* dev, size_t bytes) {
    std::vector<uint8_t> buf(std::min(bytes, IO_CHUNK));
    for (size_t off = 0; off < bytes; off += IO_CHUNK) {
        const size_t n = std::min(IO_CHUNK, bytes - off);
        ZK_CUDA(ctx, cudaMemcpyAsync(buf.data(), (const char*)dev + off, n, cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (fwrite(buf.data(), 1, n, f) != n) return fail(ctx, ZK_ERR_STATE, "pk_save: short write");
    }
    return ZK_OK;
}
int file_to_dev(zkaes_ctx* ctx, FILE* f, void* dev, size_t bytes) {
    std::vector<uint8_t> buf(std::min(bytes, IO_CHUNK));
    for (size_t off = 0; off < bytes; off += IO_CHUNK) {
        const size_t n = std::min(IO_CHUNK, bytes - off);
        if (fread(buf.data(), 1, n, f) != n) return fail(ctx, ZK_ERR_ARG, "pk_load: truncated key file");
        ZK_CUDA(ctx, cudaMemcpyAsync((char*)dev + off, buf.data(), n, cudaMemcpyHostToDevice, ctx->stream));
        ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ZK_OK;
}
}  // namespace

int pk_save(zkaes_ctx* ctx, const zkaes_pk_impl* pk, const char* path, int flags) {
    if ((flags & ~3) != 0) return fail(ctx, ZK_ERR_ARG, "pk_save: unknown flags");
    File file;
    file.f = fopen(path, "wb");
    if (!file.f) return fail(ctx, ZK_ERR_ARG, std::string("pk_save: cannot open ") + path);
    std::vector<uint8_t> hd(PK_MAGIC, PK_MAGIC + 8);
    for (uint64_t v : {(uint64_t)1, (uint64_t)pk->circ.msg_len, (uint64_t)ctx->nranks, (uint64_t)ctx->rank, (uint64_t)flags, (uint64_t)pk->D,
                       (uint64_t)pk->srs_count, (uint64_t)pk->k})
        put_u64(hd, v);
    hd.insert(hd.end(), pk->tau_seed, pk->tau_seed + 32);
    hd.insert(hd.end(), pk->gamma_seed, pk->gamma_seed + 32);
    const uint8_t* comms = reinterpret_cast<const uint8_t*>(pk->index_comms);
    hd.insert(hd.end(), comms, comms + sizeof(pk->index_comms));
    put_u64(hd, pk->vk_full.size());
    hd.insert(hd.end(), pk->vk_full.begin(), pk->vk_full.end());
    if (fwrite(hd.data(), 1, hd.size(), file.f) != hd.size()) return fail(ctx, ZK_ERR_STATE, "pk_save: short write");
    if (flags & 1) ZK_TRY(dev_to_file(ctx, file.f, pk->srs, sizeof(Aff) * pk->srs_count));
    if (flags & 2)
        for (int i = 0; i < 12; ++i) ZK_TRY(dev_to_file(ctx, file.f, pk->idx_poly[i], sizeof(Fr) * pk->k));
    if (fflush(file.f) != 0) return fail(ctx, ZK_ERR_STATE, "pk_save: flush failed");
    return ZK_OK;
}

int pk_load(zkaes_ctx* ctx, const char* path, zkaes_pk_impl** out) {
    *out = nullptr;
    File file;
    file.f = fopen(path, "rb");
    if (!file.f) return fail(ctx, ZK_ERR_ARG, std::string("pk_load: cannot open ") + path);
    uint8_t fixed[8 + 8 * 8 + 64];
    if (fread(fixed, 1, sizeof(fixed), file.f) != sizeof(fixed) || memcmp(fixed, PK_MAGIC, 8) != 0) return fail(ctx, ZK_ERR_ARG, "pk_load: not a zkaes key file");
    uint64_t v[8];
    memcpy(v, fixed + 8, sizeof(v));
    const uint64_t version = v[0], msg_len = v[1], nranks = v[2], rank = v[3], flags = v[4], D = v[5], srs_count = v[6], kk = v[7];
    if (version != 1 || (flags & ~(uint64_t)3)) return fail(ctx, ZK_ERR_ARG, "pk_load: unsupported key file version / flags");
    if (nranks != (uint64_t)ctx->nranks || rank != (uint64_t)ctx->rank)
        return fail(ctx, ZK_ERR_STATE, "pk_load: the key file holds the SRS share of another rank layout (rank " + std::to_string(rank) + " of " +
                                           std::to_string(nranks) + ")");
    std::unique_ptr<zkaes_pk_impl> pkp(new zkaes_pk_impl());
    zkaes_pk_impl& pk = *pkp;
    memcpy(pk.tau_seed, fixed + 8 + 64, 32);
    memcpy(pk.gamma_seed, fixed + 8 + 64 + 32, 32);
    PhaseTrace tr(ctx->stream);
    ZK_TRY(pk_build_shape(ctx, pk, (size_t)msg_len));
    if (pk.D != D || pk.srs_count != srs_count || pk.k != kk) return fail(ctx, ZK_ERR_ARG, "pk_load: key file does not match the circuit of its message length");
    tr.mark("load: circuit + matrices");
    if (fread(pk.index_comms, 1, sizeof(pk.index_comms), file.f) != sizeof(pk.index_comms)) return fail(ctx, ZK_ERR_ARG, "pk_load: truncated key file");
    for (int i = 0; i < 12; ++i)
        if (!pk.index_comms[i].on_curve()) return fail(ctx, ZK_ERR_ARG, "pk_load: index commitment not on the curve");
    uint64_t vk_len = 0;
    if (fread(&vk_len, 1, 8, file.f) != 8 || vk_len > ((uint64_t)1 << 20)) return fail(ctx, ZK_ERR_ARG, "pk_load: truncated key file");
    std::vector<uint8_t> vk_file(vk_len);
    if (fread(vk_file.data(), 1, vk_len, file.f) != vk_len) return fail(ctx, ZK_ERR_ARG, "pk_load: truncated key file");
    ZK_TRY(pk_build_srs(ctx, pk, !(flags & 1)));
    if (flags & 1) ZK_TRY(file_to_dev(ctx, file.f, pk.srs, sizeof(Aff) * pk.srs_count));
    tr.mark((flags & 1) ? "load: SRS share from the file" : "load: SRS share from the seeds");
    ZK_TRY(pk_build_index_polys(ctx, pk, !(flags & 2), false));
    if (flags & 2)
        for (int i = 0; i < 12; ++i) ZK_TRY(file_to_dev(ctx, file.f, pk.idx_poly[i], sizeof(Fr) * pk.k));
    tr.mark((flags & 2) ? "load: index polynomials from the file" : "load: index polynomials recomputed");
    if (fgetc(file.f) != EOF) return fail(ctx, ZK_ERR_ARG, "pk_load: trailing bytes in the key file");
    pk_build_vk(pk);
    if (pk.vk_full != vk_file) return fail(ctx, ZK_ERR_ARG, "pk_load: the verifying key in the file differs from the one its commitments and seeds give");
    ZK_TRY(pk_build_lagrange(ctx, pk));
    tr.mark("load: Lagrange-basis points from the seeds");
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = pkp.release();
    return ZK_OK;
}

void pk_free(zkaes_pk_impl* pk) { delete pk; }
const std::vector<uint8_t>& pk_vk_bytes(const zkaes_pk_impl* pk) { return pk->vk_bytes; }
const std::vector<uint8_t>& pk_verifying_key(const zkaes_pk_impl* pk) { return pk->vk_full; }
void pk_info(const zkaes_pk_impl* pk, uint64_t info[ZK_PK_INFO_WORDS]) {
    const AesCircuit& c = pk->circ;
    uint64_t v[ZK_PK_INFO_WORDS] = {c.msg_len, c.num_constraints, (uint64_t)c.num_instance + c.num_witness, pk->nnz[0], pk->nnz[1], pk->nnz[2],
                                    pk->h, pk->k, pk->x, pk->D, c.num_instance_used, pk->lag ? pk->lag_count : 0};
    memcpy(info, v, sizeof(v));
}

// Multi-GPU evaluation of `ncoset` independent cosets (rounds 2 and 3).  Coset j needs `ntask` forward transforms (task(dst, j, p)
// writes transform p of coset j into dst, `elems` field elements) which its owner folds into the coset's result (fold(j, T): T[p] =
// transform p) -- the inverse transform and the pointwise products stay with the owner.  Cosets are processed in waves of at most
// nranks cosets with distinct owners; inside a wave the forward transforms are spread over ALL ranks by coset_plan (comm.cuh):
//   1. every rank computes the transforms assigned to it (its own coset's into T, the others' into helper buffers),
//   2. one NCCL group of sends / receives moves the helpers' transforms to the owners over NVLink,
//   3. the owners fold.
// The results are then broadcast by the caller.  Field arithmetic is exact, so who computes a transform does not change a bit.
template <class Task, class Fold>
int run_coset_waves(zkaes_ctx* ctx, int ncoset, int ntask, double own_extra, size_t elems, Task task, Fold fold) {
    cudaStream_t st = ctx->stream;
    const int N = ctx->nranks, me = ctx->rank;
    const size_t bytes = sizeof(Fr) * elems;
    for (int j0 = 0; j0 < ncoset; j0 += N) {
        const int nw = std::min(N, ncoset - j0);
        const CosetPlan plan = coset_plan(N, nw, ntask, own_extra);
        const bool owner = me < nw;  // owner of coset j0 + me
        int n_help = 0;
        for (int jj = 0; jj < nw; ++jj)
            for (int p = 0; p < ntask; ++p) n_help += (plan.executor(jj, p) == me && plan.owner[jj] != me);
        const int n_own = owner ? ntask : 0;
        std::unique_ptr<DevBuf[]> T(new DevBuf[n_own ? n_own : 1]), H(new DevBuf[n_help ? n_help : 1]);  // DevBuf is neither copyable nor movable
        std::vector<Fr*> Tp(ntask, nullptr);
        for (int p = 0; p < n_own; ++p) {
            ZK_CUDA(ctx, T[p].alloc(bytes, st));
            Tp[p] = T[p].as<Fr>();
        }
        for (int i = 0; i < n_help; ++i) ZK_CUDA(ctx, H[i].alloc(bytes, st));
        // 1. compute
        int hi = 0;
        for (int jj = 0; jj < nw; ++jj)
            for (int p = 0; p < ntask; ++p)
                if (plan.executor(jj, p) == me) ZK_TRY(task(plan.owner[jj] == me ? Tp[p] : H[hi++].as<Fr>(), j0 + jj, p));
        // 2. exchange (helpers -> owners)
        if (N > 1) {
            ZK_TRY(comm_group_start(ctx));
            hi = 0;
            int rc = ZK_OK;
            for (int jj = 0; jj < nw && rc == ZK_OK; ++jj)
                for (int p = 0; p < ntask && rc == ZK_OK; ++p) {
                    const int e = plan.executor(jj, p), o = plan.owner[jj];
                    if (e == o) continue;
                    if (me == e) rc = comm_send(ctx, H[hi++].p, bytes, o);
                    if (me == o) rc = comm_recv(ctx, Tp[p], bytes, e);
                }
            const int rc_end = comm_group_end(ctx);
            if (rc != ZK_OK) return rc;
            ZK_TRY(rc_end);
        }
        // 3. fold
        if (owner) ZK_TRY(fold(j0 + me, Tp));
    }
    return ZK_OK;
}

// ======================================================================================================================
// encrypt(): witness + prove
// ======================================================================================================================
int pk_encrypt(zkaes_ctx* ctx, const zkaes_pk_impl* pkp, const uint8_t* msg, size_t msg_len, const uint8_t key[16], const uint8_t zk_seed[32],
               uint8_t* ct_out, std::vector<uint8_t>& proof) {
    const zkaes_pk_impl& pk = *pkp;
    const AesCircuit& c = pk.circ;
    if (msg_len != c.msg_len) return fail(ctx, ZK_ERR_ARG, "encrypt: message length differs from the proving key's");
    cudaStream_t st = ctx->stream;
    const size_t h = pk.h, k = pk.k, x = pk.x, D = pk.D;
    const size_t nvar = (size_t)c.num_instance + c.num_witness;
    ChaCha20Rng zk(zk_seed);
    PhaseTrace tr(st);
    ScratchScope scratch(ctx, pk.h);

    // ---- K1: witness ---------------------------------------------------------------------------------------------------
    DevBuf dmsg, dkey, dz, dct;
    ZK_CUDA(ctx, dmsg.alloc(msg_len, st));
    ZK_CUDA(ctx, dkey.alloc(16, st));
    ZK_CUDA(ctx, dz.alloc(nvar, st));
    ZK_CUDA(ctx, dct.alloc(msg_len, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(dmsg.p, msg, msg_len, cudaMemcpyHostToDevice, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(dkey.p, key, 16, cudaMemcpyHostToDevice, st));
    ZK_TRY(witness_generate(ctx, c, pk.wit, dmsg.as<uint8_t>(), dkey.as<uint8_t>(), dz.as<uint8_t>(), dct.as<uint8_t>()));
    ZK_CUDA(ctx, cudaMemcpyAsync(ct_out, dct.p, msg_len, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    const uint8_t* z = dz.as<uint8_t>();

    tr.mark("witness");
    // ---- Fiat-Shamir seed: protocol name || index vk || public input (unformatted: without the leading one, zero padded)
    std::vector<uint8_t> seed(pk.vk_bytes.size() + 11 + 32 * (x - 1));
    memcpy(seed.data(), "MARLIN-2019", 11);
    memcpy(seed.data() + 11, pk.vk_bytes.data(), pk.vk_bytes.size());
    {
        uint8_t* p = seed.data() + 11 + pk.vk_bytes.size();
        memset(p, 0, 32 * (x - 1));
        for (size_t i = 0; i < 8 * msg_len; ++i) p[32 * i] = (ct_out[i >> 3] >> (i & 7)) & 1;  // byte_to_field_array (src/helpers/mod.rs:84-93)
    }
    FiatShamirRng fs(seed);

    tr.mark("fs seed");
    // ---- first round -----------------------------------------------------------------------------------------------------
    DevBuf x_poly, x_evals, wbuf, w_poly, za, zb, mask, rem;
    ZK_CUDA(ctx, x_poly.alloc(sizeof(Fr) * x, st));
    ZK_TRY(po_bits_to_fr(ctx, x_poly.as<Fr>(), z, x));  // formatted input = the instance section of z
    ZK_TRY(ntt(ctx, x_poly.as<Fr>(), pk.log_x, true, false));
    // w, z_A, z_B are three independent chains (w: two |H| transforms and the division by v_X; z_M: a sparse product and one transform): with
    // several ranks each chain has one owner and the coefficient vectors are broadcast (2.1 GB each at 4 KiB) instead of every rank repeating all three
    const int nr = ctx->nranks;
    const int own_w = 0, own_z[2] = {nr > 1 ? 1 : 0, nr > 2 ? 2 : (nr > 1 ? 1 : 0)};
    Fr r_w = fr_rand(zk), r_a = fr_rand(zk), r_b = fr_rand(zk);
    const size_t len_w = h + 1 - x;
    ZK_CUDA(ctx, w_poly.alloc(sizeof(Fr) * len_w, st));
    if (ctx->rank == own_w) {
        ZK_CUDA(ctx, x_evals.alloc(sizeof(Fr) * h, st));
        ZK_CUDA(ctx, cudaMemsetAsync(x_evals.p, 0, sizeof(Fr) * h, st));
        ZK_CUDA(ctx, cudaMemcpyAsync(x_evals.p, x_poly.p, sizeof(Fr) * x, cudaMemcpyDeviceToDevice, st));
        ZK_TRY(ntt(ctx, x_evals.as<Fr>(), pk.log_h, false, false));
        ZK_CUDA(ctx, wbuf.alloc(sizeof(Fr) * (h + 1), st));
        ZK_TRY(po_w_evals(ctx, wbuf.as<Fr>(), z, x_evals.as<Fr>(), h, h / x, c.num_instance, c.num_witness));
        x_evals.release();
        ZK_TRY(ntt(ctx, wbuf.as<Fr>(), pk.log_h, true, false));
        ZK_CUDA(ctx, cudaMemsetAsync(wbuf.as<Fr>() + h, 0, sizeof(Fr), st));
        ZK_TRY(po_add_vanishing(ctx, wbuf.as<Fr>(), h, r_w));
        ZK_CUDA(ctx, rem.alloc(sizeof(Fr) * h, st));
        ZK_TRY(po_divide_vanishing(ctx, wbuf.as<Fr>(), h + 1, x, w_poly.as<Fr>(), rem.as<Fr>()));
        wbuf.release();
        rem.release();
    }
    ZK_CUDA(ctx, za.alloc(sizeof(Fr) * (h + 1), st));
    ZK_CUDA(ctx, zb.alloc(sizeof(Fr) * (h + 1), st));
    DevBuf* zab[2] = {&za, &zb};
    const Fr r_ab[2] = {r_a, r_b};
    for (int m = 0; m < 2; ++m) {
        if (ctx->rank != own_z[m]) continue;
        Fr* p = zab[m]->as<Fr>();
        ZK_TRY(po_spmv_bits(ctx, p, pk.csr_ptr[m], pk.csr_col[m], pk.csr_cf[m], z, c.num_constraints, h));
        ZK_TRY(ntt(ctx, p, pk.log_h, true, false));
        ZK_CUDA(ctx, cudaMemsetAsync(p + h, 0, sizeof(Fr), st));
        ZK_TRY(po_add_vanishing(ctx, p, h, r_ab[m]));
    }
    if (nr > 1) {
        ZK_TRY(comm_broadcast(ctx, w_poly.p, sizeof(Fr) * len_w, own_w));
        for (int m = 0; m < 2; ++m) ZK_TRY(comm_broadcast(ctx, zab[m]->p, sizeof(Fr) * (h + 1), own_z[m]));
    }
    tr.mark("r1: w, z_a, z_b polys");
    // mask polynomial: degree 3|H| + 2 zk - 3; force sum over H to zero by fixing the constant term
    const size_t len_mask = 3 * h;
    ZK_CUDA(ctx, mask.alloc(sizeof(Fr) * len_mask, st));
    ZK_TRY(fr_rand_device(ctx, zk, mask.as<Fr>(), len_mask));
    ZK_TRY(po_mask_fix(ctx, mask.as<Fr>(), h));  // mask[0] = -(mask[|H|] + mask[2|H|])
    tr.mark("r1: mask sample+upload");
    Committed c_w, c_za, c_zb, c_mask;
    if (pk.lag && ctx->r1_lagrange && pk.small_bits[0] && pk.small_bits[1]) {
        // w, z_A, z_B in the Lagrange basis: one small digit per element of H (the bit assignment; the row sums of A z, B z)
        DevBuf vals;
        ZK_CUDA(ctx, vals.alloc(sizeof(int32_t) * h, st));
        ZK_TRY(po_assignment_h_i32(ctx, vals.as<int32_t>(), z, h, h / x, c.num_instance, c.num_witness));
        ZK_TRY(lagrange_commit(ctx, pk, pk.lagw, vals.as<int32_t>(), 1, r_w, pk.v_hx, zk, &c_w));
        Committed* cz[2] = {&c_za, &c_zb};
        for (int m = 0; m < 2; ++m) {
            ZK_TRY(po_spmv_bits_i32(ctx, vals.as<int32_t>(), pk.csr_ptr[m], pk.csr_col[m], pk.csr_cf[m], z, c.num_constraints, h));
            ZK_TRY(lagrange_commit(ctx, pk, pk.lag, vals.as<int32_t>(), pk.small_bits[m], r_ab[m], pk.v_h, zk, cz[m]));
        }
    } else {
        ZK_TRY(pc_commit(ctx, pk, w_poly.as<Fr>(), len_w, -1, true, zk, &c_w));
        ZK_TRY(pc_commit(ctx, pk, za.as<Fr>(), h + 1, -1, true, zk, &c_za));
        ZK_TRY(pc_commit(ctx, pk, zb.as<Fr>(), h + 1, -1, true, zk, &c_zb));
    }
    ZK_TRY(pc_commit(ctx, pk, mask.as<Fr>(), len_mask, -1, false, zk, &c_mask));
    {
        std::vector<uint8_t> b;
        for (const Committed* cm : {&c_w, &c_za, &c_zb, &c_mask}) comm_to_bytes(cm->comm, b);
        fs.absorb(b);
    }
    tr.mark("r1: 4 commitments");
    Fr alpha = sample_outside_domain(fs.rng, h);
    Fr eta[3] = {fr_rand(fs.rng), fr_rand(fs.rng), fr_rand(fs.rng)};

    // ---- second round ------------------------------------------------------------------------------------------------------
    const Fr vh_alpha = vanishing(alpha, h);
    DevBuf ra, tpoly, zpoly, tmp;
    ZK_CUDA(ctx, ra.alloc(sizeof(Fr) * h, st));
    {
        size_t h0 = 0, hn = h;  // elementwise over H: a slice per rank, all-gathered in place
        shard_slice(ctx, h, &h0, &hn);
        ZK_CUDA(ctx, tmp.alloc(sizeof(Fr) * hn, st));
        Fr* rs = ra.as<Fr>() + h0;
        ZK_TRY(po_rsub_scalar(ctx, tmp.as<Fr>(), pk.elems_h + h0, alpha, hn));
        ZK_TRY(po_batch_inverse(ctx, rs, tmp.as<Fr>(), hn));
        ZK_TRY(po_scale(ctx, rs, rs, vh_alpha, hn));  // r(alpha, h_i) = v_H(alpha) / (alpha - h_i)
        if (hn != h) ZK_TRY(comm_all_gather(ctx, rs, ra.p, sizeof(Fr) * hn));
    }
    tmp.release();
    ZK_CUDA(ctx, tpoly.alloc(sizeof(Fr) * h, st));
    CscView csc[3];
    for (int m = 0; m < 3; ++m) csc[m] = CscView{pk.csc_ptr[m], pk.csc_row[m], pk.csc_cf[m]};
    ZK_TRY(po_t_evals(ctx, tpoly.as<Fr>(), csc, eta, ra.as<Fr>(), pk.heavy_flag, pk.heavy_cols, pk.n_heavy, pk.giant_cols, pk.n_giant, pk.giant_max_len, nvar, h, x));
    ZK_TRY(ntt(ctx, tpoly.as<Fr>(), pk.log_h, true, false));
    ZK_TRY(ntt(ctx, ra.as<Fr>(), pk.log_h, true, false));  // r_alpha polynomial
    ZK_CUDA(ctx, zpoly.alloc(sizeof(Fr) * (h + 1), st));
    ZK_TRY(po_z_poly(ctx, zpoly.as<Fr>(), w_poly.as<Fr>(), len_w, x_poly.as<Fr>(), x));
    tr.mark("r2: r_alpha, t, z polys");
    // rhs = r_alpha (eta_a z_a + eta_b z_b + eta_c z_a z_b) - t z: deg r_alpha = |H| - 1, deg z_a = deg z_b = |H| (one blinding coefficient),
    // deg t < |H|, deg z = |H|, so deg rhs = 3|H| - 1 -- 3|H| coefficients, like the mask.  The reference multiplies on the 4|H| domain
    // (five forward NTTs of 4|H| live at once: 43 GB at 4 KiB).  Here q_1 = mask + rhs is evaluated on THREE cosets s_j H of size |H|,
    // s_j = w_4H^j (s_j^|H| = 1, i, -1): every buffer is |H|-sized, each coset yields rhs mod (X^|H| - s_j^|H|) = c_0 + s_j^|H| c_1 + s_j^2|H| c_2,
    // and a 3 x 3 Vandermonde solve per index returns the three coefficient blocks (po_coset3_combine, as for h_2 in round 3).
    // [round 1 walked all four cosets of the 4|H| domain: 24 transforms of |H| instead of 18.]
    const size_t n3 = 3 * h;
    const int log4h = pk.log_h + 2;
    constexpr int NCOSET2 = 3;
    DevBuf e_ra, k_ra, k_za, k_zb, k_t, k_z;
    ZK_CUDA(ctx, e_ra.alloc(sizeof(Fr) * n3, st));
    for (DevBuf* bfr : {&k_ra, &k_za, &k_zb, &k_t, &k_z}) ZK_CUDA(ctx, bfr->alloc(sizeof(Fr) * h, st));
    {
        const Fr w4h = domain_gen(log4h);
        const Fr i4 = fr_pow_u64(w4h, h);  // s_j^|H| = i4^j
        // dst[a] = p(s w_H^a) for a polynomial of len <= |H| + 1 coefficients: reduce mod X^|H| - s^|H|, shift, transform
        auto to_coset_h = [&](Fr* dst, const Fr* poly, size_t len, const Fr& s, const Fr& s_h) -> int {
            const size_t lo = len < h ? len : h;
            ZK_CUDA(ctx, cudaMemcpyAsync(dst, poly, sizeof(Fr) * lo, cudaMemcpyDeviceToDevice, st));
            if (lo < h) ZK_CUDA(ctx, cudaMemsetAsync(dst + lo, 0, sizeof(Fr) * (h - lo), st));
            if (len > h) ZK_TRY(po_axpy(ctx, dst, poly + h, s_h, len - h));  // X^|H| = s^|H| on the coset
            ZK_TRY(po_scale_powers(ctx, dst, dst, s, h));
            return ntt(ctx, dst, pk.log_h, false, false);
        };
        // s_j and s_j^|H| of the cosets
        Fr s4[NCOSET2], s4h[NCOSET2];
        s4[0] = Fr::one();
        s4h[0] = Fr::one();
        for (int j = 1; j < NCOSET2; ++j) {
            s4[j] = s4[j - 1] * w4h;
            s4h[j] = s4h[j - 1] * i4;
        }
        if (ctx->nranks == 1) {
            for (int j = 0; j < NCOSET2; ++j) {
                Fr* Rj = e_ra.as<Fr>() + (size_t)j * h;
                ZK_TRY(to_coset_h(k_ra.as<Fr>(), ra.as<Fr>(), h, s4[j], s4h[j]));
                ZK_TRY(to_coset_h(k_za.as<Fr>(), za.as<Fr>(), h + 1, s4[j], s4h[j]));
                ZK_TRY(to_coset_h(k_zb.as<Fr>(), zb.as<Fr>(), h + 1, s4[j], s4h[j]));
                ZK_TRY(to_coset_h(k_t.as<Fr>(), tpoly.as<Fr>(), h, s4[j], s4h[j]));
                ZK_TRY(to_coset_h(k_z.as<Fr>(), zpoly.as<Fr>(), h + 1, s4[j], s4h[j]));
                ZK_TRY(po_round2(ctx, Rj, k_ra.as<Fr>(), k_za.as<Fr>(), k_zb.as<Fr>(), k_t.as<Fr>(), k_z.as<Fr>(), eta, h));
                ZK_TRY(ntt(ctx, Rj, pk.log_h, true, false));
                ZK_TRY(po_scale_powers(ctx, Rj, Rj, s4[j].inverse(), h));
            }
        } else {
            // Multi-GPU: coset j is assembled by rank (j mod N) and broadcast (2.1 GB at 4 KiB); its five forward transforms are
            // spread over all ranks (run_coset_waves)
            for (DevBuf* bfr : {&k_ra, &k_za, &k_zb, &k_t, &k_z}) bfr->release();
            const Fr* src[5] = {ra.as<Fr>(), za.as<Fr>(), zb.as<Fr>(), tpoly.as<Fr>(), zpoly.as<Fr>()};
            const size_t len[5] = {h, h + 1, h + 1, h, h + 1};
            ZK_TRY(run_coset_waves(
                ctx, NCOSET2, 5, 1.5, h, [&](Fr* dst, int j, int p) -> int { return to_coset_h(dst, src[p], len[p], s4[j], s4h[j]); },
                [&](int j, const std::vector<Fr*>& T) -> int {
                    Fr* Rj = e_ra.as<Fr>() + (size_t)j * h;
                    ZK_TRY(po_round2(ctx, Rj, T[0], T[1], T[2], T[3], T[4], eta, h));
                    ZK_TRY(ntt(ctx, Rj, pk.log_h, true, false));
                    return po_scale_powers(ctx, Rj, Rj, s4[j].inverse(), h);
                }));
            for (int j = 0; j < NCOSET2; ++j) ZK_TRY(comm_broadcast(ctx, e_ra.as<Fr>() + (size_t)j * h, sizeof(Fr) * h, j % ctx->nranks));
        }
        ZK_TRY(po_coset3_combine(ctx, e_ra.as<Fr>(), h, s4h));  // rhs coefficients (degree 3|H| - 1)
    }
    for (DevBuf* bfr : {&k_ra, &k_za, &k_zb, &k_t, &k_z}) bfr->release();
    ra.release();
    zpoly.release();
    ZK_TRY(po_vec(ctx, 0, e_ra.as<Fr>(), e_ra.as<Fr>(), mask.as<Fr>(), len_mask));  // q_1 = mask + rhs
    DevBuf h1, xg1;
    ZK_CUDA(ctx, h1.alloc(sizeof(Fr) * (n3 - h), st));
    ZK_CUDA(ctx, xg1.alloc(sizeof(Fr) * h, st));
    ZK_TRY(po_divide_vanishing(ctx, e_ra.as<Fr>(), n3, h, h1.as<Fr>(), xg1.as<Fr>()));
    e_ra.release();
    const Fr* g1 = xg1.as<Fr>() + 1;  // q_1 = h_1 v_H + X g_1
    const size_t len_g1 = h - 1, len_h1 = 2 * h;  // deg h_1 = deg q_1 - |H| = 2|H| - 1 (ark-marlin commits 2|H| + 1 coefficients: the top one is zero)
    tr.mark("r2: products + division");
    Committed c_t, c_g1, c_h1;
    ZK_TRY(pc_commit(ctx, pk, tpoly.as<Fr>(), h, -1, false, zk, &c_t));
    ZK_TRY(pc_commit(ctx, pk, g1, len_g1, (long)(h - 2), true, zk, &c_g1));
    ZK_TRY(pc_commit(ctx, pk, h1.as<Fr>(), len_h1, -1, true, zk, &c_h1));
    {
        std::vector<uint8_t> b;
        for (const Committed* cm : {&c_t, &c_g1, &c_h1}) comm_to_bytes(cm->comm, b);
        fs.absorb(b);
    }
    tr.mark("r2: 3 commitments");
    Fr beta = sample_outside_domain(fs.rng, h);

    // ---- third round -------------------------------------------------------------------------------------------------------
    const Fr vh_beta = vanishing(beta, h);
    const Fr vv = vh_alpha * vh_beta;
    const Fr hinv = Fr::from_u64(h).inverse();
    // The hiding part of the opening at beta, S = mask + c_za z_a + c_w w + c_h1 h_1, only needs beta, t(beta) and z_b(beta):
    // form it now (in place in the mask buffer) so that z_a, w and h_1 (10.7 GB at 4 KiB) are not carried through round 3,
    // the prover's memory peak.  Field arithmetic is exact, so ch^2 * S later equals the term-by-term combination.
    Fr ev_t, ev_zb;
    ZK_TRY(po_eval(ctx, tpoly.as<Fr>(), h, beta, &ev_t));
    ZK_TRY(po_eval(ctx, zb.as<Fr>(), h + 1, beta, &ev_zb));
    const Fr c_za_lc = bivariate_u(alpha, beta, h) * (eta[0] + eta[2] * ev_zb);
    const Fr c_w_lc = (ev_t * vanishing(beta, x)).neg();
    const Fr c_h1_lc = vh_beta.neg();
    ZK_TRY(po_axpy(ctx, mask.as<Fr>(), za.as<Fr>(), c_za_lc, h + 1));
    ZK_TRY(po_axpy(ctx, mask.as<Fr>(), w_poly.as<Fr>(), c_w_lc, len_w));
    ZK_TRY(po_axpy(ctx, mask.as<Fr>(), h1.as<Fr>(), c_h1_lc, len_h1));
    za.release(); w_poly.release(); h1.release();
    DevBuf fpoly, den, inv;
    ZK_CUDA(ctx, fpoly.alloc(sizeof(Fr) * k, st));
    {
        // f over K is elementwise (three denominators, their batch inversion, the weighted sum): with several ranks each takes a slice of K and
        // the slices are all-gathered in place (32 |K| bytes over NVLink against 7/8 of the inversions at N = 8); the inverse transform needs all of f
        size_t k0 = 0, kn = k;
        shard_slice(ctx, k, &k0, &kn);
        ZK_CUDA(ctx, den.alloc(sizeof(Fr) * kn, st));
        ZK_CUDA(ctx, inv.alloc(sizeof(Fr) * kn, st));
        Fr* fs = fpoly.as<Fr>() + k0;
        ZK_CUDA(ctx, cudaMemsetAsync(fs, 0, sizeof(Fr) * kn, st));
        for (int m = 0; m < 3; ++m) {
            ZK_TRY(po_den_k(ctx, den.as<Fr>(), pk.elems_h, pk.krow[m] + k0, pk.kcol[m] + k0, alpha, beta, kn));
            ZK_TRY(po_batch_inverse(ctx, inv.as<Fr>(), den.as<Fr>(), kn));
            ZK_TRY(po_gather_fma(ctx, fs, pk.elems_h, pk.krow[m] + k0, pk.kcoef[m] + k0, inv.as<Fr>(), eta[m] * hinv * vv, kn));
        }
        if (kn != k) ZK_TRY(comm_all_gather(ctx, fs, fpoly.p, sizeof(Fr) * kn));
    }
    den.release(); inv.release();
    ZK_TRY(ntt(ctx, fpoly.as<Fr>(), pk.log_k, true, false));
    const Fr* g2 = fpoly.as<Fr>() + 1;  // f = X g_2 + t(beta) / |K|
    const size_t len_g2 = k - 1;
    tr.mark("r3: f");
    // h_2 = (a - b f) / v_K has degree < 3|K|: it is evaluated on THREE cosets s_j K, s_j = g w_4K^j, of size |K| each
    // (the reference evaluates on the coset g*B, |B| = 4|K|, and keeps 12 tables of 4|K| in the index, ahp/indexer.rs
    // evals_on_B -- 206 GB at 4 KiB).  Every buffer is |K|-sized, v_K is the constant s_j^K - 1 on each coset, and the
    // three interpolants h_2 mod (X^K - s_j^K) = c_0 + s_j^K c_1 + s_j^2K c_2 give the coefficient blocks c_b by a 3x3 solve.
    const int log4k = pk.log_k + 2;
    const Fr ab = alpha * beta;
    const Fr g = coset_gen(), w4k = domain_gen(log4k);
    DevBuf V;
    ZK_CUDA(ctx, V.alloc(sizeof(Fr) * 3 * k, st));
    auto to_coset = [&](Fr* dst, const Fr* poly, const Fr& shift) -> int {  // dst[i] = poly(shift * w_K^i)
        ZK_TRY(po_scale_powers(ctx, dst, poly, shift, k));
        return ntt(ctx, dst, pk.log_k, false, false);
    };
    constexpr int NCOSET = 3;
    Fr s3[NCOSET], uj[NCOSET];
    s3[0] = g;
    for (int j = 1; j < NCOSET; ++j) s3[j] = s3[j - 1] * w4k;
    for (int j = 0; j < NCOSET; ++j) uj[j] = fr_pow_u64(s3[j], k);
    // forward transform p of coset j: 0..2 the denominators of A, B, C (linear in (row, col, row_col): the coefficient vectors are
    // combined first, one NTT instead of three), 3 = f, 4..6 = val of A, B, C
    auto r3_task = [&](Fr* dst, int j, int p) -> int {
        if (p < 3) {
            Fr* const* P = &pk.idx_poly[4 * p];
            ZK_TRY(po_lincomb_den(ctx, dst, P[0], P[1], P[3], alpha.neg(), beta.neg(), ab, k));
            return to_coset(dst, dst, s3[j]);
        }
        return to_coset(dst, p == 3 ? fpoly.as<Fr>() : pk.idx_poly[4 * (p - 4) + 2], s3[j]);
    };
    // V_j <- (a - b f) / v_K on coset j from the seven transforms, then back to the coefficients of the degree-<|K| interpolant
    auto r3_finish = [&](int j) -> int {
        Fr* Vj = V.as<Fr>() + (size_t)j * k;
        const Fr vk_inv = (uj[j] - Fr::one()).inverse();
        ZK_TRY(po_scale(ctx, Vj, Vj, vk_inv, k));
        ZK_TRY(ntt(ctx, Vj, pk.log_k, true, false));
        return po_scale_powers(ctx, Vj, Vj, s3[j].inverse(), k);  // the shift undone
    };
    if (ctx->nranks == 1) {
        // one GPU: four |K|-sized work buffers, transforms consumed as they are produced
        DevBuf dden[3], tA;
        for (int m = 0; m < 3; ++m) ZK_CUDA(ctx, dden[m].alloc(sizeof(Fr) * k, st));
        ZK_CUDA(ctx, tA.alloc(sizeof(Fr) * k, st));
        for (int j = 0; j < NCOSET; ++j) {
            Fr* Vj = V.as<Fr>() + (size_t)j * k;
            for (int m = 0; m < 3; ++m) ZK_TRY(r3_task(dden[m].as<Fr>(), j, m));
            ZK_TRY(r3_task(tA.as<Fr>(), j, 3));
            ZK_TRY(po_mul3(ctx, Vj, dden[0].as<Fr>(), dden[1].as<Fr>(), dden[2].as<Fr>(), Fr::one().neg(), k));
            ZK_TRY(po_vec(ctx, 2, Vj, Vj, tA.as<Fr>(), k));  // - b * f
            for (int m = 0; m < 3; ++m) {
                ZK_TRY(r3_task(tA.as<Fr>(), j, 4 + m));
                ZK_TRY(po_fma3(ctx, Vj, tA.as<Fr>(), dden[(m + 1) % 3].as<Fr>(), dden[(m + 2) % 3].as<Fr>(), vv * eta[m], k));  // + a
            }
            ZK_TRY(r3_finish(j));
        }
    } else {
        // Multi-GPU: coset j is assembled by rank (j mod N) and broadcast (NVLink; 4.3 GB per coset at 4 KiB); its seven forward
        // transforms are spread over all ranks (run_coset_waves) instead of leaving N - 3 ranks idle
        ZK_TRY(run_coset_waves(
            ctx, NCOSET, 7, 1.5, k, r3_task, [&](int j, const std::vector<Fr*>& T) -> int {
                Fr* Vj = V.as<Fr>() + (size_t)j * k;
                ZK_TRY(po_mul3(ctx, Vj, T[0], T[1], T[2], Fr::one().neg(), k));
                ZK_TRY(po_vec(ctx, 2, Vj, Vj, T[3], k));
                for (int m = 0; m < 3; ++m) ZK_TRY(po_fma3(ctx, Vj, T[4 + m], T[(m + 1) % 3], T[(m + 2) % 3], vv * eta[m], k));
                return r3_finish(j);
            }));
        for (int j = 0; j < NCOSET; ++j) ZK_TRY(comm_broadcast(ctx, V.as<Fr>() + (size_t)j * k, sizeof(Fr) * k, j % ctx->nranks));
    }
    ZK_TRY(po_coset3_combine(ctx, V.as<Fr>(), k, uj));
    Fr* h2 = V.as<Fr>();  // 3|K| - 3 coefficients
    const size_t len_h2 = 3 * k - 3;
    tr.mark("r3: h_2 on the coset");
    Committed c_g2, c_h2;
    ZK_TRY(pc_commit(ctx, pk, g2, len_g2, (long)(k - 2), false, zk, &c_g2));
    ZK_TRY(pc_commit(ctx, pk, h2, len_h2, -1, false, zk, &c_h2));
    {
        std::vector<uint8_t> b;
        for (const Committed* cm : {&c_g2, &c_h2}) comm_to_bytes(cm->comm, b);
        fs.absorb(b);
    }
    tr.mark("r3: 2 commitments");
    Fr gamma = fr_rand(fs.rng);

    // ---- evaluations (sorted by label: a_denom b_denom c_denom g_1 g_2 t z_b) ---------------------------------------------
    Fr ev_g1, ev_g2, ev_den[3];
    ZK_TRY(po_eval(ctx, g1, len_g1, beta, &ev_g1));
    ZK_TRY(po_eval(ctx, g2, len_g2, gamma, &ev_g2));
    for (int m = 0; m < 3; ++m) {
        Fr er, ec, erc;
        ZK_TRY(po_eval(ctx, pk.idx_poly[4 * m + 0], k, gamma, &er));
        ZK_TRY(po_eval(ctx, pk.idx_poly[4 * m + 1], k, gamma, &ec));
        ZK_TRY(po_eval(ctx, pk.idx_poly[4 * m + 3], k, gamma, &erc));
        ev_den[m] = ab - alpha * er - beta * ec + erc;
    }
    const Fr evals[7] = {ev_den[0], ev_den[1], ev_den[2], ev_g1, ev_g2, ev_t, ev_zb};
    {
        std::vector<uint8_t> b;
        for (const Fr& e : evals) put_fr(b, e);
        fs.absorb(b);
    }
    uint64_t lo = fs.rng.next_u64(), hi = fs.rng.next_u64();
    const Fr ch = fr_from_u128(lo, hi);
    Fr chp[6];
    chp[0] = Fr::one();
    for (int i = 1; i < 6; ++i) chp[i] = chp[i - 1] * ch;

    tr.mark("evaluations");
    // ---- openings (marlin_pc open_combinations -> batch_open per query point) ------------------------------------------
    // beta: g_1 [ch^0, shifted ch^1], outer_sumcheck [ch^2], t [ch^3], z_b [ch^4]
    Aff w_beta, w_gamma;
    Fr rv_beta;
    {
        // x(beta) for the constant term is not part of the opened polynomial (constants are dropped by open_combinations)
        DevBuf& P = mask;  // holds S = mask + c_za z_a + c_w w + c_h1 h_1 since the start of round 3
        ZK_TRY(po_scale(ctx, P.as<Fr>(), P.as<Fr>(), chp[2], len_mask));                   // ch^2 * S
        ZK_TRY(po_axpy(ctx, P.as<Fr>(), g1, chp[0], len_g1));
        ZK_TRY(po_axpy(ctx, P.as<Fr>(), tpoly.as<Fr>(), chp[3], h));
        ZK_TRY(po_axpy(ctx, P.as<Fr>(), zb.as<Fr>(), chp[4], h + 1));
        Blind r;
        r.add_scaled(chp[0], c_g1.rand);
        r.add_scaled(chp[2] * c_za_lc, c_za.rand);
        r.add_scaled(chp[2] * c_w_lc, c_w.rand);
        r.add_scaled(chp[2] * c_h1_lc, c_h1.rand);
        r.add_scaled(chp[4], c_zb.rand);
        DevBuf Q;
        ZK_CUDA(ctx, Q.alloc(sizeof(Fr) * len_mask, st));
        ZK_TRY(po_div_linear(ctx, P.as<Fr>(), len_mask, beta, Q.as<Fr>()));
        ZK_TRY(msm_commit(ctx, pk, Q.as<Fr>(), len_mask - 1, 0, &w_beta));
        Blind rw = r.div_linear(beta);
        for (size_t i = 0; i < rw.c.size(); ++i) w_beta = g1_add(w_beta, g1_mul(pk.gamma_g[i], rw.c[i]));
        rv_beta = r.eval(beta);
        // shifted part: witness of g_1 alone on the powers from D - (|H| - 2), times ch^1
        ZK_TRY(po_div_linear(ctx, g1, len_g1, beta, Q.as<Fr>()));
        Aff sw;
        ZK_TRY(msm_commit(ctx, pk, Q.as<Fr>(), len_g1 - 1, D - (h - 2), &sw));
        Blind srw = c_g1.shifted_rand.div_linear(beta);
        for (size_t i = 0; i < srw.c.size(); ++i) sw = g1_add(sw, g1_mul(pk.gamma_g[i], srw.c[i]));
        w_beta = g1_add(w_beta, g1_mul(sw, chp[1]));
        rv_beta = rv_beta + chp[1] * c_g1.shifted_rand.eval(beta);
    }
    // nothing below needs the round-1 / round-2 polynomials: return them to the pool before the 3|K|-sized opening
    mask.release(); zb.release(); tpoly.release(); xg1.release();
    tr.mark("opening at beta");
    // gamma: a_denom [ch^0], b_denom [ch^1], c_denom [ch^2], g_2 [ch^3, shifted ch^4], inner_sumcheck [ch^5]; nothing is hiding
    {
        const Fr vk_gamma = vanishing(gamma, k);
        const size_t len_p = len_h2;
        DevBuf& P = V;  // h_2 is consumed here: build the combination in place (V holds 3|K| >= len_p elements)
        ZK_TRY(po_scale(ctx, P.as<Fr>(), h2, (chp[5] * vk_gamma).neg(), len_p));
        const Fr inner_c[3] = {eta[0] * ev_den[1] * ev_den[2] * vv, eta[1] * ev_den[0] * ev_den[2] * vv, eta[2] * ev_den[1] * ev_den[0] * vv};
        for (int m = 0; m < 3; ++m) {
            ZK_TRY(po_axpy(ctx, P.as<Fr>(), pk.idx_poly[4 * m + 2], chp[5] * inner_c[m], k));
            ZK_TRY(po_axpy(ctx, P.as<Fr>(), pk.idx_poly[4 * m + 0], (chp[m] * alpha).neg(), k));
            ZK_TRY(po_axpy(ctx, P.as<Fr>(), pk.idx_poly[4 * m + 1], (chp[m] * beta).neg(), k));
            ZK_TRY(po_axpy(ctx, P.as<Fr>(), pk.idx_poly[4 * m + 3], chp[m], k));
        }
        ZK_TRY(po_axpy(ctx, P.as<Fr>(), g2, chp[3], len_g2));
        DevBuf Q;
        ZK_CUDA(ctx, Q.alloc(sizeof(Fr) * len_p, st));
        ZK_TRY(po_div_linear(ctx, P.as<Fr>(), len_p, gamma, Q.as<Fr>()));
        ZK_TRY(msm_commit(ctx, pk, Q.as<Fr>(), len_p - 1, 0, &w_gamma));
        ZK_TRY(po_div_linear(ctx, g2, len_g2, gamma, Q.as<Fr>()));
        Aff sw;
        ZK_TRY(msm_commit(ctx, pk, Q.as<Fr>(), len_g2 - 1, D - (k - 2), &sw));
        w_gamma = g1_add(w_gamma, g1_mul(sw, chp[4]));
    }
    ZK_CUDA(ctx, cudaStreamSynchronize(st));

    tr.mark("opening at gamma");
    // ---- ark-serialize 0.3.0 CanonicalSerialize of ark_marlin::Proof ----------------------------------------------------------
    proof.clear();
    put_u64(proof, 3);
    const std::vector<const Committed*> rounds[3] = {{&c_w, &c_za, &c_zb, &c_mask}, {&c_t, &c_g1, &c_h1}, {&c_g2, &c_h2}};
    for (const auto& rnd : rounds) {
        put_u64(proof, rnd.size());
        for (const Committed* cm : rnd) {
            g1_serialize(cm->comm.comm, proof);
            proof.push_back(cm->comm.has_shifted ? 1 : 0);
            if (cm->comm.has_shifted) g1_serialize(cm->comm.shifted, proof);
        }
    }
    put_u64(proof, 7);
    for (const Fr& e : evals) put_fr(proof, e);
    put_u64(proof, 3);
    proof.push_back(0); proof.push_back(0); proof.push_back(0);  // three ProverMsg::EmptyMessage
    put_u64(proof, 2);
    g1_serialize(w_beta, proof);
    proof.push_back(1);
    put_fr(proof, rv_beta);
    g1_serialize(w_gamma, proof);
    proof.push_back(0);  // random_v = None
    proof.push_back(0);  // BatchLCProof.evals = None
    return ZK_OK;
}

}  // namespace zk
