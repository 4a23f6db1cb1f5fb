// K2: radix-2 NTT over Fr for sm_100a (natural order in, natural order out).
//
// Replaces ark-poly 0.3.0 `Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place`
// (reference Cargo.lock:234-236; reached from src/lib.rs:111 through ark-marlin's prover rounds).
// Same domain convention: group generator w = ROOT^(2^(TWO_ADICITY-log_n)), coset shift = multiplicative
// generator GEN, inverse scales by n^-1.
//
// Schedule: in-place bit-reversal (optionally fused with the coset pre-scale), then ceil(log_n/8)-ish
// decimation-in-time passes.  Each pass stages a tile of 2^k x G elements in shared memory as eight 32-bit
// limb planes (conflict-free butterflies), runs k butterfly stages, and writes back in place, so HBM traffic
// is 64 B/element per pass.  Twiddles come from a cached table w^e, e < n/2; the inverse transform reuses
// the same table through w^-e = -w^(n/2-e).
#include "ntt.cuh"

namespace zk {

static constexpr int TILE_LOG = 11;          // 2048 elements x 32 B = 64 KB of shared memory per block
static constexpr int TILE = 1 << TILE_LOG;
static constexpr int NTT_THREADS = 256;
static constexpr int POW_LO_LOG = 10;

template <class F>
__device__ __forceinline__ F load_fr(const F* p) {
    F r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class F>
__device__ __forceinline__ F load_fr_ro(const F* p) {
    F r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class F>
__device__ __forceinline__ void store_fr(F* p, const F& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// out[i] = c * base^(i << shift), i < count   (square-and-multiply per entry; tables are built once and cached)
template <class F>
__global__ void k_pow_table(F* out, size_t count, F base, F c, int shift) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long e = (unsigned long long)i << shift;
    F r = c, b = base;
    while (e) {
        if (e & 1) r = r * b;
        b = b.sqr();
        e >>= 1;
    }
    out[i] = r;
}

// in-place bit reversal; if lo/hi tables are given also multiplies element i by lo[i & mask] * hi[i >> POW_LO_LOG]
template <class F>
__global__ void __launch_bounds__(256) k_bitrev(F* data, int log_n, const F* __restrict__ pw_lo, const F* __restrict__ pw_hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)1 << log_n;
    if (i >= n) return;
    size_t j = log_n ? (__brevll((unsigned long long)i) >> (64 - log_n)) : 0;
    if (j < i) return;
    F a = load_fr(data + i);
    if (pw_lo) {
        a = a * load_fr_ro(pw_lo + (i & ((1u << POW_LO_LOG) - 1)));
        if (i >> POW_LO_LOG) a = a * load_fr_ro(pw_hi + (i >> POW_LO_LOG));
    }
    if (j == i) {
        if (pw_lo) store_fr(data + i, a);
        return;
    }
    F b = load_fr(data + j);
    if (pw_lo) {
        b = b * load_fr_ro(pw_lo + (j & ((1u << POW_LO_LOG) - 1)));
        if (j >> POW_LO_LOG) b = b * load_fr_ro(pw_hi + (j >> POW_LO_LOG));
    }
    store_fr(data + i, b);
    store_fr(data + j, a);
}

// data[i] *= lo[i & mask] * hi[i >> POW_LO_LOG]
template <class F>
__global__ void __launch_bounds__(256) k_scale_pow(F* data, size_t n, const F* __restrict__ pw_lo, const F* __restrict__ pw_hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F a = load_fr(data + i);
    a = a * load_fr_ro(pw_lo + (i & ((1u << POW_LO_LOG) - 1)));
    if (i >> POW_LO_LOG) a = a * load_fr_ro(pw_hi + (i >> POW_LO_LOG));
    store_fr(data + i, a);
}

// One DIT pass: global stages s0+1 .. s0+k on bit-reversed-ordered data.
//   element index = hi * 2^(s0+k) + mid * 2^s0 + lo ; a tile holds all 2^k `mid` for G consecutive `lo`
//   (s0 > 0) or for G consecutive `hi` (s0 == 0).  G = 2^glog.
template <class F>
__global__ void __launch_bounds__(NTT_THREADS) k_ntt_pass(F* data, const F* __restrict__ tw, int log_n, int s0, int k, int glog,
                                                          int inverse, int do_scale, F scale) {
    extern __shared__ uint32_t sm[];  // 8 planes of (1 << (k + glog)) words
    const int tlog = k + glog;
    const uint32_t tsize = 1u << tlog;
    const uint32_t G = 1u << glog;
    const size_t tile = blockIdx.x;
    size_t base;
    uint32_t lo0 = 0;
    if (s0 == 0) {
        base = tile << tlog;
    } else {
        size_t tph = ((size_t)1 << s0) >> glog;  // tiles per hi
        size_t hi = tile / tph;
        lo0 = (uint32_t)(tile % tph) << glog;
        base = (hi << (s0 + k)) + lo0;
    }
    // load: local id L -> (mid, g).  s0 == 0: L = g*2^k + mid (contiguous).  s0 > 0: L = mid*G + g.
    for (uint32_t L = threadIdx.x; L < tsize; L += NTT_THREADS) {
        uint32_t mid, g;
        size_t gi;
        if (s0 == 0) {
            g = L >> k;
            mid = L & ((1u << k) - 1);
            gi = base + L;
        } else {
            mid = L >> glog;
            g = L & (G - 1);
            gi = base + ((size_t)mid << s0) + g;
        }
        F a = load_fr(data + gi);
        uint32_t si = (mid << glog) | g;
#pragma unroll
        for (int l = 0; l < 8; ++l) sm[l * tsize + si] = a.v[l];
    }
    __syncthreads();
    const uint32_t half_n = 1u << (log_n - 1);
    const uint32_t nbf = tsize >> 1;
    for (int t = 1; t <= k; ++t) {
        const int s = s0 + t;
        for (uint32_t b = threadIdx.x; b < nbf; b += NTT_THREADS) {
            uint32_t g = b & (G - 1);
            uint32_t r = b >> glog;  // in [0, 2^(k-1))
            uint32_t mid_low = r & ((1u << (t - 1)) - 1);
            uint32_t mid0 = ((r >> (t - 1)) << t) | mid_low;
            uint32_t i0 = (mid0 << glog) | g;
            uint32_t i1 = i0 + ((1u << (t - 1)) << glog);
            uint32_t j = (mid_low << s0) + (s0 ? lo0 + g : 0);
            uint32_t e = j << (log_n - s);  // < n/2
            F x, y;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                x.v[l] = sm[l * tsize + i0];
                y.v[l] = sm[l * tsize + i1];
            }
            F o0, o1;
            // w^0 = 1 needs no product, but only the very first stage has e = 0 for a whole warp: in the stages after it a
            // per-lane shortcut splits the warp (half, a quarter, ... of the lanes idle through the product the others still
            // execute; ncu: 26.5 of 32 lanes active, profiles/r1_ncu_full_ntt_pass_2p24.txt), so the test is block-uniform
            if (s == 1) {
                o0 = x + y;
                o1 = x - y;
            } else if (!inverse) {
                y = y * load_fr_ro(tw + e);
                o0 = x + y;
                o1 = x - y;
            } else {
                // w^-e = -w^(n/2-e) for e > 0; the lanes with e = 0 multiply by w^0 and keep the signs (w^(n/2) is not in the table)
                y = y * load_fr_ro(tw + (e ? half_n - e : 0));
                const F a = x + y, b = x - y;
                o0 = e ? b : a;
                o1 = e ? a : b;
            }
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                sm[l * tsize + i0] = o0.v[l];
                sm[l * tsize + i1] = o1.v[l];
            }
        }
        __syncthreads();
    }
    for (uint32_t L = threadIdx.x; L < tsize; L += NTT_THREADS) {
        uint32_t mid, g;
        size_t gi;
        if (s0 == 0) {
            g = L >> k;
            mid = L & ((1u << k) - 1);
            gi = base + L;
        } else {
            mid = L >> glog;
            g = L & (G - 1);
            gi = base + ((size_t)mid << s0) + g;
        }
        uint32_t si = (mid << glog) | g;
        F a;
#pragma unroll
        for (int l = 0; l < 8; ++l) a.v[l] = sm[l * tsize + si];
        if (do_scale) a = a * scale;
        store_fr(data + gi, a);
    }
}

template <class FrP>
static Fp<FrP> host_domain_gen(int log_n) {
    Fp<FrP> g;
    for (int i = 0; i < 8; ++i) g.v[i] = FrP::ROOT(i);
    for (int i = log_n; i < FrP::TWO_ADICITY; ++i) g = g.sqr();
    return g;
}
template <class FrP>
static Fp<FrP> host_coset_gen() {
    Fp<FrP> g;
    for (int i = 0; i < 8; ++i) g.v[i] = FrP::GEN(i);
    return g;
}

enum TableKind : uint64_t { TBL_TWIDDLE = 1, TBL_COSET_LO = 2, TBL_COSET_HI = 3, TBL_ICOSET_LO = 4, TBL_ICOSET_HI = 5 };
static uint64_t table_key(int curve, uint64_t kind, int log_n) { return ((uint64_t)curve << 32) | (kind << 8) | (uint64_t)log_n; }

template <class F>
static int get_pow_table(zkaes_ctx* ctx, uint64_t key, size_t count, const F& base, const F& c, int shift, const F** out) {
    auto it = ctx->tables.find(key);
    if (it != ctx->tables.end()) {
        *out = reinterpret_cast<const F*>(it->second);
        return ZK_OK;
    }
    void* p = nullptr;
    ZK_CUDA(ctx, cudaMalloc(&p, sizeof(F) * (count ? count : 1)));
    if (count) {
        k_pow_table<F><<<cdiv(count, 256), 256, 0, ctx->stream>>>(reinterpret_cast<F*>(p), count, base, c, shift);
        ctx->launches++;
        ZK_CUDA(ctx, cudaGetLastError());
    }
    ctx->tables[key] = p;
    *out = reinterpret_cast<const F*>(p);
    return ZK_OK;
}

template <class FrP>
int ntt_device(zkaes_ctx* ctx, int curve_id, void* d_data, int log_n, int inverse, int coset) {
    using F = Fp<FrP>;
    if (log_n < 0 || log_n > FrP::TWO_ADICITY || log_n > 30) return fail(ctx, ZK_ERR_ARG, "ntt: log_n out of range");
    cudaStream_t st = ctx->stream;
    F* data = reinterpret_cast<F*>(d_data);
    const size_t n = (size_t)1 << log_n;
    const F w = host_domain_gen<FrP>(log_n);
    const F one = F::one();

    const F* tw = nullptr;
    if (log_n >= 1) ZK_TRY(get_pow_table<F>(ctx, table_key(curve_id, TBL_TWIDDLE, log_n), n / 2, w, one, 0, &tw));

    const F* pw_lo = nullptr;
    const F* pw_hi = nullptr;
    const size_t lo_cnt = n < ((size_t)1 << POW_LO_LOG) ? n : ((size_t)1 << POW_LO_LOG);
    const size_t hi_cnt = n >> POW_LO_LOG;
    if (coset && !inverse) {
        const F g = host_coset_gen<FrP>();
        ZK_TRY(get_pow_table<F>(ctx, table_key(curve_id, TBL_COSET_LO, log_n), lo_cnt, g, one, 0, &pw_lo));
        ZK_TRY(get_pow_table<F>(ctx, table_key(curve_id, TBL_COSET_HI, log_n), hi_cnt, g, one, POW_LO_LOG, &pw_hi));
    }
    if (log_n >= 1 || pw_lo) {
        k_bitrev<F><<<cdiv(n, 256), 256, 0, st>>>(data, log_n, pw_lo, pw_hi);
        ctx->launches++;
    }

    // n^-1 (only for the inverse transform without coset; the coset variant folds it into its power table)
    F ninv = one;
    if (inverse) {
        F nn = F::from_u64((uint64_t)n);
        ninv = nn.inverse();
    }

    if (log_n >= 1) {
        int npass = log_n <= TILE_LOG ? 1 : (log_n + 7) / 8;
        int basek = log_n / npass, rem = log_n % npass;
        int s0 = 0;
        for (int pi = 0; pi < npass; ++pi) {
            int k = basek + (pi < rem ? 1 : 0);
            int glog = TILE_LOG - k;
            if (s0 > 0 && glog > s0) glog = s0;
            if (s0 == 0 && glog > log_n - k) glog = log_n - k;
            size_t tiles = n >> (k + glog);
            size_t smem = (size_t)32 << (k + glog);
            bool last = (pi == npass - 1);
            int do_scale = (inverse && !coset && last) ? 1 : 0;
            ZK_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 << TILE_LOG));
            k_ntt_pass<F><<<(unsigned)tiles, NTT_THREADS, smem, st>>>(data, tw, log_n, s0, k, glog, inverse, do_scale, ninv);
            ctx->launches++;
            s0 += k;
        }
    } else if (inverse) {
        // n = 1: nothing to do (n^-1 = 1)
    }
    if (coset && inverse) {
        const F gi = host_coset_gen<FrP>().inverse();
        const F* ilo = nullptr;
        const F* ihi = nullptr;
        ZK_TRY(get_pow_table<F>(ctx, table_key(curve_id, TBL_ICOSET_LO, log_n), lo_cnt, gi, ninv, 0, &ilo));
        ZK_TRY(get_pow_table<F>(ctx, table_key(curve_id, TBL_ICOSET_HI, log_n), hi_cnt, gi, one, POW_LO_LOG, &ihi));
        k_scale_pow<F><<<cdiv(n, 256), 256, 0, st>>>(data, n, ilo, ihi);
        ctx->launches++;
    }
    ZK_CUDA(ctx, cudaGetLastError());
    return ZK_OK;
}

template int ntt_device<Fr377Params>(zkaes_ctx*, int, void*, int, int, int);
template int ntt_device<Fr381Params>(zkaes_ctx*, int, void*, int, int, int);

}  // namespace zk
