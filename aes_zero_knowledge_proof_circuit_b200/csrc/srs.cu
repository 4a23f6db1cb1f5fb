// Test-SRS generation on the device: out[i] = tau^i * G (fixed-base scalar multiplication).
//
// Stands in for ark-poly-commit 0.3.0 `KZG10::setup` (powers_of_g), which the reference reaches through
// simpleworks::marlin::generate_universal_srs (src/lib.rs:141).  As in the reference (README.md:26) the
// trapdoor is derived from a seed and is NOT secret: this is a test SRS.
//
// Fixed-base method: 8-bit windows, table T[w][d-1] = d * 2^(8w) * G (32 x 255 affine points, built on the
// device), each output = sum_w T[w][byte_w(tau^i)] with 32 mixed additions, then normalised to affine.
// [r2] The normalisation is batched: a Fermat inversion per point (~570 products) cost more than the 32 mixed additions
// (320 products).  The multiplication kernels now leave XYZZ sums in a chunk buffer and k_fb_normalise runs Montgomery's trick
// over runs of FB_RUN consecutive points (3 products per point + one inversion per run), x = X (ZZ / ZZZ)^2, y = Y / ZZZ.
#include "srs.cuh"

namespace zk {

static constexpr int FB_WINDOWS = 32;
static constexpr int FB_ROW = 255;

template <class C>
__global__ void k_fb_rows(Affine<C>* rows) {  // rows[w] = 2^(8w) * G
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= FB_WINDOWS) return;
    XYZZ<C> p = XYZZ<C>::from_affine(Affine<C>::generator());
    for (int i = 0; i < 8 * w; ++i) p = p.dbl();
    rows[w] = p.to_affine();
}
template <class C>
__global__ void k_fb_fill(const Affine<C>* __restrict__ rows, Affine<C>* __restrict__ table) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= FB_WINDOWS * FB_ROW) return;
    int w = t / FB_ROW, d = t % FB_ROW + 1;
    XYZZ<C> base = XYZZ<C>::from_affine(rows[w]);
    XYZZ<C> r = XYZZ<C>::inf();
    for (int b = 7; b >= 0; --b) {
        r = r.dbl();
        if ((d >> b) & 1) r.add(base);
    }
    table[t] = r.to_affine();
}

// scalar_i = lo[i & 1023] * hi[i >> 10]  (Montgomery), converted to canonical bytes
template <class C>
__global__ void __launch_bounds__(128) k_fb_mul(const Affine<C>* __restrict__ table, const Fp<typename C::FrP>* __restrict__ pw_lo,
                                               const Fp<typename C::FrP>* __restrict__ pw_hi, size_t start, size_t stride, size_t n,
                                               XYZZ<C>* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    using Fr = Fp<typename C::FrP>;
    const size_t e = start + i * stride;  // exponent of tau
    Fr s = pw_lo[e & 1023];
    if (e >> 10) s = s * pw_hi[e >> 10];
    s = s.from_mont();
    XYZZ<C> acc = XYZZ<C>::inf();
    for (int w = 0; w < FB_WINDOWS; ++w) {
        uint32_t d = (s.v[w >> 2] >> (8 * (w & 3))) & 0xff;
        if (d) acc.madd(table[w * FB_ROW + d - 1]);
    }
    out[i] = acc;
}

// out[i] = scalars[start + i * stride] * G for Montgomery-form Fr scalars in HBM (the Lagrange-basis points of the prover key)
template <class C>
__global__ void __launch_bounds__(128) k_fb_mul_vec(const Affine<C>* __restrict__ table, const Fp<typename C::FrP>* __restrict__ scalars, size_t start,
                                                   size_t stride, size_t n, XYZZ<C>* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    using Fr = Fp<typename C::FrP>;
    Fr s = scalars[start + i * stride].from_mont();
    XYZZ<C> acc = XYZZ<C>::inf();
    for (int w = 0; w < FB_WINDOWS; ++w) {
        uint32_t d = (s.v[w >> 2] >> (8 * (w & 3))) & 0xff;
        if (d) acc.madd(table[w * FB_ROW + d - 1]);
    }
    out[i] = acc;
}

// XYZZ -> affine for runs of FB_RUN consecutive points per thread with ONE inversion per run (Montgomery's trick on the ZZZ coordinates;
// the prefix products are parked in the x slot of the output).  Points at infinity (ZZ = ZZZ = 0) are skipped and come out as (0, 0).
static constexpr int FB_RUN = 32;
template <class C>
__global__ void __launch_bounds__(128) k_fb_normalise(const XYZZ<C>* __restrict__ in, size_t n, Affine<C>* __restrict__ out) {
    using Fq = typename Affine<C>::Fq;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = t * FB_RUN;
    if (lo >= n) return;
    const size_t hi = lo + FB_RUN < n ? lo + FB_RUN : n;
    Fq acc = Fq::one();
    for (size_t i = lo; i < hi; ++i) {
        out[i].x = acc;
        const Fq z = in[i].zzz;
        if (!z.is_zero()) acc = acc * z;
    }
    Fq inv = acc.inverse();
    for (size_t i = hi; i-- > lo;) {
        const XYZZ<C> p = in[i];
        if (p.zzz.is_zero()) {
            out[i] = Affine<C>::inf();
            continue;
        }
        const Fq izzz = inv * out[i].x;   // 1 / ZZZ_i
        inv = inv * p.zzz;
        const Fq izz = (izzz * p.zz).sqr();  // (ZZ / ZZZ)^2 = 1 / ZZ   because ZZ^3 = ZZZ^2
        Affine<C> r;
        r.x = p.x * izz;
        r.y = p.y * izzz;
        out[i] = r;
    }
}
// chunked: `mul(first, count, tmp)` launches a multiplication kernel that writes the XYZZ sums of outputs [first, first + count) to tmp
static constexpr size_t FB_CHUNK = (size_t)1 << 22;  // 805 MB of XYZZ sums
template <class C, class Mul>
static int fb_mul_chunks(zkaes_ctx* ctx, size_t n, Affine<C>* out, Mul mul) {
    if (!n) return ZK_OK;
    DevBuf tmp;
    ZK_CUDA(ctx, tmp.alloc(sizeof(XYZZ<C>) * (n < FB_CHUNK ? n : FB_CHUNK), ctx->stream));
    for (size_t first = 0; first < n; first += FB_CHUNK) {
        const size_t count = n - first < FB_CHUNK ? n - first : FB_CHUNK;
        mul(first, count, tmp.as<XYZZ<C>>());
        k_fb_normalise<C><<<cdiv(cdiv(count, FB_RUN), 128), 128, 0, ctx->stream>>>(tmp.as<XYZZ<C>>(), count, out + first);
        ctx->launches += 2;
    }
    ZK_CUDA(ctx, cudaGetLastError());
    return ZK_OK;
}

template <class F>
__global__ void k_pow_table_srs(F* out, size_t count, F base, int shift) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long e = (unsigned long long)i << shift;
    F r = F::one(), b = base;
    while (e) {
        if (e & 1) r = r * b;
        b = b.sqr();
        e >>= 1;
    }
    out[i] = r;
}

// the 32 x 255 table of d * 2^(8w) * G
template <class C>
static int fb_table_build(zkaes_ctx* ctx, DevBuf& table) {
    cudaStream_t st = ctx->stream;
    DevBuf rows;
    ZK_CUDA(ctx, rows.alloc(sizeof(Affine<C>) * FB_WINDOWS, st));
    ZK_CUDA(ctx, table.alloc(sizeof(Affine<C>) * FB_WINDOWS * FB_ROW, st));
    k_fb_rows<C><<<1, FB_WINDOWS, 0, st>>>(rows.as<Affine<C>>());
    k_fb_fill<C><<<cdiv(FB_WINDOWS * FB_ROW, 64), 64, 0, st>>>(rows.as<Affine<C>>(), table.as<Affine<C>>());
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaGetLastError());
    return ZK_OK;
}

template <class C>
int srs_powers_device(zkaes_ctx* ctx, const uint8_t seed32[32], size_t n, void* d_out, size_t start, size_t stride) {
    using Fr = Fp<typename C::FrP>;
    cudaStream_t st = ctx->stream;
    // tau: seed as a 252-bit integer (always < r for both curves), lifted to Montgomery form
    Fr tau = Fr::zero();
    for (int i = 0; i < 32; ++i) tau.v[i >> 2] |= (uint32_t)seed32[i] << (8 * (i & 3));
    tau.v[7] &= 0x0fffffffu;
    tau = tau.to_mont();
    DevBuf table, lo, hi;
    ZK_TRY(fb_table_build<C>(ctx, table));
    size_t hi_cnt = ((start + n * stride) >> 10) + 1;
    ZK_CUDA(ctx, lo.alloc(sizeof(Fr) * 1024, st));
    ZK_CUDA(ctx, hi.alloc(sizeof(Fr) * hi_cnt, st));
    k_pow_table_srs<Fr><<<4, 256, 0, st>>>(lo.as<Fr>(), 1024, tau, 0);
    k_pow_table_srs<Fr><<<cdiv(hi_cnt, 256), 256, 0, st>>>(hi.as<Fr>(), hi_cnt, tau, 10);
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaGetLastError());
    return fb_mul_chunks<C>(ctx, n, reinterpret_cast<Affine<C>*>(d_out), [&](size_t first, size_t count, XYZZ<C>* tmp) {
        k_fb_mul<C><<<cdiv(count, 128), 128, 0, st>>>(table.as<Affine<C>>(), lo.as<Fr>(), hi.as<Fr>(), start + first * stride, stride, count, tmp);
    });
}

template <class C>
int fb_mul_scalars_device(zkaes_ctx* ctx, const void* d_scalars, size_t n, void* d_out, size_t start, size_t stride) {
    using Fr = Fp<typename C::FrP>;
    DevBuf table;
    ZK_TRY(fb_table_build<C>(ctx, table));
    return fb_mul_chunks<C>(ctx, n, reinterpret_cast<Affine<C>*>(d_out), [&](size_t first, size_t count, XYZZ<C>* tmp) {
        k_fb_mul_vec<C><<<cdiv(count, 128), 128, 0, ctx->stream>>>(table.as<Affine<C>>(), reinterpret_cast<const Fr*>(d_scalars), start + first * stride, stride,
                                                                   count, tmp);
    });
}
template int fb_mul_scalars_device<G1_377Params>(zkaes_ctx*, const void*, size_t, void*, size_t, size_t);
template int fb_mul_scalars_device<G1_381Params>(zkaes_ctx*, const void*, size_t, void*, size_t, size_t);

template int srs_powers_device<G1_377Params>(zkaes_ctx*, const uint8_t*, size_t, void*, size_t, size_t);
template int srs_powers_device<G1_381Params>(zkaes_ctx*, const uint8_t*, size_t, void*, size_t, size_t);

}  // namespace zk
