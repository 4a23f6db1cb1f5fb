// Prime-field arithmetic for the B200 prover: Fp<P> over 32-bit limbs in Montgomery form.
//
// Replaces, on the hot path, ark-ff 0.3.0's Fp256/Fp384 ([u64;4]/[u64;6] Montgomery; reference
// Cargo.lock:159-186).  The in-memory layout is identical to arkworks' (little-endian limbs of the
// Montgomery representative, R = 2^256 / 2^384), so buffers cross the C-ABI without conversion.
//
// Two multipliers:
//   * mont_mul_portable : CIOS on uint64_t, compiles for host and device (host side is unit-tested on CPU)
//   * mont_mul_raw_{8,12}: generated inline PTX (tools/gen_mont_asm.py), even/odd split carry chains so
//     every 32x32->64 product is one IMAD.WIDE feeding a carry chain; device only.
// ZK_FF_PORTABLE forces the portable multiplier on device (used by the on-GPU self test to cross-check).
#pragma once
#include <cstdint>
#include "params_gen.h"

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_DEV __device__ __forceinline__
#else
#define ZK_HD inline
#define ZK_DEV inline
#endif

namespace zk {

#if defined(__CUDACC__)
#include "ff_mont_asm.inc"
// -p^-1 mod 2^32 per field, read from constant memory on the device ON PURPOSE: when ptxas sees the literal
// (0xffffffff for three of the four fields) it rewrites m = -t0 and then splits every IMAD.WIDE of the reduction rows
// into IMAD + IMAD.HI (half rate on sm_100): 11.7 -> ~5 clk/SM per Fq product (profiles/ubench_r1.txt).
static __constant__ uint32_t zk_c_mont_inv[4] = {Fr377Params::INV, Fq377Params::INV, Fr381Params::INV, Fq381Params::INV};
// An opaque zero (ptxas cannot fold a constant-bank load) used by Fp::mul_after to order two otherwise independent
// products: without it ptxas interleaves their carry chains, runs out of the 7 predicate registers and spills carries
// through P2R/LOP3/ISETP (about one extra ALU instruction per IMAD.WIDE in the madd inner loop).
static __constant__ uint32_t zk_c_zero = 0;
#endif

template <int N>
ZK_HD uint32_t add_n(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
}

// r = a - b, returns borrow (0/1)
template <int N>
ZK_HD uint32_t sub_n(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        c += (int64_t)a[i] - (int64_t)b[i];
        r[i] = (uint32_t)c;
        c >>= 32;  // arithmetic shift: 0 or -1
    }
    return (uint32_t)(c & 1);
}

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t v[N];

    static ZK_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = 0;
        return r;
    }
    static ZK_HD Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = P::ONE(i);
        return r;
    }
    static ZK_HD Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = P::R2(i);
        return r;
    }
    ZK_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= v[i];
        return o == 0;
    }
    ZK_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    ZK_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    // if v >= p: v -= p   (v < 2p on entry)
    ZK_HD void reduce_once() {
        uint32_t t[N], m[N];
#pragma unroll
        for (int i = 0; i < N; ++i) m[i] = P::MOD(i);
        uint32_t borrow = sub_n<N>(t, v, m);
        if (!borrow) {
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = t[i];
        }
    }

    friend ZK_HD Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        add_n<N>(r.v, a.v, b.v);  // p < 2^(32N-1) for all four fields: no carry out
        r.reduce_once();
        return r;
    }
    friend ZK_HD Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t borrow = sub_n<N>(r.v, a.v, b.v);
        if (borrow) {
            uint32_t m[N];
#pragma unroll
            for (int i = 0; i < N; ++i) m[i] = P::MOD(i);
            add_n<N>(r.v, r.v, m);
        }
        return r;
    }
    ZK_HD Fp neg() const {
        if (is_zero()) return *this;
        Fp r;
        uint32_t m[N];
#pragma unroll
        for (int i = 0; i < N; ++i) m[i] = P::MOD(i);
        sub_n<N>(r.v, m, v);
        return r;
    }
    ZK_HD Fp dbl() const { return *this + *this; }

    static ZK_HD void mont_mul_portable(uint32_t* r, const uint32_t* a, const uint32_t* b) {
        // CIOS, one extra limb of headroom; result < 2p before the final subtract
        uint32_t t[N + 2];
#pragma unroll
        for (int i = 0; i < N + 2; ++i) t[i] = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                c += (uint64_t)a[j] * b[i] + t[j];
                t[j] = (uint32_t)c;
                c >>= 32;
            }
            c += t[N];
            t[N] = (uint32_t)c;
            t[N + 1] = (uint32_t)(c >> 32);
            uint32_t m = t[0] * P::INV;
            c = (uint64_t)m * P::MOD(0) + t[0];
            c >>= 32;
#pragma unroll
            for (int j = 1; j < N; ++j) {
                c += (uint64_t)m * P::MOD(j) + t[j];
                t[j - 1] = (uint32_t)c;
                c >>= 32;
            }
            c += t[N];
            t[N - 1] = (uint32_t)c;
            t[N] = t[N + 1] + (uint32_t)(c >> 32);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = t[i];
    }

#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__)
    // Host-only CIOS over 64-bit limbs (the same little-endian memory as the 32-bit limbs): ~4x the portable loop.  Used by
    // everything the host computes -- the MSM Horner fold, hiding commitments, the verifier's pairing (csrc/pairing.h).
    static inline uint64_t host_inv64() {
        // -p^-1 mod 2^64 from -p^-1 mod 2^32 by one Newton step: x' = x (2 - p x)
        const uint64_t p0 = (uint64_t)P::MOD(0) | ((uint64_t)P::MOD(1) << 32);
        uint64_t x = (uint64_t)(0u - P::INV);  // p^-1 mod 2^32
        x = x * (2 - p0 * x);                  // mod 2^64
        return 0 - x;
    }
    static inline void mont_mul_host64(uint32_t* r32, const uint32_t* a32, const uint32_t* b32) {
        constexpr int M = N / 2;
        static_assert(N % 2 == 0, "64-bit host path needs an even number of 32-bit limbs");
        uint64_t a[M], b[M], p[M], t[M + 2];
        for (int i = 0; i < M; ++i) {
            a[i] = (uint64_t)a32[2 * i] | ((uint64_t)a32[2 * i + 1] << 32);
            b[i] = (uint64_t)b32[2 * i] | ((uint64_t)b32[2 * i + 1] << 32);
            p[i] = (uint64_t)P::MOD(2 * i) | ((uint64_t)P::MOD(2 * i + 1) << 32);
        }
        static const uint64_t inv = host_inv64();
        for (int i = 0; i < M + 2; ++i) t[i] = 0;
        for (int i = 0; i < M; ++i) {
            unsigned __int128 c = 0;
            for (int j = 0; j < M; ++j) {
                c += (unsigned __int128)a[j] * b[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[M];
            t[M] = (uint64_t)c;
            t[M + 1] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * inv;
            c = (unsigned __int128)m * p[0] + t[0];
            c >>= 64;
            for (int j = 1; j < M; ++j) {
                c += (unsigned __int128)m * p[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[M];
            t[M - 1] = (uint64_t)c;
            t[M] = t[M + 1] + (uint64_t)(c >> 64);
        }
        for (int i = 0; i < M; ++i) {
            r32[2 * i] = (uint32_t)t[i];
            r32[2 * i + 1] = (uint32_t)(t[i] >> 32);
        }
    }
#endif

    friend ZK_HD Fp operator*(const Fp& a, const Fp& b) {
        Fp r;
#if defined(__CUDA_ARCH__) && !defined(ZK_FF_PORTABLE)
        uint32_t m[N];
#pragma unroll
        for (int i = 0; i < N; ++i) m[i] = P::MOD(i);
        const uint32_t inv = zk_c_mont_inv[P::FIELD_ID];
        if constexpr (N == 8)
            mont_mul_raw_8(r.v, a.v, b.v, m, inv, zk_c_zero);
        else
            mont_mul_raw_12(r.v, a.v, b.v, m, inv, zk_c_zero);
#elif !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__) && !defined(ZK_FF_PORTABLE)
        mont_mul_host64(r.v, a.v, b.v);
#else
        mont_mul_portable(r.v, a.v, b.v);
#endif
        r.reduce_once();
        return r;
    }
    // Squaring.  Device, 12 limbs: the generated dedicated squaring (tools/gen_mont_asm.py sqr_rows_for: off-diagonal products once and
    // doubled, then a separate reduction -- 222 wide multiplies instead of 288).  Everything else: a product.
    ZK_HD Fp sqr() const {
#if defined(__CUDA_ARCH__) && !defined(ZK_FF_PORTABLE)
        if constexpr (N == 12) {
            Fp r;
            uint32_t m[N];
#pragma unroll
            for (int i = 0; i < N; ++i) m[i] = P::MOD(i);
            mont_sqr_raw_12(r.v, v, m, zk_c_mont_inv[P::FIELD_ID], zk_c_zero);
            r.reduce_once();
            return r;
        }
#endif
        return *this * *this;
    }

    // a * b, scheduled after `tok` was produced (device: a false data dependency on b[0]; host: plain product).
    // `tok` is then replaced by a limb of the result so that calls chain.
    static ZK_HD Fp mul_after(const Fp& a, const Fp& b, uint32_t& tok) {
#if defined(__CUDA_ARCH__) && !defined(ZK_FF_PORTABLE)
        Fp bb = b;
        bb.v[0] |= tok & zk_c_zero;
        Fp r = a * bb;
        tok = r.v[N - 1];
        return r;
#else
        (void)tok;
        return a * b;
#endif
    }

    // a * b through ONE out-of-line copy of the multiplier (device: a real CALL with both operands and the result in registers --
    // nvcc passes the 2 x N words by value in registers, no stack traffic; checked in SASS, profiles/r2_sass_madd_call.txt).
    // The MSM inner loop issues its ten products through this so that the loop body is ~10 KB of SASS instead of the 84 KB of ten
    // inlined copies, which thrashed the 32 KB instruction cache (20 % "no instruction" stalls in the round-1 ncu capture).
    static ZK_HD Fp mul_call(const Fp& a, const Fp& b);
    static ZK_HD Fp sqr_call(const Fp& a);  // the same for sqr(): one out-of-line copy of the dedicated squaring

    // wire-format hooks: the arkworks limbs ARE the working form
    static ZK_HD Fp unpack(const uint32_t* w) {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = w[i];
        return r;
    }
    ZK_HD void pack(uint32_t* w) const {
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = v[i];
    }
    static ZK_HD Fp from_std(const uint32_t* w) { return unpack(w); }
    ZK_HD void to_std(uint32_t* w) const { pack(w); }

    // canonical integer -> Montgomery and back
    ZK_HD Fp to_mont() const { return *this * r2(); }
    ZK_HD Fp from_mont() const {
        Fp o = zero();
        o.v[0] = 1;
        return *this * o;
    }

    // generic square-and-multiply with a multi-limb exponent (LE 32-bit limbs)
    ZK_HD Fp pow(const uint32_t* e, int nlimbs) const {
        Fp r = one();
        bool started = false;
        for (int i = nlimbs - 1; i >= 0; --i) {
            for (int b = 31; b >= 0; --b) {
                if (started) r = r.sqr();
                if ((e[i] >> b) & 1) {
                    r = started ? r * *this : *this;
                    started = true;
                }
            }
        }
        return r;
    }
    // Fermat inverse (0 -> 0).  Only used off the inner loops (batch inversion amortises it).
    ZK_HD Fp inverse() const {
        uint32_t e[N], two[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            e[i] = P::MOD(i);
            two[i] = (i == 0) ? 2u : 0u;
        }
        sub_n<N>(e, e, two);  // p - 2 (BLS12-377 moduli end in ...0001, so the borrow must ripple)
        return pow(e, N);
    }
    static ZK_HD Fp from_u64(uint64_t x) {
        Fp r = zero();
        r.v[0] = (uint32_t)x;
        r.v[1] = (uint32_t)(x >> 32);
        return r.to_mont();
    }
    // lexicographic compare of canonical values: used for ark-serialize's y-sign flag
    ZK_HD bool canonical_gt(const Fp& b) const {
        for (int i = N - 1; i >= 0; --i) {
            if (v[i] != b.v[i]) return v[i] > b.v[i];
        }
        return false;
    }
};

#if defined(__CUDACC__)
template <class P>
struct FpWords {
    uint32_t v[P::N];
};
template <class P>
__device__ __noinline__ FpWords<P> fp_mul_out_of_line(FpWords<P> a, FpWords<P> b) {
    Fp<P> x, y;
#pragma unroll
    for (int i = 0; i < P::N; ++i) {
        x.v[i] = a.v[i];
        y.v[i] = b.v[i];
    }
    const Fp<P> r = x * y;
    FpWords<P> o;
#pragma unroll
    for (int i = 0; i < P::N; ++i) o.v[i] = r.v[i];
    return o;
}
#endif
#if defined(__CUDACC__)
template <class P>
__device__ __noinline__ FpWords<P> fp_sqr_out_of_line(FpWords<P> a) {
    Fp<P> x;
#pragma unroll
    for (int i = 0; i < P::N; ++i) x.v[i] = a.v[i];
    const Fp<P> r = x.sqr();
    FpWords<P> o;
#pragma unroll
    for (int i = 0; i < P::N; ++i) o.v[i] = r.v[i];
    return o;
}
#endif
template <class P>
ZK_HD Fp<P> Fp<P>::sqr_call(const Fp<P>& a) {
#if defined(__CUDA_ARCH__) && !defined(ZK_FF_PORTABLE)
    FpWords<P> x;
#pragma unroll
    for (int i = 0; i < P::N; ++i) x.v[i] = a.v[i];
    const FpWords<P> o = fp_sqr_out_of_line<P>(x);
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; ++i) r.v[i] = o.v[i];
    return r;
#else
    return a * a;
#endif
}
template <class P>
ZK_HD Fp<P> Fp<P>::mul_call(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__) && !defined(ZK_FF_PORTABLE)
    FpWords<P> x, y;
#pragma unroll
    for (int i = 0; i < P::N; ++i) {
        x.v[i] = a.v[i];
        y.v[i] = b.v[i];
    }
    const FpWords<P> o = fp_mul_out_of_line<P>(x, y);
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; ++i) r.v[i] = o.v[i];
    return r;
#else
    return a * b;
#endif
}

using Fr377 = Fp<Fr377Params>;
using Fq377 = Fp<Fq377Params>;
using Fr381 = Fp<Fr381Params>;
using Fq381 = Fp<Fq381Params>;

}  // namespace zk
