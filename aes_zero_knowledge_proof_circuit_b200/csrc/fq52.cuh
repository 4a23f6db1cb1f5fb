// Fq on the FP64 pipe: 8 limbs of 52 bits held in doubles (exact integers), Montgomery product from DFMA hi/lo limb products.
//
// EXPERIMENT (round 2) -- used by tools/ubench.cu and the CPU-tier test tests/test_fq52.py, not by the prover.
// B200 issues 58.8 DFMA per clock per SM next to 31.1 IMAD.WIDE (profiles/ubench_r2.txt): the two pipes overlap, so FP64-limb
// warps next to IMAD.WIDE warps could add throughput.  This file is the product such warps would run, bit-exact against the
// 32-bit-limb multiplier (same Montgomery form is NOT kept here: R = 2^416; see DESIGN.md 4.9 for what a kernel would still need).
//
// Limb product (Emmart et al., "Faster modular exponentiation using double precision floating point arithmetic on the GPU"):
//   hi = fma_rz(x, y, 2^104)                  = 2^104 + 2^52 floor(xy / 2^52)        mantissa field = floor(xy / 2^52)
//   lo = fma_rz(x, y, (2^104 + 2^52) - hi)    = 2^52 + (xy mod 2^52)                 mantissa field = xy mod 2^52
// The raw bit patterns are ADDED as 64-bit integers into column sums; the exponent fields they drag along are known per column and
// subtracted when a column is finalised.
#pragma once
#include <cstdint>
#include <cstring>
#if !defined(__CUDA_ARCH__)
#include <cmath>
#endif

namespace zk {

struct Fq52 {
    double v[8];
};

#if defined(__CUDA_ARCH__)
#define ZK52_HD __host__ __device__ __forceinline__
#define ZK52_FMA_RZ(x, y, z) __fma_rz((x), (y), (z))
#define ZK52_BITS(d) ((uint64_t)__double_as_longlong(d))
#define ZK52_DBL(u) __longlong_as_double((long long)(u))
#else
#if defined(__CUDACC__)
#define ZK52_HD __host__ __device__ inline
#else
#define ZK52_HD inline
#endif
// host: the caller runs under fesetround(FE_TOWARDZERO); std::fma is correctly rounded in the current mode
#define ZK52_FMA_RZ(x, y, z) std::fma((x), (y), (z))
static inline uint64_t zk52_bits(double d) {
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
}
static inline double zk52_dbl(uint64_t u) {
    double d;
    memcpy(&d, &u, 8);
    return d;
}
#define ZK52_BITS(d) zk52_bits(d)
#define ZK52_DBL(u) zk52_dbl(u)
#endif

static constexpr uint64_t ZK52_MASK = ((uint64_t)1 << 52) - 1;
static constexpr uint64_t ZK52_EXP_HI = (uint64_t)(1023 + 104) << 52;  // bit pattern of 2^104
static constexpr uint64_t ZK52_EXP_LO = (uint64_t)(1023 + 52) << 52;   // bit pattern of 2^52

// P52: the modulus as 8 x 52-bit limbs; INV52 = -p^-1 mod 2^52
template <class P52>
ZK52_HD Fq52 fq52_mont_mul(const Fq52& a, const Fq52& b) {
    const double C1 = 0x1p104, C2 = 0x1p104 + 0x1p52;
    uint64_t col[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) col[k] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double hi = ZK52_FMA_RZ(a.v[i], b.v[j], C1);
            const double lo = ZK52_FMA_RZ(a.v[i], b.v[j], C2 - hi);
            col[i + j] += ZK52_BITS(lo);
            col[i + j + 1] += ZK52_BITS(hi);
        }
        // column i is complete: (i + 1) + i low parts and i + i high parts have been added to it (rows 0..i of a*b, rows 0..i-1 of q*p)
        const uint64_t c = col[i] - (uint64_t)(2 * i + 1) * ZK52_EXP_LO - (uint64_t)(2 * i) * ZK52_EXP_HI;
        const uint64_t q = ((c & ZK52_MASK) * P52::INV52) & ZK52_MASK;
        const double qd = ZK52_DBL(q | ZK52_EXP_LO) - 0x1p52;
        uint64_t lo0 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double hi = ZK52_FMA_RZ(qd, P52::limb(j), C1);
            const double lo = ZK52_FMA_RZ(qd, P52::limb(j), C2 - hi);
            if (j == 0)
                lo0 = ZK52_BITS(lo) - ZK52_EXP_LO;
            else
                col[i + j] += ZK52_BITS(lo);
            col[i + j + 1] += ZK52_BITS(hi);
        }
        col[i + 1] += (c + lo0) >> 52;  // c + lo(q p_0) = 0 mod 2^52: only its carry survives
    }
    // columns 8..15: 2 (15 - k) low parts and 2 (16 - k) high parts each, plus the carries pushed into them
    uint64_t r[8], carry = 0;
#pragma unroll
    for (int k = 8; k < 16; ++k) {
        const uint64_t v = col[k] - (uint64_t)(2 * (15 - k)) * ZK52_EXP_LO - (uint64_t)(2 * (16 - k)) * ZK52_EXP_HI + carry;
        r[k - 8] = v & ZK52_MASK;
        carry = v >> 52;
    }
    // result < 2p: one conditional subtraction of p
    uint64_t t[8];
    int64_t borrow = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t d = (int64_t)r[k] - (int64_t)P52::ilimb(k) + borrow;
        t[k] = (uint64_t)d & ZK52_MASK;
        borrow = d >> 52;  // 0 or -1
    }
    Fq52 out;
#pragma unroll
    for (int k = 0; k < 8; ++k) out.v[k] = ZK52_DBL((borrow ? r[k] : t[k]) | ZK52_EXP_LO) - 0x1p52;
    return out;
}

// BLS12-377 base field q as 8 x 52-bit limbs (little-endian), -q^-1 mod 2^52
struct Fq377P52 {
    static constexpr uint64_t INV52 = 0x8bfffffffffffull;
    static ZK52_HD uint64_t ilimb(int j) {
        constexpr uint64_t L[8] = {0x8c00000000001ull, 0x4430000000850ull, 0xa094800170b5dull, 0x138f1ef3622fbull, 0xb1a22d9f300f5ull, 0x3b05c06ca1493ull, 0xa4617c510eac6ull, 0x1ae3ull};
        return L[j];
    }
    static ZK52_HD double limb(int j) { return (double)ilimb(j); }
};

// 32-bit-limb value (12 words, little-endian) <-> 52-bit limbs: pure bit re-slicing
ZK52_HD Fq52 fq52_from_words(const uint32_t* w) {
    Fq52 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int bit = 52 * k, lo = bit >> 5, sh = bit & 31;
        uint64_t v = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t)
            if (lo + t < 12) {
                const int s = 32 * t - sh;  // position of word lo + t relative to the limb's bit 0
                if (s < 52) v |= s >= 0 ? (uint64_t)w[lo + t] << s : (uint64_t)w[lo + t] >> -s;
            }
        r.v[k] = ZK52_DBL((v & ZK52_MASK) | ZK52_EXP_LO) - 0x1p52;
    }
    return r;
}
ZK52_HD void fq52_to_words(const Fq52& a, uint32_t* w) {
    uint64_t l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) l[k] = ZK52_BITS(a.v[k] + 0x1p52) & ZK52_MASK;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int bit = 32 * i, k = bit / 52, sh = bit % 52;
        uint64_t v = l[k] >> sh;
        if (sh > 20 && k + 1 < 8) v |= l[k + 1] << (52 - sh);
        w[i] = (uint32_t)v;
    }
}

}  // namespace zk
