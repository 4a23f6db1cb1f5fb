// Base-field arithmetic in radix 2^29 for the MSM kernels (device; the host build exists for unit tests).
//
// Same mathematical object as Fp<FqP> in ff.cuh (ark-ff 0.3.0's Fp384; reference Cargo.lock:159-186), different
// machine schedule.  Measured on B200 (profiles/ubench_r1.txt): IMAD.WIDE.U32 issues at 63.5 /clk/SM but the carry-flag
// form IMAD.WIDE.U32.X that a saturated 32-bit-limb Montgomery product needs issues at 31.6 /clk/SM.  With 29-bit limbs
// a 64-bit accumulator absorbs all 2 x NL partial products of a column without overflowing, so the whole product runs on
// plain IMAD.WIDE; carries are resolved once per row / once at the end with shifts and adds on the ALU pipe, which has
// twice the issue rate and is otherwise idle.  Cost per product: NL^2 + NL(NL+1) IMAD vs 2 x 12^2 half-rate IMAD.WIDE.X.
//
// Invariant: every Fq29 value is CANONICAL (limbs < 2^29, value < p), Montgomery form with R' = 2^(29 NL).
// Conversion from / to arkworks' wire form (12 x u32 limbs, R = 2^384) is one product each (from_std / to_std); bases that
// stay resident (the SRS) are kept "packed": the R' representative as a plain little-endian integer in the same 48 bytes.
#pragma once
#include "ff.cuh"

namespace zk {

template <class P>
struct Fq29 {
    static constexpr int NL = P::NL;
    static constexpr int LB = P::LB;
    static constexpr uint32_t MASK = (1u << LB) - 1;
    uint32_t v[NL];

    static ZK_HD Fq29 zero() {
        Fq29 r;
#pragma unroll
        for (int i = 0; i < NL; ++i) r.v[i] = 0;
        return r;
    }
    static ZK_HD Fq29 one() {
        Fq29 r;
#pragma unroll
        for (int i = 0; i < NL; ++i) r.v[i] = P::ONE(i);
        return r;
    }
    ZK_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) o |= v[i];
        return o == 0;
    }
    ZK_HD bool operator==(const Fq29& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) o |= v[i] ^ b.v[i];
        return o == 0;
    }

    // t (limbs < 2^29 except possibly the top one, value < 2p) -> canonical
    static ZK_HD Fq29 cond_sub(const uint32_t* t) {
        Fq29 d;
        int32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            int32_t x = (int32_t)t[i] - (int32_t)P::MOD(i) + borrow;
            if (i < NL - 1) {
                d.v[i] = (uint32_t)x & MASK;
                borrow = x >> LB;  // arithmetic: 0 or -1
            } else {
                d.v[i] = (uint32_t)x;
                borrow = x >> 31;
            }
        }
        Fq29 r;
#pragma unroll
        for (int i = 0; i < NL; ++i) r.v[i] = borrow ? t[i] : d.v[i];
        return r;
    }

    friend ZK_HD Fq29 operator+(const Fq29& a, const Fq29& b) {
        uint32_t s[NL];
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            uint32_t x = a.v[i] + b.v[i] + c;
            if (i < NL - 1) {
                s[i] = x & MASK;
                c = x >> LB;
            } else {
                s[i] = x;  // < 2p: may use bit 29 of the top limb
            }
        }
        return cond_sub(s);
    }
    friend ZK_HD Fq29 operator-(const Fq29& a, const Fq29& b) {
        // d = a - b; if negative add p back
        uint32_t d[NL];
        int32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            int32_t x = (int32_t)a.v[i] - (int32_t)b.v[i] + borrow;
            if (i < NL - 1) {
                d[i] = (uint32_t)x & MASK;
                borrow = x >> LB;
            } else {
                d[i] = (uint32_t)x;
                borrow = x >> 31;
            }
        }
        const uint32_t m = (uint32_t)borrow;  // all ones if negative
        Fq29 r;
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            uint32_t x = d[i] + (P::MOD(i) & m) + c;
            if (i < NL - 1) {
                r.v[i] = x & MASK;
                c = x >> LB;
            } else {
                r.v[i] = x & MASK;  // wraps the borrowed top limb back into range
            }
        }
        return r;
    }
    ZK_HD Fq29 neg() const {
        if (is_zero()) return *this;
        Fq29 r;
        int32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            int32_t x = (int32_t)P::MOD(i) - (int32_t)v[i] + borrow;
            r.v[i] = (uint32_t)x & MASK;
            borrow = x >> LB;
        }
        return r;
    }
    ZK_HD Fq29 dbl() const { return *this + *this; }

    static ZK_HD void mad_wide(uint64_t& acc, uint32_t x, uint32_t y) {
#if defined(__CUDA_ARCH__)
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(y));
#else
        acc += (uint64_t)x * y;
#endif
    }
    // Montgomery product a * b / 2^(29 NL) mod p: operand scanning with the reduction row interleaved; all partial
    // products accumulate in 64-bit registers through mad.wide.u32 (no carry flags).
    friend ZK_HD Fq29 operator*(const Fq29& a, const Fq29& b) {
        uint64_t t[NL];
#if defined(__CUDA_ARCH__)
        const uint32_t inv = zk_c_r29_inv[P::NL == 13 ? 0 : 1];
#else
        const uint32_t inv = P::INV;
#endif
#pragma unroll
        for (int j = 0; j < NL; ++j) t[j] = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            const uint32_t bi = b.v[i];
#pragma unroll
            for (int j = 0; j < NL; ++j) mad_wide(t[j], a.v[j], bi);
            const uint32_t m = ((uint32_t)t[0] * inv) & MASK;
#pragma unroll
            for (int j = 0; j < NL; ++j) mad_wide(t[j], m, P::MOD(j));
            // t[0] = 0 mod 2^29: drop it, carry its upper part into the next column, shift the window down
            const uint64_t c = t[0] >> LB;
#pragma unroll
            for (int j = 0; j < NL - 1; ++j) t[j] = t[j + 1];
            t[NL - 1] = 0;
            t[0] += c;
        }
        uint32_t r[NL];
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            if (j < NL - 1) {
                r[j] = (uint32_t)t[j] & MASK;
                t[j + 1] += t[j] >> LB;
            } else {
                r[j] = (uint32_t)t[j];  // result < 2p
            }
        }
        return cond_sub(r);
    }
    ZK_HD Fq29 sqr() const { return *this * *this; }
    static ZK_HD Fq29 mul_after(const Fq29& a, const Fq29& b, uint32_t& tok) {
        (void)tok;
        return a * b;
    }

    // ---- wire formats ------------------------------------------------------------------------------------------------
    // plain little-endian integer in STD_WORDS u32 words <-> NL limbs of 29 bits (no arithmetic, value unchanged)
    static ZK_HD Fq29 unpack(const uint32_t* w) {
        Fq29 r;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            const int bit = LB * i, lo = bit >> 5, sh = bit & 31;
            uint32_t x = w[lo] >> sh;
            if (sh + LB > 32 && lo + 1 < P::STD_WORDS) x |= w[lo + 1] << (32 - sh);
            r.v[i] = x & MASK;
        }
        return r;
    }
    ZK_HD void pack(uint32_t* w) const {
#pragma unroll
        for (int k = 0; k < P::STD_WORDS; ++k) w[k] = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            const int bit = LB * i, lo = bit >> 5, sh = bit & 31;
            if (lo < P::STD_WORDS) w[lo] |= v[i] << sh;
            if (sh + LB > 32 && lo + 1 < P::STD_WORDS) w[lo + 1] |= v[i] >> (32 - sh);
        }
    }
    // arkworks Montgomery limbs (R = 2^(32 STD_WORDS)) <-> internal (R' = 2^(29 NL))
    static ZK_HD Fq29 from_std(const uint32_t* w) {
        Fq29 k;
#pragma unroll
        for (int i = 0; i < NL; ++i) k.v[i] = P::TO_INT(i);
        return unpack(w) * k;
    }
    ZK_HD void to_std(uint32_t* w) const {
        Fq29 k;
#pragma unroll
        for (int i = 0; i < NL; ++i) k.v[i] = P::TO_STD(i);
        (*this * k).pack(w);
    }
};

}  // namespace zk
