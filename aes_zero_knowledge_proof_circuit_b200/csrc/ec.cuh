// G1 arithmetic for y^2 = x^3 + b (a = 0) over Fq: affine inputs, XYZZ accumulators.
//
// Replaces, on the hot path, ark-ec 0.3.0's short-Weierstrass-Jacobian GroupAffine/GroupProjective
// (reference Cargo.lock:118-120).  Wire layout of an affine point is arkworks': x || y, each 12 x u32
// LE Montgomery limbs (96 B).  The point at infinity is encoded as x = y = 0 on this side of the ABI
// (it is not on the curve because b != 0); ark's separate `infinity: bool` is mapped at the boundary.
//
// XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) is used for buckets because the mixed addition is 8M+2S with
// no field inversion and no Z^2/Z^3 recomputation.  Results leave the device as XYZZ window sums and are
// normalised to affine on the host (host code uses the portable multiplier in ff.cuh).
#pragma once
#include <type_traits>

#include "ff.cuh"
#include "fq29.cuh"

// Point doubling / full addition are off the inner loop (bucket reduction, rare mixed-add corner cases):
// keeping them out of line cuts both compile time and instruction-cache footprint of the hot kernels.
#if defined(__CUDACC__)
#define ZK_HD_COLD __host__ __device__ __noinline__
#else
#define ZK_HD_COLD __attribute__((noinline))
#endif

namespace zk {

// base-field type of a curve description: Fp<FqP> (arkworks layout) unless the description names its own (C::FqCustom:
// the radix-2^29 form the MSM kernels compute in)
template <class C, class = void>
struct FqOf {
    using type = Fp<typename C::FqP>;
};
template <class C>
struct FqOf<C, std::void_t<typename C::FqCustom>> {
    using type = typename C::FqCustom;
};
// internal curve descriptions used by msm.cu
struct G1_377R29 {
    using FqP = Fq377Params;
    using FrP = Fr377Params;
    using FqCustom = Fq29<Fq377R29Params>;
    static constexpr int CURVE_ID = 377;
};
struct G1_381R29 {
    using FqP = Fq381Params;
    using FrP = Fr381Params;
    using FqCustom = Fq29<Fq381R29Params>;
    static constexpr int CURVE_ID = 381;
};
// Which form the MSM kernels compute in.  Default: the arkworks 32-bit-limb form.  -DZK_MSM_R29 selects the radix-2^29
// form; it is bit-exact (tests run both) but SLOWER on B200: every 32x32->64 IMAD (WIDE or HI, with or without carry)
// issues at half rate, so 13 x 13 limb products cannot beat 12 x 12 (profiles/ubench_r1.txt).
template <class C> struct InternalCurve { using type = C; };
#if defined(ZK_MSM_R29)
template <> struct InternalCurve<G1_377Params> { using type = G1_377R29; };
template <> struct InternalCurve<G1_381Params> { using type = G1_381R29; };
#endif

template <class C>
struct Affine {
    using Fq = typename FqOf<C>::type;
    Fq x, y;
    ZK_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static ZK_HD Affine inf() {
        Affine r;
        r.x = Fq::zero();
        r.y = Fq::zero();
        return r;
    }
    static ZK_HD Affine generator() {
        Affine r;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            r.x.v[i] = C::GX(i);
            r.y.v[i] = C::GY(i);
        }
        return r;
    }
    ZK_HD Affine neg() const {
        Affine r;
        r.x = x;
        r.y = y.neg();
        return r;
    }
    ZK_HD bool on_curve() const {
        if (is_inf()) return true;
        Fq b;
#pragma unroll
        for (int i = 0; i < 12; ++i) b.v[i] = C::B(i);
        return y.sqr() == x.sqr() * x + b;
    }
};

template <class C>
struct XYZZ {
    using Fq = typename FqOf<C>::type;
    Fq x, y, zz, zzz;

    ZK_HD bool is_inf() const { return zz.is_zero(); }
    static ZK_HD XYZZ inf() {
        XYZZ r;
        r.x = Fq::zero();
        r.y = Fq::zero();
        r.zz = Fq::zero();
        r.zzz = Fq::zero();
        return r;
    }
    static ZK_HD XYZZ from_affine(const Affine<C>& p) {
        XYZZ r;
        if (p.is_inf()) return inf();
        r.x = p.x;
        r.y = p.y;
        r.zz = Fq::one();
        r.zzz = Fq::one();
        return r;
    }

    // 2*P for an affine P (mdbl-2008-s-1, a = 0)
    static ZK_HD_COLD XYZZ dbl_affine(const Affine<C>& p) {
        if (p.is_inf() || p.y.is_zero()) return inf();
        XYZZ r;
        Fq u = p.y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = p.x * v;
        Fq x2 = p.x.sqr();
        Fq m = x2.dbl() + x2;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * p.y;
        r.zz = v;
        r.zzz = w;
        return r;
    }

    // dbl-2008-s-1, a = 0
    ZK_HD_COLD XYZZ dbl() const {
        if (is_inf() || y.is_zero()) return inf();
        XYZZ r;
        Fq u = y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = x * v;
        Fq x2 = x.sqr();
        Fq m = x2.dbl() + x2;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }

    // this += P (affine), madd-2008-s.  Handles P = inf, this = inf, P = +-this.
    // The ten field products are issued as one chain (Fp::mul_after) so that ptxas keeps a single set of carry
    // predicates live; see ff.cuh.
    ZK_HD void madd(const Affine<C>& p) {
        if (p.is_inf()) return;
        if (is_inf()) {
            *this = from_affine(p);
            return;
        }
        uint32_t tok = 0;
        Fq u2 = Fq::mul_after(p.x, zz, tok);
        Fq s2 = Fq::mul_after(p.y, zzz, tok);
        Fq pp_ = u2 - x;
        Fq r_ = s2 - y;
        if (pp_.is_zero()) {
            if (r_.is_zero())
                *this = dbl_affine(p);
            else
                *this = inf();
            return;
        }
        Fq pp = Fq::mul_after(pp_, pp_, tok);
        Fq ppp = Fq::mul_after(pp_, pp, tok);
        Fq q = Fq::mul_after(x, pp, tok);
        Fq r2 = Fq::mul_after(r_, r_, tok);
        Fq x3 = r2 - ppp - q.dbl();
        Fq yp = Fq::mul_after(y, ppp, tok);
        zz = Fq::mul_after(zz, pp, tok);
        zzz = Fq::mul_after(zzz, ppp, tok);
        y = Fq::mul_after(r_, q - x3, tok) - yp;
        x = x3;
    }

    // this += o, add-2008-s
    ZK_HD_COLD void add(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) {
            *this = o;
            return;
        }
        Fq u1 = x * o.zz;
        Fq u2 = o.x * zz;
        Fq s1 = y * o.zzz;
        Fq s2 = o.y * zzz;
        Fq pp_ = u2 - u1;
        Fq r_ = s2 - s1;
        if (pp_.is_zero()) {
            if (r_.is_zero())
                *this = dbl();
            else
                *this = inf();
            return;
        }
        Fq pp = pp_.sqr();
        Fq ppp = pp_ * pp;
        Fq q = u1 * pp;
        Fq x3 = r_.sqr() - ppp - q.dbl();
        y = r_ * (q - x3) - s1 * ppp;
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }

    ZK_HD XYZZ neg() const {
        XYZZ r = *this;
        r.y = y.neg();
        return r;
    }

    // one field inversion; host-side use (window combine, serialisation)
    ZK_HD_COLD Affine<C> to_affine() const {
        if (is_inf()) return Affine<C>::inf();
        Affine<C> r;
        Fq izzz = zzz.inverse();          // 1/ZZZ
        Fq izz = (izzz * zz).sqr();       // (ZZ/ZZZ)^2 = 1/ZZ   because ZZ^3 = ZZZ^2
        r.x = x * izz;
        r.y = y * izzz;
        return r;
    }
};

using G1Affine377 = Affine<G1_377Params>;
using G1XYZZ377 = XYZZ<G1_377Params>;
using G1Affine381 = Affine<G1_381Params>;
using G1XYZZ381 = XYZZ<G1_381Params>;

}  // namespace zk
