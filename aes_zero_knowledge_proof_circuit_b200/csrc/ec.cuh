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

// Point doubling / full addition are off the mixed-addition loop (bucket reduction, rare mixed-add corner cases): they are
// out of line and issue their products through the out-of-line multiplier (Fp::mul_call), so a kernel that uses them carries
// one copy of the multiplier instead of fourteen per formula (60 KB of SASS each in round 1).
#if defined(__CUDACC__)
#define ZK_HD_COLD __host__ __device__ __noinline__
#else
#define ZK_HD_COLD __attribute__((noinline))
#endif

namespace zk {

template <class C>
struct FqOf {
    using type = Fp<typename C::FqP>;
};

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
    // (argument by value: a reference would force the caller's point into local memory on every loop iteration -- the
    // 96-byte stack frame and the 843 M dead local stores of the round-1 ncu capture)
    static ZK_HD_COLD XYZZ dbl_affine(Affine<C> p) {
        if (p.is_inf() || p.y.is_zero()) return inf();
        XYZZ r;
        Fq u = p.y.dbl();
        Fq v = Fq::sqr_call(u);
        Fq w = Fq::mul_call(u, v);
        Fq s = Fq::mul_call(p.x, v);
        Fq x2 = Fq::sqr_call(p.x);
        Fq m = x2.dbl() + x2;
        r.x = Fq::sqr_call(m) - s.dbl();
        r.y = Fq::mul_call(m, s - r.x) - Fq::mul_call(w, p.y);
        r.zz = v;
        r.zzz = w;
        return r;
    }

    // dbl-2008-s-1, a = 0
    ZK_HD_COLD XYZZ dbl() const {
        if (is_inf() || y.is_zero()) return inf();
        XYZZ r;
        Fq u = y.dbl();
        Fq v = Fq::sqr_call(u);
        Fq w = Fq::mul_call(u, v);
        Fq s = Fq::mul_call(x, v);
        Fq x2 = Fq::sqr_call(x);
        Fq m = x2.dbl() + x2;
        r.x = Fq::sqr_call(m) - s.dbl();
        r.y = Fq::mul_call(m, s - r.x) - Fq::mul_call(w, y);
        r.zz = Fq::mul_call(v, zz);
        r.zzz = Fq::mul_call(w, zzz);
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
        Fq pp = Fq::mul_after(pp_, pp_, tok);  // (the inlined form keeps products only: its ordering tokens chain through mul_after)
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

    // madd with the ten products issued through the out-of-line multiplier (Fp::mul_call): same values, ~1/8 of the code
    ZK_HD void madd_call(const Affine<C>& p) {
        if (p.is_inf()) return;
        if (is_inf()) {
            *this = from_affine(p);
            return;
        }
        Fq u2 = Fq::mul_call(p.x, zz);
        Fq s2 = Fq::mul_call(p.y, zzz);
        Fq pp_ = u2 - x;
        Fq r_ = s2 - y;
        if (pp_.is_zero()) {
            if (r_.is_zero())
                *this = dbl_affine(p);
            else
                *this = inf();
            return;
        }
        Fq pp = Fq::sqr_call(pp_);
        Fq ppp = Fq::mul_call(pp_, pp);
        Fq q = Fq::mul_call(x, pp);
        Fq r2 = Fq::sqr_call(r_);
        Fq x3 = r2 - ppp - q.dbl();
        Fq yp = Fq::mul_call(y, ppp);
        zz = Fq::mul_call(zz, pp);
        zzz = Fq::mul_call(zzz, ppp);
        y = Fq::mul_call(r_, q - x3) - yp;
        x = x3;
    }

    // this += o, add-2008-s
    ZK_HD_COLD void add(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) {
            *this = o;
            return;
        }
        Fq u1 = Fq::mul_call(x, o.zz);
        Fq u2 = Fq::mul_call(o.x, zz);
        Fq s1 = Fq::mul_call(y, o.zzz);
        Fq s2 = Fq::mul_call(o.y, zzz);
        Fq pp_ = u2 - u1;
        Fq r_ = s2 - s1;
        if (pp_.is_zero()) {
            if (r_.is_zero())
                *this = dbl();
            else
                *this = inf();
            return;
        }
        Fq pp = Fq::sqr_call(pp_);
        Fq ppp = Fq::mul_call(pp_, pp);
        Fq q = Fq::mul_call(u1, pp);
        Fq x3 = Fq::sqr_call(r_) - ppp - q.dbl();
        y = Fq::mul_call(r_, q - x3) - Fq::mul_call(s1, ppp);
        x = x3;
        zz = Fq::mul_call(Fq::mul_call(zz, o.zz), pp);
        zzz = Fq::mul_call(Fq::mul_call(zzz, o.zzz), ppp);
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
