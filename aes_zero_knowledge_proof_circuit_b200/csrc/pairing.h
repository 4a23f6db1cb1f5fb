// BLS12-377 ate pairing on the HOST, for the verifier only (csrc/verifier.cpp).
//
// Stands in for what the reference's `verify_encryption` (src/lib.rs:116-136) reaches through
// simpleworks::marlin::verify_proof -> ark-poly-commit 0.3.0 marlin_pc::check_combinations -> KZG10::batch_check ->
// Bls12_377::product_of_pairings (ark-ec 0.3.0 / ark-bls12-377 0.3.0, Cargo.lock:77,118).  Verification is two pairing
// equations per proof: CPU work in the reference and CPU work here ("GPU acceleration not
// warranted"), so this is plain portable C++ over ff.cuh's host multiplier.
//
// Tower (ark-bls12-377's): Fq2 = Fq[u]/(u^2 + 5); Fq12 = Fq2[w]/(w^6 - u) (ark's Fq6 = Fq2[v]/(v^3 - u), Fq12 = Fq6[w]/(w^2 - v)
// flattened); G2 is the D-type twist y^2 = x^3 + 1/u over Fq2, untwisted by (x', y') -> (x' w^2, y' w^3).
// The Miller loop runs over the BLS parameter x with affine G2 arithmetic; the final exponentiation is the plain power
// (q^12 - 1)/r.  Only pairing EQUATIONS are checked, which hold for any bilinear non-degenerate pairing; GT values are
// cross-checked bit for bit against the big-integer model in tools/pairing_model.py (tests/test_verifier.py).
#pragma once
#include <cstdint>

#include "ec.cuh"
#include "pairing_params_gen.h"

namespace zk {
namespace pairing {

using Fq = Fp<Fq377Params>;
using G1A = Affine<G1_377Params>;

inline Fq fq_from_limbs(const uint32_t* w) {
    Fq r;
    for (int i = 0; i < 12; ++i) r.v[i] = w[i];
    return r;
}

struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    static Fq2 from_fq(const Fq& a) { return {a, Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    static Fq times5(const Fq& a) {
        Fq a2 = a.dbl();
        return a2.dbl() + a;
    }
    // (a0 + a1 u)(b0 + b1 u) = a0 b0 - 5 a1 b1 + (a0 b1 + a1 b0) u, three base-field products
    Fq2 operator*(const Fq2& o) const {
        Fq v0 = c0 * o.c0, v1 = c1 * o.c1;
        Fq mid = (c0 + c1) * (o.c0 + o.c1) - v0 - v1;
        return {v0 - times5(v1), mid};
    }
    Fq2 scale(const Fq& s) const { return {c0 * s, c1 * s}; }
    Fq2 mul_by_u() const { return {times5(c1).neg(), c0}; }  // (a0 + a1 u) u = -5 a1 + a0 u
    Fq2 inverse() const {
        Fq n = (c0 * c0 + times5(c1 * c1)).inverse();  // norm = a0^2 + 5 a1^2
        return {c0 * n, (c1 * n).neg()};
    }
};

struct G2A {  // affine point on the twist; inf flag explicit
    Fq2 x, y;
    bool inf = false;
    static G2A infinity() {
        G2A r;
        r.x = Fq2::zero();
        r.y = Fq2::zero();
        r.inf = true;
        return r;
    }
    static G2A generator() {
        using namespace pairing_params;
        G2A r;
        r.x = {fq_from_limbs(G2_X_C0), fq_from_limbs(G2_X_C1)};
        r.y = {fq_from_limbs(G2_Y_C0), fq_from_limbs(G2_Y_C1)};
        return r;
    }
    bool on_curve() const {
        if (inf) return true;
        Fq2 b = {Fq::zero(), fq_from_limbs(pairing_params::TWIST_B_C1)};
        return y * y == x * x * x + b;
    }
    G2A neg() const {
        G2A r = *this;
        r.y = y.neg();
        return r;
    }
};

// slope of the tangent at t / of the chord t -> q (callers exclude the vertical cases)
inline Fq2 g2_tangent(const G2A& t) {
    Fq2 x2 = t.x * t.x;
    return (x2 + x2 + x2) * (t.y + t.y).inverse();
}
inline Fq2 g2_chord(const G2A& t, const G2A& q) { return (q.y - t.y) * (q.x - t.x).inverse(); }
inline G2A g2_step(const G2A& t, const G2A& q, const Fq2& lam) {  // third point of the line through t, q with slope lam
    G2A r;
    r.x = lam * lam - t.x - q.x;
    r.y = lam * (t.x - r.x) - t.y;
    return r;
}
inline G2A g2_add(const G2A& a, const G2A& b) {
    if (a.inf) return b;
    if (b.inf) return a;
    if (a.x == b.x) {
        if ((a.y + b.y).is_zero()) return G2A::infinity();
        return g2_step(a, a, g2_tangent(a));
    }
    return g2_step(a, b, g2_chord(a, b));
}
// s: canonical (non-Montgomery) little-endian limbs
inline G2A g2_mul(const G2A& p, const uint32_t* s, int nlimbs) {
    G2A acc = G2A::infinity();
    for (int i = nlimbs * 32 - 1; i >= 0; --i) {
        acc = g2_add(acc, acc);
        if ((s[i >> 5] >> (i & 31)) & 1) acc = g2_add(acc, p);
    }
    return acc;
}

struct Fq12 {
    Fq2 c[6];  // sum c[i] w^i, w^6 = u
    static Fq12 one() {
        Fq12 r;
        for (int i = 0; i < 6; ++i) r.c[i] = Fq2::zero();
        r.c[0] = Fq2::one();
        return r;
    }
    bool operator==(const Fq12& o) const {
        for (int i = 0; i < 6; ++i)
            if (!(c[i] == o.c[i])) return false;
        return true;
    }
    Fq12 operator*(const Fq12& o) const {
        Fq2 t[11];
        for (int i = 0; i < 11; ++i) t[i] = Fq2::zero();
        for (int i = 0; i < 6; ++i) {
            if (c[i].is_zero()) continue;
            for (int j = 0; j < 6; ++j) {
                if (o.c[j].is_zero()) continue;
                t[i + j] = t[i + j] + c[i] * o.c[j];
            }
        }
        Fq12 r;
        for (int k = 0; k < 6; ++k) r.c[k] = t[k];
        for (int k = 6; k < 11; ++k) r.c[k - 6] = r.c[k - 6] + t[k].mul_by_u();
        return r;
    }
    Fq12 pow(const uint32_t* e, int nlimbs) const {
        Fq12 r = one();
        bool started = false;
        for (int i = nlimbs * 32 - 1; i >= 0; --i) {
            if (started) r = r * r;
            if ((e[i >> 5] >> (i & 31)) & 1) {
                r = started ? r * *this : *this;
                started = true;
            }
        }
        return r;
    }
};

// line through T (slope lam on the twist) evaluated at P in G1, untwisted: yP - lam xP w + (lam x_T - y_T) w^3
inline Fq12 line_at(const G2A& t, const Fq2& lam, const G1A& p) {
    Fq12 l;
    for (int i = 0; i < 6; ++i) l.c[i] = Fq2::zero();
    l.c[0] = Fq2::from_fq(p.y);
    l.c[1] = lam.scale(p.x).neg();
    l.c[3] = lam * t.x - t.y;
    return l;
}

// f_{x,Q}(P); P or Q at infinity contribute 1
inline Fq12 miller_loop(const G1A& p, const G2A& q) {
    Fq12 f = Fq12::one();
    if (p.is_inf() || q.inf) return f;
    G2A t = q;
    const uint64_t x = pairing_params::BLS_X;
    int top = 63;
    while (!((x >> top) & 1)) --top;
    for (int i = top - 1; i >= 0; --i) {
        Fq2 lam = g2_tangent(t);
        f = (f * f) * line_at(t, lam, p);
        t = g2_step(t, t, lam);
        if ((x >> i) & 1) {
            lam = g2_chord(t, q);
            f = f * line_at(t, lam, p);
            t = g2_step(t, q, lam);
        }
    }
    return f;
}
inline Fq12 final_exponentiation(const Fq12& f) { return f.pow(pairing_params::FINAL_EXP, pairing_params::FINAL_EXP_LIMBS); }
inline Fq12 pairing(const G1A& p, const G2A& q) { return final_exponentiation(miller_loop(p, q)); }
// e(a1, b1) * e(a2, b2) == 1 with one final exponentiation
inline bool pairing_product_is_one(const G1A& a1, const G2A& b1, const G1A& a2, const G2A& b2) {
    return final_exponentiation(miller_loop(a1, b1) * miller_loop(a2, b2)) == Fq12::one();
}

}  // namespace pairing
}  // namespace zk
