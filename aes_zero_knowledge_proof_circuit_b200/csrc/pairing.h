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
// The Miller loop runs over the BLS parameter x with affine G2 arithmetic; the final exponentiation of pairing() is the plain power
// (q^12 - 1)/r, the equation check uses the structured form (its cube, see below).  Only pairing EQUATIONS are checked, which hold for any bilinear non-degenerate pairing; GT values are
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

// ---- structured final exponentiation (for the equation checks; the plain power above defines the GT VALUES the tests compare) --------------------
// f^(3 (q^12 - 1) / r) = easy part f^((q^6 - 1)(q^2 + 1)), then the BLS12 hard part through
//     3 (q^4 - q^2 + 1) / r = (x - 1)^2 (x + q) (x^2 + q^2 - 1) + 3           (checked for this curve by tools/pairing_model.py's integers),
// i.e. five powers by the 64-bit x, two Frobenius maps and a handful of products instead of a 4,300-bit power.  The result is the CUBE of the
// plain power; 3 does not divide r, so "== 1" is the same statement (tests/pairing_check.cpp compares the two bit for bit).
struct Fq6 {  // Fq2[v] / (v^3 - u): the even (or odd) coefficients of an Fq12 element, v = w^2
    Fq2 a, b, c;
    Fq6 operator+(const Fq6& o) const { return {a + o.a, b + o.b, c + o.c}; }
    Fq6 operator-(const Fq6& o) const { return {a - o.a, b - o.b, c - o.c}; }
    Fq6 operator*(const Fq6& o) const {
        const Fq2 t0 = a * o.a, t1 = a * o.b + b * o.a, t2 = a * o.c + b * o.b + c * o.a, t3 = b * o.c + c * o.b, t4 = c * o.c;
        return {t0 + t3.mul_by_u(), t1 + t4.mul_by_u(), t2};
    }
    Fq6 mul_by_v() const { return {c.mul_by_u(), a, b}; }
    Fq6 inverse() const {
        const Fq2 t0 = a * a - (b * c).mul_by_u(), t1 = (c * c).mul_by_u() - a * b, t2 = b * b - a * c;
        const Fq2 d = (a * t0 + (c * t1 + b * t2).mul_by_u()).inverse();
        return {t0 * d, t1 * d, t2 * d};
    }
};
inline Fq12 fq12_conj6(const Fq12& f) {  // f^(q^6): w -> -w
    Fq12 r = f;
    for (int i = 1; i < 6; i += 2) r.c[i] = f.c[i].neg();
    return r;
}
inline Fq12 fq12_inverse(const Fq12& f) {  // (A + w B)^-1 = (A - w B) / (A^2 - v B^2)
    const Fq6 A = {f.c[0], f.c[2], f.c[4]}, B = {f.c[1], f.c[3], f.c[5]};
    const Fq6 n = (A * A - (B * B).mul_by_v()).inverse();
    const Fq6 ra = A * n, rb = B * n;
    Fq12 r;
    r.c[0] = ra.a; r.c[2] = ra.b; r.c[4] = ra.c;
    r.c[1] = rb.a.neg(); r.c[3] = rb.b.neg(); r.c[5] = rb.c.neg();
    return r;
}
// w^(q-1) = u^((q-1)/6) =: g, so (c_i w^i)^q = conj(c_i) g^i w^i (u^q = -u because -5 is a non-residue); for q^2 the factor is norm(g)^i in Fq
struct FrobeniusTable {
    Fq2 g[6];
    Fq n[6];
    FrobeniusTable() {
        uint32_t e[12];  // (q - 1) / 6 by long division of the modulus limbs
        uint64_t rem = 0;
        for (int i = 11; i >= 0; --i) {
            const uint64_t cur = (rem << 32) | (uint64_t)(Fq377Params::MOD(i) - (i == 0 ? 1u : 0u));  // q is odd: the low limb does not borrow
            e[i] = (uint32_t)(cur / 6);
            rem = cur % 6;
        }
        Fq2 base = {Fq::zero(), Fq::one()}, acc = Fq2::one();
        for (int i = 12 * 32 - 1; i >= 0; --i) {
            acc = acc * acc;
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * base;
        }
        g[0] = Fq2::one();
        n[0] = Fq::one();
        const Fq nn = acc.c0 * acc.c0 + Fq2::times5(acc.c1 * acc.c1);
        for (int i = 1; i < 6; ++i) {
            g[i] = g[i - 1] * acc;
            n[i] = n[i - 1] * nn;
        }
    }
};
inline const FrobeniusTable& frobenius_table() {
    static const FrobeniusTable t;
    return t;
}
inline Fq12 fq12_frobenius(const Fq12& f) {
    const FrobeniusTable& t = frobenius_table();
    Fq12 r;
    for (int i = 0; i < 6; ++i) r.c[i] = Fq2{f.c[i].c0, f.c[i].c1.neg()} * t.g[i];
    return r;
}
inline Fq12 fq12_frobenius2(const Fq12& f) {
    const FrobeniusTable& t = frobenius_table();
    Fq12 r;
    for (int i = 0; i < 6; ++i) r.c[i] = f.c[i].scale(t.n[i]);
    return r;
}
inline Fq12 fq12_pow_x(const Fq12& f) {
    const uint32_t x[2] = {(uint32_t)pairing_params::BLS_X, (uint32_t)(pairing_params::BLS_X >> 32)};
    return f.pow(x, 2);
}
inline Fq12 final_exponentiation_cubed(const Fq12& f) {
    Fq12 m = fq12_conj6(f) * fq12_inverse(f);  // f^(q^6 - 1): from here on the inverse is the conjugate
    m = fq12_frobenius2(m) * m;                // ^(q^2 + 1)
    const Fq12 t0 = fq12_pow_x(m) * fq12_conj6(m);                                       // m^(x - 1)
    const Fq12 t1 = fq12_pow_x(t0) * fq12_conj6(t0);                                     // m^((x - 1)^2)
    const Fq12 t2 = fq12_pow_x(t1) * fq12_frobenius(t1);                                 // ^(x + q)
    const Fq12 t3 = fq12_pow_x(fq12_pow_x(t2)) * fq12_frobenius2(t2) * fq12_conj6(t2);   // ^(x^2 + q^2 - 1)
    return t3 * (m * m * m);
}
// e(a1, b1) * e(a2, b2) == 1 with one final exponentiation
inline bool pairing_product_is_one(const G1A& a1, const G2A& b1, const G1A& a2, const G2A& b2) {
    const Fq12 f = miller_loop(a1, b1) * miller_loop(a2, b2);
    // a Miller value is never zero for points of order r; if it were, the plain power decides (0 != 1)
    bool zero = true;
    for (int i = 0; i < 6; ++i) zero = zero && f.c[i].is_zero();
    if (zero) return false;
    return final_exponentiation_cubed(f) == Fq12::one();
}

}  // namespace pairing
}  // namespace zk
