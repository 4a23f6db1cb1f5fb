// CPU check of the structured final exponentiation in csrc/pairing.h (test infrastructure, built on demand by tests/test_verifier.py).
// For a Miller value f = f_{x,Q}(P) of points P = a G1, Q = b G2:
//   1  final_exponentiation_cubed(f) == final_exponentiation(f)^3, bit for bit
//   2  the Frobenius maps are the q-th and q^2-th powers (compared with plain powers by the modulus), conj6 is frobenius^6
//   4  fq12_inverse(f) * f == 1
// Returns 0 when all hold, otherwise the number of the first failing check.
#include <cstdint>
#include <cstring>

#include "../aes_zero_knowledge_proof_circuit_b200/csrc/pairing.h"

using namespace zk;
using namespace zk::pairing;

static void scalar_limbs(uint64_t s, uint32_t out[8]) {
    memset(out, 0, 32);
    out[0] = (uint32_t)s;
    out[1] = (uint32_t)(s >> 32);
}

extern "C" int pairing_fast_check(uint64_t a, uint64_t b) {
    uint32_t sa[8], sb[8];
    scalar_limbs(a, sa);
    scalar_limbs(b, sb);
    XYZZ<G1_377Params> acc = XYZZ<G1_377Params>::inf();
    const G1A g = G1A::generator();
    for (int i = 63; i >= 0; --i) {
        acc = acc.dbl();
        if ((a >> i) & 1) acc.madd(g);
    }
    const G1A p = acc.to_affine();
    const G2A q = g2_mul(G2A::generator(), sb, 8);
    const Fq12 f = miller_loop(p, q);
    const Fq12 plain = final_exponentiation(f);
    if (!(final_exponentiation_cubed(f) == plain * plain * plain)) return 1;
    uint32_t mod[12];
    for (int i = 0; i < 12; ++i) mod[i] = Fq377Params::MOD(i);
    const Fq12 fq = f.pow(mod, 12);
    if (!(fq12_frobenius(f) == fq)) return 2;
    if (!(fq12_frobenius2(f) == fq.pow(mod, 12))) return 3;
    if (!(fq12_inverse(f) * f == Fq12::one())) return 4;
    Fq12 c = f;
    for (int i = 0; i < 6; ++i) c = fq12_frobenius(c);
    if (!(c == fq12_conj6(f))) return 5;
    // bilinearity through the equation form the verifier uses: e(aP, bQ) * e(-abP, Q) == 1, and != 1 when the scalar is off by one
    XYZZ<G1_377Params> ab = XYZZ<G1_377Params>::inf();
    const unsigned __int128 prod = (unsigned __int128)a * b;
    for (int i = 127; i >= 0; --i) {
        ab = ab.dbl();
        if ((prod >> i) & 1) ab.madd(g);
    }
    const G1A abg = ab.to_affine();
    if (!pairing_product_is_one(p, q, abg.neg(), G2A::generator())) return 6;
    XYZZ<G1_377Params> ab1 = ab;
    ab1.madd(g);
    if (pairing_product_is_one(p, q, ab1.to_affine().neg(), G2A::generator())) return 7;
    return 0;
}
