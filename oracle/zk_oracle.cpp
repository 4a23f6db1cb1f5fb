// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.  The product path
// (libzkaes_b200.so) never links or calls it.
//
// CPU restatement (plain C++17, 64-bit limbs + unsigned __int128) of the arithmetic that the reference's
// `encrypt()` hot path performs inside its pinned, un-vendored dependencies.  The reference itself is
// Rust and cannot be compiled in this image (no cargo/rustc), so each function below names
// the dependency + version whose published algorithm it follows and the reference call site that
// reaches it:
//   * Fp256 / Fp384 Montgomery arithmetic ......... ark-ff 0.3.0           (Cargo.lock:159)
//   * G1 Jacobian add / mixed add / double ........ ark-ec 0.3.0           (Cargo.lock:118)
//   * VariableBaseMSM::multi_scalar_mul ........... ark-ec 0.3.0 msm       (Cargo.lock:118; reached from src/lib.rs:111)
//   * Radix2EvaluationDomain fft/ifft/coset ....... ark-poly 0.3.0         (Cargo.lock:234; reached from src/lib.rs:111)
//   * byte-level AES-128 steps .................... src/aes.rs:10-269 (in tree)
// PARITY PINNING: AES functions are pinned by the reference's FIPS-197 tables (tests/integration_tests.rs:
// 50-310, src/aes.rs:277-360).  MSM/NTT have NO golden vectors in the reference ("parity unpinned" there);
// they are canonical mathematical objects and are cross-checked here against independent definitions
// (double-and-add, O(n^2) DFT) -- see tests/test_oracle.py.
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <thread>
#include <atomic>
#include <functional>

// this image's gcc has no libgomp, so the rayon-style data parallelism of the arkworks kernels is
// restated with std::thread: dynamic chunked parallel-for over [0, n)
static int g_threads = 0;
static int n_threads() {
    if (g_threads > 0) return g_threads;
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}
template <class F>
static void parallel_for(long long n, long long grain, F fn) {
    int nt = n_threads();
    if (nt <= 1 || n <= grain) {
        for (long long i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<long long> next(0);
    auto worker = [&]() {
        for (;;) {
            long long lo = next.fetch_add(grain);
            if (lo >= n) break;
            long long hi = std::min(n, lo + grain);
            for (long long i = lo; i < hi; ++i) fn(i);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

typedef unsigned __int128 u128;
typedef uint64_t u64;

// ---------------------------------------------------------------------------------------------
// Prime field, N 64-bit limbs, Montgomery form (ark-ff 0.3.0 `Fp256`/`Fp384`: R = 2^(64N)).
// ---------------------------------------------------------------------------------------------
template <int N>
struct FieldCtx {
    u64 mod[N];
    u64 inv;    // -mod^-1 mod 2^64
    u64 one[N]; // R mod p
    u64 r2[N];  // R^2 mod p
};

template <int N>
static inline bool geq(const u64* a, const u64* b) {
    for (int i = N - 1; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}
template <int N>
static inline u64 add_limbs(u64* r, const u64* a, const u64* b) {
    u128 c = 0;
    for (int i = 0; i < N; ++i) {
        c += (u128)a[i] + b[i];
        r[i] = (u64)c;
        c >>= 64;
    }
    return (u64)c;
}
template <int N>
static inline u64 sub_limbs(u64* r, const u64* a, const u64* b) {
    u64 borrow = 0;
    for (int i = 0; i < N; ++i) {
        u128 t = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)t;
        borrow = (u64)(t >> 64) & 1;
    }
    return borrow;
}

template <int N>
struct Fe {
    u64 v[N];
    bool is_zero() const {
        u64 o = 0;
        for (int i = 0; i < N; ++i) o |= v[i];
        return o == 0;
    }
    bool operator==(const Fe& b) const { return memcmp(v, b.v, sizeof v) == 0; }
};

template <int N>
static inline Fe<N> f_add(const FieldCtx<N>& F, const Fe<N>& a, const Fe<N>& b) {
    Fe<N> r;
    add_limbs<N>(r.v, a.v, b.v);
    if (geq<N>(r.v, F.mod)) sub_limbs<N>(r.v, r.v, F.mod);
    return r;
}
template <int N>
static inline Fe<N> f_sub(const FieldCtx<N>& F, const Fe<N>& a, const Fe<N>& b) {
    Fe<N> r;
    if (sub_limbs<N>(r.v, a.v, b.v)) add_limbs<N>(r.v, r.v, F.mod);
    return r;
}
template <int N>
static inline Fe<N> f_neg(const FieldCtx<N>& F, const Fe<N>& a) {
    if (a.is_zero()) return a;
    Fe<N> r;
    sub_limbs<N>(r.v, F.mod, a.v);
    return r;
}
// Montgomery product, coarsely-integrated operand scanning (the schedule ark-ff's mul_assign unrolls)
template <int N>
static inline Fe<N> f_mul(const FieldCtx<N>& F, const Fe<N>& a, const Fe<N>& b) {
    u64 t[N + 2] = {0};
    for (int i = 0; i < N; ++i) {
        u128 c = 0;
        for (int j = 0; j < N; ++j) {
            c += (u128)a.v[j] * b.v[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[N];
        t[N] = (u64)c;
        t[N + 1] = (u64)(c >> 64);
        u64 m = t[0] * F.inv;
        c = ((u128)m * F.mod[0] + t[0]) >> 64;
        for (int j = 1; j < N; ++j) {
            c += (u128)m * F.mod[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[N];
        t[N - 1] = (u64)c;
        t[N] = t[N + 1] + (u64)(c >> 64);
    }
    Fe<N> r;
    memcpy(r.v, t, sizeof r.v);
    if (t[N] || geq<N>(r.v, F.mod)) sub_limbs<N>(r.v, r.v, F.mod);
    return r;
}
template <int N>
static inline Fe<N> f_sqr(const FieldCtx<N>& F, const Fe<N>& a) { return f_mul(F, a, a); }
template <int N>
static inline Fe<N> f_one(const FieldCtx<N>& F) {
    Fe<N> r;
    memcpy(r.v, F.one, sizeof r.v);
    return r;
}
template <int N>
static inline Fe<N> f_zero() {
    Fe<N> r;
    memset(r.v, 0, sizeof r.v);
    return r;
}
template <int N>
static Fe<N> f_pow(const FieldCtx<N>& F, const Fe<N>& a, const u64* e, int nl) {
    Fe<N> r = f_one(F);
    for (int i = nl - 1; i >= 0; --i)
        for (int b = 63; b >= 0; --b) {
            r = f_sqr(F, r);
            if ((e[i] >> b) & 1) r = f_mul(F, r, a);
        }
    return r;
}
template <int N>
static Fe<N> f_inv(const FieldCtx<N>& F, const Fe<N>& a) {
    u64 e[N], two[N] = {2};
    sub_limbs<N>(e, F.mod, two);
    return f_pow(F, a, e, N);
}
template <int N>
static Fe<N> f_from_u64(const FieldCtx<N>& F, u64 x) {
    Fe<N> r = f_zero<N>(), r2;
    r.v[0] = x;
    memcpy(r2.v, F.r2, sizeof r2.v);
    return f_mul(F, r, r2);
}
template <int N>
static Fe<N> f_to_mont(const FieldCtx<N>& F, const Fe<N>& a) {
    Fe<N> r2;
    memcpy(r2.v, F.r2, sizeof r2.v);
    return f_mul(F, a, r2);
}
template <int N>
static Fe<N> f_from_mont(const FieldCtx<N>& F, const Fe<N>& a) {
    Fe<N> o = f_zero<N>();
    o.v[0] = 1;
    return f_mul(F, a, o);
}

static int hexval(char c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; }
template <int N>
static void parse_hex(const char* s, u64* out) {
    memset(out, 0, 8 * N);
    int len = (int)strlen(s);
    for (int i = 0; i < len; ++i) {
        int nib = hexval(s[len - 1 - i]);
        out[i / 16] |= (u64)nib << (4 * (i % 16));
    }
}
template <int N>
static void make_ctx(FieldCtx<N>& F, const char* hexmod) {
    parse_hex<N>(hexmod, F.mod);
    // -p^-1 mod 2^64 by Newton iteration
    u64 x = 1;
    for (int i = 0; i < 6; ++i) x *= 2 - F.mod[0] * x;
    F.inv = (u64)0 - x;
    // R mod p and R^2 mod p by repeated doubling of 1 (64N and 128N doublings)
    u64 t[N] = {1};
    for (int i = 0; i < 128 * N; ++i) {
        u64 c = add_limbs<N>(t, t, t);
        if (c || geq<N>(t, F.mod)) sub_limbs<N>(t, t, F.mod);
        if (i == 64 * N - 1) memcpy(F.one, t, sizeof F.one);
    }
    memcpy(F.r2, t, sizeof F.r2);
}

// ---------------------------------------------------------------------------------------------
// Curves: BLS12-377 (the reference's curve: src/lib.rs:47) and BLS12-381.
// ---------------------------------------------------------------------------------------------
struct Curve {
    FieldCtx<4> fr;
    FieldCtx<6> fq;
    Fe<6> b;          // curve coefficient, Montgomery
    Fe<6> gx, gy;     // generator, Montgomery
    int two_adicity;
    Fe<4> root;       // 2^two_adicity-th root of unity, Montgomery
    Fe<4> gen;        // multiplicative generator (coset shift), Montgomery
    int fr_bits;
};
static Curve g_curves[2];
static bool g_init = false;

static Fe<4> fr_pow_u64(const Curve& C, Fe<4> a, const u64* e, int nl) { return f_pow<4>(C.fr, a, e, nl); }

static void init_curve(Curve& C, const char* r, const char* q, u64 fr_gen, int s, u64 b, const char* gx, const char* gy,
                       int fr_bits) {
    make_ctx<4>(C.fr, r);
    make_ctx<6>(C.fq, q);
    C.b = f_from_u64<6>(C.fq, b);
    Fe<6> t;
    parse_hex<6>(gx, t.v);
    C.gx = f_to_mont<6>(C.fq, t);
    parse_hex<6>(gy, t.v);
    C.gy = f_to_mont<6>(C.fq, t);
    C.two_adicity = s;
    C.gen = f_from_u64<4>(C.fr, fr_gen);
    // root = gen^((r-1)/2^s)
    u64 e[4], one[4] = {1};
    sub_limbs<4>(e, C.fr.mod, one);
    for (int i = 0; i < s; ++i) {  // e >>= 1
        for (int k = 0; k < 4; ++k) e[k] = (e[k] >> 1) | (k < 3 ? e[k + 1] << 63 : 0);
    }
    C.root = fr_pow_u64(C, C.gen, e, 4);
    C.fr_bits = fr_bits;
}
static void ensure_init() {
    if (g_init) return;
    init_curve(g_curves[0], "12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001",
               "1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001", 22, 47, 1,
               "008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef",
               "01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6", 253);
    init_curve(g_curves[1], "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001",
               "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab", 7, 32, 4,
               "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb",
               "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1", 255);
    g_init = true;
}
static const Curve* curve_by_id(int id) {
    ensure_init();
    if (id == 377) return &g_curves[0];
    if (id == 381) return &g_curves[1];
    return nullptr;
}

// ---------------------------------------------------------------------------------------------
// G1 in Jacobian coordinates -- ark-ec 0.3.0 models/short_weierstrass_jacobian.rs
// (double_in_place: "dbl-2009-l" for a = 0; add_assign_mixed: "madd-2007-bl"; add_assign: "add-2007-bl")
// ---------------------------------------------------------------------------------------------
typedef Fe<6> Fq;
struct Aff { Fq x, y; bool inf; };
struct Jac { Fq x, y, z; };

#define QF C.fq
static inline Fq qa(const Curve& C, const Fq& a, const Fq& b) { return f_add<6>(QF, a, b); }
static inline Fq qs(const Curve& C, const Fq& a, const Fq& b) { return f_sub<6>(QF, a, b); }
static inline Fq qm(const Curve& C, const Fq& a, const Fq& b) { return f_mul<6>(QF, a, b); }
static inline Fq qd(const Curve& C, const Fq& a) { return f_add<6>(QF, a, a); }

static Jac jac_zero(const Curve& C) {
    Jac r;
    r.x = f_zero<6>();
    r.y = f_one<6>(C.fq);
    r.z = f_zero<6>();
    return r;
}
static bool jac_is_zero(const Jac& p) { return p.z.is_zero(); }

static void jac_double(const Curve& C, Jac& p) {
    if (jac_is_zero(p)) return;
    Fq a = qm(C, p.x, p.x);
    Fq b = qm(C, p.y, p.y);
    Fq c = qm(C, b, b);
    Fq xb = qa(C, p.x, b);
    Fq d = qd(C, qs(C, qs(C, qm(C, xb, xb), a), c));
    Fq e = qa(C, qd(C, a), a);
    Fq f = qm(C, e, e);
    Fq z3 = qd(C, qm(C, p.z, p.y));
    Fq x3 = qs(C, qs(C, f, d), d);
    Fq c8 = qd(C, qd(C, qd(C, c)));
    Fq y3 = qs(C, qm(C, qs(C, d, x3), e), c8);
    p.x = x3;
    p.y = y3;
    p.z = z3;
}
static void jac_add_mixed(const Curve& C, Jac& p, const Aff& q) {
    if (q.inf) return;
    if (jac_is_zero(p)) {
        p.x = q.x;
        p.y = q.y;
        p.z = f_one<6>(C.fq);
        return;
    }
    Fq z1z1 = qm(C, p.z, p.z);
    Fq u2 = qm(C, q.x, z1z1);
    Fq s2 = qm(C, qm(C, q.y, p.z), z1z1);
    if (p.x == u2 && p.y == s2) {
        jac_double(C, p);
        return;
    }
    Fq h = qs(C, u2, p.x);
    Fq hh = qm(C, h, h);
    Fq i = qd(C, qd(C, hh));
    Fq j = qm(C, h, i);
    Fq r = qd(C, qs(C, s2, p.y));
    Fq v = qm(C, p.x, i);
    Fq x3 = qs(C, qs(C, qs(C, qm(C, r, r), j), v), v);
    Fq y3 = qs(C, qm(C, r, qs(C, v, x3)), qd(C, qm(C, p.y, j)));
    Fq zh = qa(C, p.z, h);
    Fq z3 = qs(C, qs(C, qm(C, zh, zh), z1z1), hh);
    p.x = x3;
    p.y = y3;
    p.z = z3;
}
static void jac_add(const Curve& C, Jac& p, const Jac& q) {
    if (jac_is_zero(q)) return;
    if (jac_is_zero(p)) {
        p = q;
        return;
    }
    Fq z1z1 = qm(C, p.z, p.z);
    Fq z2z2 = qm(C, q.z, q.z);
    Fq u1 = qm(C, p.x, z2z2);
    Fq u2 = qm(C, q.x, z1z1);
    Fq s1 = qm(C, qm(C, p.y, q.z), z2z2);
    Fq s2 = qm(C, qm(C, q.y, p.z), z1z1);
    if (u1 == u2 && s1 == s2) {
        jac_double(C, p);
        return;
    }
    Fq h = qs(C, u2, u1);
    Fq i = qd(C, h);
    i = qm(C, i, i);
    Fq j = qm(C, h, i);
    Fq r = qd(C, qs(C, s2, s1));
    Fq v = qm(C, u1, i);
    Fq x3 = qs(C, qs(C, qs(C, qm(C, r, r), j), v), v);
    Fq y3 = qs(C, qm(C, r, qs(C, v, x3)), qd(C, qm(C, s1, j)));
    Fq zz = qa(C, p.z, q.z);
    Fq z3 = qm(C, qs(C, qs(C, qm(C, zz, zz), z1z1), z2z2), h);
    p.x = x3;
    p.y = y3;
    p.z = z3;
}
static Aff jac_to_affine(const Curve& C, const Jac& p) {
    Aff r;
    if (jac_is_zero(p)) {
        r.x = f_zero<6>();
        r.y = f_zero<6>();
        r.inf = true;
        return r;
    }
    Fq zi = f_inv<6>(C.fq, p.z);
    Fq zi2 = qm(C, zi, zi);
    r.x = qm(C, p.x, zi2);
    r.y = qm(C, p.y, qm(C, zi2, zi));
    r.inf = false;
    return r;
}
// wire format shared with the product: 96 B = x||y (Montgomery LE limbs); x=y=0 means infinity
static Aff load_aff(const u64* p) {
    Aff a;
    memcpy(a.x.v, p, 48);
    memcpy(a.y.v, p + 6, 48);
    a.inf = a.x.is_zero() && a.y.is_zero();
    return a;
}
static void store_aff(u64* p, const Aff& a) {
    if (a.inf) {
        memset(p, 0, 96);
        return;
    }
    memcpy(p, a.x.v, 48);
    memcpy(p + 6, a.y.v, 48);
}

static int bigint_bit(const u64* s, int i) { return (int)((s[i >> 6] >> (i & 63)) & 1); }

// scalar (canonical 4x64 LE) times point, MSB-first double-and-add: the independent MSM definition
static Jac scalar_mul(const Curve& C, const Aff& p, const u64* s) {
    Jac r = jac_zero(C);
    for (int i = 255; i >= 0; --i) {
        jac_double(C, r);
        if (bigint_bit(s, i)) jac_add_mixed(C, r, p);
    }
    return r;
}

static int ark_log2_ceil(size_t x) {
    if (x <= 1) return 0;
    int l = 0;
    size_t v = x - 1;
    while (v) {
        ++l;
        v >>= 1;
    }
    return l;
}

// ark-ec 0.3.0 msm/variable_base.rs VariableBaseMSM::multi_scalar_mul, restated:
// window c = 3 if n < 32 else (ceil_log2(n)*69/100)+2; windows start at 0,c,2c,.. < MODULUS_BITS, processed in
// parallel (rayon there, std::thread here); unit scalars short-cut into the first window; buckets 2^c-1; running-sum
// reduction; fold windows high->low with c doublings each.
static Jac msm_pippenger(const Curve& C, const u64* bases, const u64* scalars, size_t n) {
    int c = n < 32 ? 3 : (ark_log2_ceil(n) * 69 / 100) + 2;
    int num_bits = C.fr_bits;
    std::vector<int> starts;
    for (int w = 0; w < num_bits; w += c) starts.push_back(w);
    std::vector<Jac> sums(starts.size());
    const u64 one[4] = {1, 0, 0, 0};
    parallel_for((long long)starts.size(), 1, [&](long long wi) {
        int w_start = starts[wi];
        Jac res = jac_zero(C);
        std::vector<Jac> buckets(((size_t)1 << c) - 1, jac_zero(C));
        for (size_t i = 0; i < n; ++i) {
            const u64* s = scalars + 4 * i;
            if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;
            Aff base = load_aff(bases + 12 * i);
            if (memcmp(s, one, 32) == 0) {
                if (w_start == 0) jac_add_mixed(C, res, base);
                continue;
            }
            // (scalar >> w_start) mod 2^c
            int limb = w_start >> 6, off = w_start & 63;
            u64 v = s[limb] >> off;
            if (off && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
            v &= (((u64)1) << c) - 1;
            if (v) jac_add_mixed(C, buckets[v - 1], base);
        }
        Jac running = jac_zero(C);
        for (size_t b = buckets.size(); b-- > 0;) {
            jac_add(C, running, buckets[b]);
            jac_add(C, res, running);
        }
        sums[wi] = res;
    });
    Jac total = jac_zero(C);
    for (size_t wi = sums.size(); wi-- > 1;) {
        jac_add(C, total, sums[wi]);
        for (int k = 0; k < c; ++k) jac_double(C, total);
    }
    jac_add(C, total, sums[0]);
    return total;
}

// ---------------------------------------------------------------------------------------------
// Radix-2 NTT over Fr -- ark-poly 0.3.0 domain/radix2 (in-order in, in-order out).
// group_gen = ROOT^(2^(two_adicity - log_n)); ifft uses group_gen^-1 and scales by n^-1;
// coset_fft multiplies coefficient i by GEN^i first; coset_ifft multiplies by GEN^-i after.
// ---------------------------------------------------------------------------------------------
typedef Fe<4> Fr;
#define RF C.fr
static Fr domain_gen(const Curve& C, int log_n) {
    Fr g = C.root;
    for (int i = log_n; i < C.two_adicity; ++i) g = f_sqr<4>(RF, g);
    return g;
}
static void distribute_powers(const Curve& C, Fr* a, size_t n, const Fr& g) {
    // a[i] *= g^i ; chunked so threads are independent
    const size_t CH = 4096;
    parallel_for((long long)((n + CH - 1) / CH), 1, [&](long long c) {
        size_t lo = (size_t)c * CH, hi = std::min(n, lo + CH);
        u64 e[1] = {lo};
        Fr p = f_pow<4>(RF, g, e, 1);
        for (size_t i = lo; i < hi; ++i) {
            a[i] = f_mul<4>(RF, a[i], p);
            p = f_mul<4>(RF, p, g);
        }
    });
}
static void ntt_in_place(const Curve& C, Fr* a, int log_n, const Fr& omega) {
    size_t n = (size_t)1 << log_n;
    // bit-reversal permutation
    for (size_t i = 0; i < n; ++i) {
        size_t j = 0;
        for (int b = 0; b < log_n; ++b) j |= ((i >> b) & 1) << (log_n - 1 - b);
        if (i < j) std::swap(a[i], a[j]);
    }
    // precompute twiddles w^k for k < n/2
    std::vector<Fr> tw(n / 2 ? n / 2 : 1);
    tw[0] = f_one<4>(RF);
    for (size_t k = 1; k < n / 2; ++k) tw[k] = f_mul<4>(RF, tw[k - 1], omega);
    for (int s = 1; s <= log_n; ++s) {
        size_t m = (size_t)1 << s, half = m >> 1, stride = n / m;
        parallel_for((long long)(n / 2), 2048, [&](long long idx) {
            size_t k = (size_t)idx / half, j = (size_t)idx % half;
            size_t i0 = k * m + j, i1 = i0 + half;
            Fr t = f_mul<4>(RF, a[i1], tw[j * stride]);
            Fr u = a[i0];
            a[i0] = f_add<4>(RF, u, t);
            a[i1] = f_sub<4>(RF, u, t);
        });
    }
}

// ---------------------------------------------------------------------------------------------
// Byte-level AES-128 (the reference's plain mirror, src/aes.rs).
// ---------------------------------------------------------------------------------------------
static uint8_t rotl8(uint8_t b, int n) { return (uint8_t)((b << n) | (b >> (8 - n))); }
// src/aes.rs:24-62: S-box generated from the 3 / 3^-1 orbit in GF(2^8) plus the affine map
static void build_sbox(uint8_t sbox[256]) {
    uint8_t p = 1, q = 1;
    memset(sbox, 0, 256);
    do {
        p = (uint8_t)(p ^ (p << 1) ^ (((p >> 7) & 1) * 0x1B));
        q ^= (uint8_t)(q << 1);
        q ^= (uint8_t)(q << 2);
        q ^= (uint8_t)(q << 4);
        q ^= (uint8_t)(((q >> 7) & 1) * 0x09);
        uint8_t x = (uint8_t)(q ^ rotl8(q, 1) ^ rotl8(q, 2) ^ rotl8(q, 3) ^ rotl8(q, 4));
        sbox[p] = (uint8_t)(x ^ 0x63);
    } while (p != 1);
    sbox[0] = 0x63;
}
static uint8_t g_sbox[256];
static bool g_sbox_init = false;
static const uint8_t* sbox() {
    if (!g_sbox_init) {
        build_sbox(g_sbox);
        g_sbox_init = true;
    }
    return g_sbox;
}
// src/aes.rs:10-18
static void aes_add_round_key(const uint8_t* in, const uint8_t* key, uint8_t* out) {
    for (int i = 0; i < 16; ++i) out[i] = in[i] ^ key[i];
}
// src/aes.rs:64-81 (byte substitution only; the UInt128 witness there has no effect on the bytes)
static void aes_sub_bytes(const uint8_t* in, uint8_t* out) {
    for (int i = 0; i < 16; ++i) out[i] = sbox()[in[i]];
}
// src/aes.rs:96-151: column-major state, row r rotated left by r
static void aes_shift_rows(const uint8_t* in, uint8_t* out) {
    uint8_t m[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) m[i][j] = in[i + 4 * j];
    for (int r = 0; r < 4; ++r) {
        uint8_t t[4];
        for (int j = 0; j < 4; ++j) t[j] = m[r][(j + r) % 4];
        memcpy(m[r], t, 4);
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = m[j][i];
}
// src/aes.rs:153-174
static void aes_gmix_column(const uint8_t a[4], uint8_t o[4]) {
    uint8_t b[4];
    for (int i = 0; i < 4; ++i) {
        uint8_t h = (a[i] >> 7) & 1;
        b[i] = (uint8_t)((a[i] << 1) ^ (h * 0x1B));
    }
    o[0] = b[0] ^ a[3] ^ a[2] ^ b[1] ^ a[1];
    o[1] = b[1] ^ a[0] ^ a[3] ^ b[2] ^ a[2];
    o[2] = b[2] ^ a[1] ^ a[0] ^ b[3] ^ a[3];
    o[3] = b[3] ^ a[2] ^ a[1] ^ b[0] ^ a[0];
}
// src/aes.rs:176-196
static void aes_mix_columns(const uint8_t* in, uint8_t* out) {
    for (int c = 0; c < 4; ++c) aes_gmix_column(in + 4 * c, out + 4 * c);
}
// src/aes.rs:203-250
static void aes_derive_keys(const uint8_t key[16], uint8_t rk[11][16]) {
    static const uint8_t rcon[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36};
    uint32_t w[44];
    for (int i = 0; i < 4; ++i)
        w[i] = ((uint32_t)key[4 * i] << 24) | ((uint32_t)key[4 * i + 1] << 16) | ((uint32_t)key[4 * i + 2] << 8) | key[4 * i + 3];
    for (int i = 4; i < 44; ++i) {
        if (i % 4 == 0) {
            uint32_t t = w[i - 1];
            uint8_t by[4] = {(uint8_t)(t >> 16), (uint8_t)(t >> 8), (uint8_t)t, (uint8_t)(t >> 24)};  // rotate_word
            uint32_t s = ((uint32_t)sbox()[by[0]] << 24) | ((uint32_t)sbox()[by[1]] << 16) | ((uint32_t)sbox()[by[2]] << 8) |
                         sbox()[by[3]];
            w[i] = w[i - 4] ^ s ^ ((uint32_t)rcon[i / 4 - 1] << 24);
        } else {
            w[i] = w[i - 4] ^ w[i - 1];
        }
    }
    for (int r = 0; r < 11; ++r)
        for (int k = 0; k < 4; ++k) {
            uint32_t t = w[4 * r + k];
            rk[r][4 * k] = (uint8_t)(t >> 24);
            rk[r][4 * k + 1] = (uint8_t)(t >> 16);
            rk[r][4 * k + 2] = (uint8_t)(t >> 8);
            rk[r][4 * k + 3] = (uint8_t)t;
        }
}

// =============================================================================================
// C interface for ctypes (tests/, bench.py cpu_baseline)
// =============================================================================================
extern "C" {

int orc_threads() { return n_threads(); }
void orc_set_threads(int t) { g_threads = t; }

// field ops; which: 0 = Fr, 1 = Fq.  op: 0 add, 1 sub, 2 mul, 3 inverse(a), 4 to_mont(a), 5 from_mont(a), 6 neg(a)
int orc_field_op(int curve, int which, int op, const u64* a, const u64* b, u64* out, size_t count) {
    const Curve* C = curve_by_id(curve);
    if (!C) return -1;
    if (which == 0) {
        for (size_t i = 0; i < count; ++i) {
            Fr x, y = f_zero<4>(), r;
            memcpy(x.v, a + 4 * i, 32);
            if (b) memcpy(y.v, b + 4 * i, 32);
            switch (op) {
                case 0: r = f_add<4>(C->fr, x, y); break;
                case 1: r = f_sub<4>(C->fr, x, y); break;
                case 2: r = f_mul<4>(C->fr, x, y); break;
                case 3: r = f_inv<4>(C->fr, x); break;
                case 4: r = f_to_mont<4>(C->fr, x); break;
                case 5: r = f_from_mont<4>(C->fr, x); break;
                case 6: r = f_neg<4>(C->fr, x); break;
                default: return -2;
            }
            memcpy(out + 4 * i, r.v, 32);
        }
    } else {
        for (size_t i = 0; i < count; ++i) {
            Fq x, y = f_zero<6>(), r;
            memcpy(x.v, a + 6 * i, 48);
            if (b) memcpy(y.v, b + 6 * i, 48);
            switch (op) {
                case 0: r = f_add<6>(C->fq, x, y); break;
                case 1: r = f_sub<6>(C->fq, x, y); break;
                case 2: r = f_mul<6>(C->fq, x, y); break;
                case 3: r = f_inv<6>(C->fq, x); break;
                case 4: r = f_to_mont<6>(C->fq, x); break;
                case 5: r = f_from_mont<6>(C->fq, x); break;
                case 6: r = f_neg<6>(C->fq, x); break;
                default: return -2;
            }
            memcpy(out + 6 * i, r.v, 48);
        }
    }
    return 0;
}

// constants, for the CPU-side check of the product's generated tables: what = 0 Fr modulus, 1 Fq modulus,
// 2 Fr R, 3 Fq R, 4 Fr root of unity (mont), 5 Fr generator (mont), 6 G1 generator (96 B), 7 Fr R^2, 8 Fq R^2
int orc_constant(int curve, int what, u64* out) {
    const Curve* C = curve_by_id(curve);
    if (!C) return -1;
    switch (what) {
        case 0: memcpy(out, C->fr.mod, 32); return 4;
        case 1: memcpy(out, C->fq.mod, 48); return 6;
        case 2: memcpy(out, C->fr.one, 32); return 4;
        case 3: memcpy(out, C->fq.one, 48); return 6;
        case 4: memcpy(out, C->root.v, 32); return 4;
        case 5: memcpy(out, C->gen.v, 32); return 4;
        case 6: memcpy(out, C->gx.v, 48); memcpy(out + 6, C->gy.v, 48); return 12;
        case 7: memcpy(out, C->fr.r2, 32); return 4;
        case 8: memcpy(out, C->fq.r2, 48); return 6;
    }
    return -2;
}

int orc_g1_on_curve(int curve, const u64* pts, size_t n) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    for (size_t i = 0; i < n; ++i) {
        Aff a = load_aff(pts + 12 * i);
        if (a.inf) continue;
        Fq lhs = qm(C, a.y, a.y);
        Fq rhs = qa(C, qm(C, qm(C, a.x, a.x), a.x), C.b);
        if (!(lhs == rhs)) return 0;
    }
    return 1;
}

// out[i] = scalars[i] * G  (canonical scalars), affine 96 B each.  Fixture generator for tests.
int orc_g1_mul_gen(int curve, const u64* scalars, size_t n, u64* out) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    Aff g;
    g.x = C.gx;
    g.y = C.gy;
    g.inf = false;
    parallel_for((long long)n, 16, [&](long long i) {
        Jac r = scalar_mul(C, g, scalars + 4 * i);
        store_aff(out + 12 * i, jac_to_affine(C, r));
    });
    return 0;
}

// cheap synthetic bases for large fixtures: P_0 = k0*G, P_{i+1} = P_i + step*G  (all distinct, all on curve)
int orc_g1_walk(int curve, u64 k0, u64 step, size_t n, u64* out) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    Aff g;
    g.x = C.gx;
    g.y = C.gy;
    g.inf = false;
    u64 s0[4] = {k0, 0, 0, 0}, s1[4] = {step, 0, 0, 0};
    Jac p = scalar_mul(C, g, s0);
    Aff st = jac_to_affine(C, scalar_mul(C, g, s1));
    // batch normalisation in chunks (Montgomery's trick) to avoid one inversion per point
    const size_t CH = 1024;
    std::vector<Jac> buf(CH);
    std::vector<Fq> pre(CH);
    for (size_t base = 0; base < n; base += CH) {
        size_t m = std::min(CH, n - base);
        for (size_t i = 0; i < m; ++i) {
            buf[i] = p;
            jac_add_mixed(C, p, st);
        }
        Fq acc = f_one<6>(C.fq);
        for (size_t i = 0; i < m; ++i) {
            pre[i] = acc;
            acc = qm(C, acc, buf[i].z);
        }
        Fq inv = f_inv<6>(C.fq, acc);
        for (size_t i = m; i-- > 0;) {
            Fq zi = qm(C, inv, pre[i]);
            inv = qm(C, inv, buf[i].z);
            Fq zi2 = qm(C, zi, zi);
            Aff a;
            a.inf = false;
            a.x = qm(C, buf[i].x, zi2);
            a.y = qm(C, buf[i].y, qm(C, zi2, zi));
            store_aff(out + 12 * (base + i), a);
        }
    }
    return 0;
}

// MSM.  algo 0: ark-ec 0.3.0 Pippenger restatement; algo 1: sum of double-and-add (definition).
int orc_g1_msm(int curve, const u64* bases, const u64* scalars, size_t n, int algo, u64* out_affine) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    Jac r;
    if (algo == 0) {
        r = msm_pippenger(C, bases, scalars, n);
    } else {
        r = jac_zero(C);
        for (size_t i = 0; i < n; ++i) {
            Aff b = load_aff(bases + 12 * i);
            Jac t = scalar_mul(C, b, scalars + 4 * i);
            jac_add(C, r, t);
        }
    }
    store_aff(out_affine, jac_to_affine(C, r));
    return 0;
}

// point addition / negation helper for linearity tests: out = a + b (affine in, affine out)
int orc_g1_add(int curve, const u64* a, const u64* b, u64* out) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    Aff A = load_aff(a), B = load_aff(b);
    Jac r = jac_zero(C);
    jac_add_mixed(C, r, A);
    jac_add_mixed(C, r, B);
    store_aff(out, jac_to_affine(C, r));
    return 0;
}

// NTT in place.  data: n x 32 B Montgomery.  inverse / coset as in ark-poly's four entry points.
int orc_ntt(int curve, u64* data, int log_n, int inverse, int coset) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    if (log_n < 0 || log_n > C.two_adicity) return -2;
    size_t n = (size_t)1 << log_n;
    Fr* a = (Fr*)data;
    Fr w = domain_gen(C, log_n);
    if (!inverse) {
        if (coset) distribute_powers(C, a, n, C.gen);
        ntt_in_place(C, a, log_n, w);
    } else {
        Fr wi = f_inv<4>(C.fr, w);
        ntt_in_place(C, a, log_n, wi);
        Fr ninv = f_inv<4>(C.fr, f_from_u64<4>(C.fr, (u64)n));
        parallel_for((long long)n, 4096, [&](long long i) { a[i] = f_mul<4>(C.fr, a[i], ninv); });
        if (coset) distribute_powers(C, a, n, f_inv<4>(C.fr, C.gen));
    }
    return 0;
}

// O(n^2) evaluation of the polynomial with coefficients `in` at shift*w^i: the definition the NTT must match
int orc_dft_naive(int curve, const u64* in, u64* out, int log_n, int coset) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    size_t n = (size_t)1 << log_n;
    Fr w = domain_gen(C, log_n);
    const Fr* a = (const Fr*)in;
    Fr x = coset ? C.gen : f_one<4>(C.fr);
    for (size_t i = 0; i < n; ++i) {
        Fr acc = f_zero<4>();
        for (size_t k = n; k-- > 0;) acc = f_add<4>(C.fr, f_mul<4>(C.fr, acc, x), a[k]);  // Horner
        memcpy(out + 4 * i, acc.v, 32);
        x = f_mul<4>(C.fr, x, w);
    }
    return 0;
}


// ---------------------------------------------------------------------------------------------
// Vector primitives over Fr used by the Marlin oracle (oracle/marlin_oracle.py): the polynomial
// arithmetic that ark-poly 0.3.0 DensePolynomial / Evaluations and ark-ff batch_inversion provide
// to ark-marlin's prover rounds.  Plain loops; parallel where it is embarrassingly so.
// ---------------------------------------------------------------------------------------------
// op 0 add, 1 sub, 2 mul; b is broadcast when nb == 1
int orc_fr_vec(int curve, int op, const u64* a, size_t na, const u64* b, size_t nb, u64* out) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp || (nb != na && nb != 1)) return -1;
    const Curve& C = *Cp;
    const Fr* A = (const Fr*)a;
    const Fr* B = (const Fr*)b;
    Fr* O = (Fr*)out;
    parallel_for((long long)na, 4096, [&](long long i) {
        const Fr& y = B[nb == 1 ? 0 : i];
        O[i] = op == 0 ? f_add<4>(RF, A[i], y) : op == 1 ? f_sub<4>(RF, A[i], y) : f_mul<4>(RF, A[i], y);
    });
    return 0;
}
// ark-ff batch_inversion semantics: zeros are left as zeros
int orc_fr_batch_inv(int curve, const u64* a, u64* out, size_t n) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    const Fr* A = (const Fr*)a;
    Fr* O = (Fr*)out;
    const size_t CH = 1 << 14;
    parallel_for((long long)((n + CH - 1) / CH), 1, [&](long long c) {
        size_t lo = (size_t)c * CH, hi = std::min(n, lo + CH);
        std::vector<Fr> pre(hi - lo);
        Fr acc = f_one<4>(RF);
        for (size_t i = lo; i < hi; ++i) {
            pre[i - lo] = acc;
            if (!A[i].is_zero()) acc = f_mul<4>(RF, acc, A[i]);
        }
        Fr inv = f_inv<4>(RF, acc);
        for (size_t i = hi; i-- > lo;) {
            if (A[i].is_zero()) { O[i] = A[i]; continue; }
            Fr ai = A[i];
            O[i] = f_mul<4>(RF, inv, pre[i - lo]);
            inv = f_mul<4>(RF, inv, ai);
        }
    });
    return 0;
}
// Horner evaluation of sum c_i x^i
int orc_fr_poly_eval(int curve, const u64* coeffs, size_t n, const u64* x, u64* out) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    const Fr* A = (const Fr*)coeffs;
    Fr X;
    memcpy(X.v, x, 32);
    // chunked Horner: value = sum_c x^(c*CH) * horner(chunk c)
    const size_t CH = 1 << 14;
    size_t nch = (n + CH - 1) / CH;
    std::vector<Fr> part(nch ? nch : 1, f_zero<4>());
    parallel_for((long long)nch, 1, [&](long long c) {
        size_t lo = (size_t)c * CH, hi = std::min(n, lo + CH);
        Fr acc = f_zero<4>();
        for (size_t i = hi; i-- > lo;) acc = f_add<4>(RF, f_mul<4>(RF, acc, X), A[i]);
        part[c] = acc;
    });
    u64 e[1] = {CH};
    Fr xc = f_pow<4>(RF, X, e, 1);
    Fr acc = f_zero<4>();
    for (size_t c = nch; c-- > 0;) acc = f_add<4>(RF, f_mul<4>(RF, acc, xc), part[c]);
    memcpy(out, acc.v, 32);
    return 0;
}
// out[i] = c * base^i
int orc_fr_powers(int curve, const u64* base, const u64* c, size_t n, u64* out) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    Fr g, c0;
    memcpy(g.v, base, 32);
    memcpy(c0.v, c, 32);
    Fr* O = (Fr*)out;
    const size_t CH = 4096;
    parallel_for((long long)((n + CH - 1) / CH), 1, [&](long long ch) {
        size_t lo = (size_t)ch * CH, hi = std::min(n, lo + CH);
        u64 e[1] = {lo};
        Fr p = f_mul<4>(RF, c0, f_pow<4>(RF, g, e, 1));
        for (size_t i = lo; i < hi; ++i) {
            O[i] = p;
            p = f_mul<4>(RF, p, g);
        }
    });
    return 0;
}
// quotient of p(X) / (X - z) (synthetic division, remainder dropped): n coefficients in, n-1 out
int orc_fr_div_linear(int curve, const u64* coeffs, size_t n, const u64* z, u64* out) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    if (n < 2) return 0;
    const Fr* A = (const Fr*)coeffs;
    Fr* Q = (Fr*)out;
    Fr Z;
    memcpy(Z.v, z, 32);
    Fr acc = A[n - 1];
    Q[n - 2] = acc;
    for (size_t i = n - 1; i-- > 1;) {
        acc = f_add<4>(RF, f_mul<4>(RF, acc, Z), A[i]);
        Q[i - 1] = acc;
    }
    return 0;
}
// out[i] = scalars[i] * pts[i]   (canonical scalars, affine points in/out)
int orc_g1_mul(int curve, const u64* pts, const u64* scalars, size_t n, u64* out) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    parallel_for((long long)n, 4, [&](long long i) {
        Aff b = load_aff(pts + 12 * i);
        Jac r = scalar_mul(C, b, scalars + 4 * i);
        store_aff(out + 12 * i, jac_to_affine(C, r));
    });
    return 0;
}
// test SRS: out[i] = tau^i * G, i < n (tau canonical).  Fixed-base windows of 8 bits, like the product's srs.cu,
// but written independently: table T[w][d] = d * 2^(8w) * G built by repeated addition.
int orc_g1_srs(int curve, const u64* tau, size_t n, u64* out) {
    const Curve* Cp = curve_by_id(curve);
    if (!Cp) return -1;
    const Curve& C = *Cp;
    Aff g;
    g.x = C.gx; g.y = C.gy; g.inf = false;
    std::vector<Aff> table(32 * 256);
    Jac base = jac_zero(C);
    jac_add_mixed(C, base, g);
    for (int w = 0; w < 32; ++w) {
        Aff b = jac_to_affine(C, base);
        Jac acc = jac_zero(C);
        table[w * 256].inf = true;
        for (int d = 1; d < 256; ++d) {
            jac_add_mixed(C, acc, b);
            table[w * 256 + d] = jac_to_affine(C, acc);
        }
        for (int k = 0; k < 8; ++k) jac_double(C, base);
    }
    Fr t;
    memcpy(t.v, tau, 32);
    Fr tm = f_to_mont<4>(RF, t);
    const size_t CH = 256;
    parallel_for((long long)((n + CH - 1) / CH), 1, [&](long long ch) {
        size_t lo = (size_t)ch * CH, hi = std::min(n, lo + CH);
        u64 e[1] = {lo};
        Fr p = f_pow<4>(RF, tm, e, 1);
        std::vector<Jac> buf(hi - lo);
        for (size_t i = lo; i < hi; ++i) {
            Fr s = f_from_mont<4>(RF, p);
            Jac acc = jac_zero(C);
            for (int w = 0; w < 32; ++w) {
                unsigned d = (unsigned)((s.v[w >> 3] >> (8 * (w & 7))) & 0xff);
                if (d) jac_add_mixed(C, acc, table[w * 256 + d]);
            }
            buf[i - lo] = acc;
            p = f_mul<4>(RF, p, tm);
        }
        for (size_t i = lo; i < hi; ++i) store_aff(out + 12 * i, jac_to_affine(C, buf[i - lo]));
    });
    return 0;
}

// ---- AES (byte level) ----
void orc_aes_sbox(uint8_t out[256]) { memcpy(out, sbox(), 256); }
void orc_aes_add_round_key(const uint8_t* in, const uint8_t* key, uint8_t* out) { aes_add_round_key(in, key, out); }
void orc_aes_sub_bytes(const uint8_t* in, uint8_t* out) { aes_sub_bytes(in, out); }
void orc_aes_shift_rows(const uint8_t* in, uint8_t* out) { aes_shift_rows(in, out); }
void orc_aes_mix_columns(const uint8_t* in, uint8_t* out) { aes_mix_columns(in, out); }
void orc_aes_derive_keys(const uint8_t* key, uint8_t* out176) { aes_derive_keys(key, (uint8_t(*)[16])out176); }
// ECB over whole 16-byte blocks, round structure of src/lib.rs:194-277.  Optionally dumps the per-round trace:
// trace (if non-null) receives per block 41 x 16 bytes: state after round-0 ARK, then for rounds 1..9
// [sub, shift, mix, ark], then round 10 [sub, shift, ark]... laid out as 1 + 9*4 + 3 = 40 states (+1 unused).
int orc_aes128_ecb(const uint8_t* msg, size_t len, const uint8_t* key, uint8_t* ct, uint8_t* trace) {
    if (len % 16) return -1;
    uint8_t rk[11][16];
    aes_derive_keys(key, rk);
    for (size_t b = 0; b < len / 16; ++b) {
        uint8_t s[16], t[16];
        uint8_t* tr = trace ? trace + b * 40 * 16 : nullptr;
        int k = 0;
        aes_add_round_key(msg + 16 * b, key, s);
        if (tr) memcpy(tr + 16 * k++, s, 16);
        for (int r = 1; r <= 9; ++r) {
            aes_sub_bytes(s, t);
            if (tr) memcpy(tr + 16 * k++, t, 16);
            aes_shift_rows(t, s);
            if (tr) memcpy(tr + 16 * k++, s, 16);
            aes_mix_columns(s, t);
            if (tr) memcpy(tr + 16 * k++, t, 16);
            aes_add_round_key(t, rk[r], s);
            if (tr) memcpy(tr + 16 * k++, s, 16);
        }
        aes_sub_bytes(s, t);
        if (tr) memcpy(tr + 16 * k++, t, 16);
        aes_shift_rows(t, s);
        if (tr) memcpy(tr + 16 * k++, s, 16);
        aes_add_round_key(s, rk[10], t);
        if (tr) memcpy(tr + 16 * k++, t, 16);
        memcpy(ct + 16 * b, t, 16);
    }
    return 0;
}

}  // extern "C"
