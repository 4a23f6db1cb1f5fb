// Fr vector kernels of the Marlin prover rounds (see polyops.cuh).  Every kernel is a single coalesced pass over its
// operands with 2 x 128-bit loads/stores per element; grids are sized from the element count (>= 148 x resident CTAs
// for every vector of the prover's sizes), so these run at HBM speed: bytes = 32 x (inputs + outputs) x n.
#include "polyops.cuh"

#include <algorithm>
#include <vector>

namespace zk {

namespace {
using F = FrS;
constexpr int TB = 256;

__device__ __forceinline__ F ld_fr(const F* p) {
    F r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(F* p, const F& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// v * x for a small integer v
__device__ __forceinline__ F mul_small(const F& x, int v) {
    int a = v < 0 ? -v : v;
    F t;
    if (a == 1) t = x;
    else if (a == 2) t = x.dbl();
    else if (a == 0) t = F::zero();
    else t = x * F::from_u64((uint64_t)a);
    return v < 0 ? t.neg() : t;
}
__device__ __forceinline__ F small_to_fr(int v) {
    if (v == 0) return F::zero();
    if (v == 1) return F::one();
    return mul_small(F::one(), v);
}

__global__ void k_powers(F* out, size_t n, F base, F c) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long e = i;
    F r = c, b = base;
    while (e) {
        if (e & 1) r = r * b;
        b = b.sqr();
        e >>= 1;
    }
    st_fr(out + i, r);
}
__global__ void k_fill(F* out, size_t n, F v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(out + i, v);
}
__global__ void k_vec(int op, F* out, const F* a, const F* b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = ld_fr(a + i), y = ld_fr(b + i);
    st_fr(out + i, op == 0 ? x + y : op == 1 ? x - y : x * y);
}
__global__ void k_scale(F* out, const F* a, F s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(out + i, ld_fr(a + i) * s);
}
__global__ void k_axpy(F* acc, const F* x, F s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(acc + i, ld_fr(acc + i) + ld_fr(x + i) * s);
}
__global__ void k_rsub_scalar(F* out, const F* a, F s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(out + i, s - ld_fr(a + i));
}
// Montgomery's trick on runs of INV_CH consecutive elements per thread: 3 products per element + one Fermat inverse per run
constexpr int INV_CH = 64;
__global__ void __launch_bounds__(128) k_batch_inverse(F* out, const F* in, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * INV_CH;
    if (lo >= n) return;
    size_t hi = lo + INV_CH < n ? lo + INV_CH : n;
    F acc = F::one();
    for (size_t i = lo; i < hi; ++i) {
        st_fr(out + i, acc);  // prefix product of the non-zero entries before i
        F a = ld_fr(in + i);
        if (!a.is_zero()) acc = acc * a;
    }
    F inv = acc.inverse();
    for (size_t i = hi; i-- > lo;) {
        F a = ld_fr(in + i);
        if (a.is_zero()) {
            st_fr(out + i, a);
            continue;
        }
        F pre = ld_fr(out + i);
        st_fr(out + i, inv * pre);
        inv = inv * a;
    }
}
__global__ void k_gather(F* out, const F* table, const uint32_t* idx, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(out + i, ld_fr(table + idx[i]));
}
__global__ void k_gather_scaled(F* out, const F* table, const uint32_t* idx, const int8_t* coeff, F s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = coeff[i];
    if (v == 0) {
        st_fr(out + i, F::zero());
        return;
    }
    st_fr(out + i, mul_small(ld_fr(table + idx[i]) * s, v));
}
__global__ void k_gather_fma(F* acc, const F* table, const uint32_t* idx, const int8_t* coeff, const F* w, F s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = coeff[i];
    if (v == 0) return;
    F t = mul_small(ld_fr(table + idx[i]) * s, v) * ld_fr(w + i);
    st_fr(acc + i, ld_fr(acc + i) + t);
}
__global__ void k_den_k(F* out, const F* table, const uint32_t* ridx, const uint32_t* cidx, F alpha, F beta, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(out + i, (beta - ld_fr(table + ridx[i])) * (alpha - ld_fr(table + cidx[i])));
}
__global__ void k_spmv_bits(F* out, const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, const int8_t* __restrict__ coeff,
                            const uint8_t* __restrict__ z, size_t nrows, size_t n_out) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_out) return;
    int s = 0;
    if (r < nrows)
        for (uint32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e) s += (int)coeff[e] * (int)z[col[e]];
    st_fr(out + r, small_to_fr(s));
}
// the same row sums as plain integers (the Lagrange-basis commitment of z_A / z_B takes them as one signed digit each)
__global__ void k_spmv_bits_i32(int32_t* out, const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, const int8_t* __restrict__ coeff,
                                const uint8_t* __restrict__ z, size_t nrows, size_t n_out) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_out) return;
    int s = 0;
    if (r < nrows)
        for (uint32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e) s += (int)coeff[e] * (int)z[col[e]];
    out[r] = s;
}
// the full assignment in H order: position j * ratio holds instance variable j, the positions in between the witness (zero padded)
__global__ void k_assignment_h_i32(int32_t* out, const uint8_t* __restrict__ z, size_t h, size_t ratio, size_t num_instance, size_t num_witness) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= h) return;
    int v;
    if (k % ratio == 0) {
        const size_t j = k / ratio;
        v = j < num_instance ? (int)z[j] : 0;
    } else {
        const size_t wi = k - k / ratio - 1;
        v = wi < num_witness ? (int)z[num_instance + wi] : 0;
    }
    out[k] = v;
}
__global__ void k_scale_strided(F* out, const F* a, F s, size_t stride, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) st_fr(out + i * stride, ld_fr(a + i * stride) * s);
}
__global__ void k_w_evals(F* out, const uint8_t* __restrict__ z, const F* x_evals, size_t h, size_t ratio, size_t num_instance,
                          size_t num_witness) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= h) return;
    if (k % ratio == 0) {
        st_fr(out + k, F::zero());
        return;
    }
    size_t wi = k - k / ratio - 1;
    F w = (wi < num_witness && z[num_instance + wi]) ? F::one() : F::zero();
    st_fr(out + k, w - ld_fr(x_evals + k));
}
__global__ void k_add_vanishing(F* c, size_t n, F r) {
    st_fr(c, ld_fr(c) - r);
    st_fr(c + n, ld_fr(c + n) + r);
}
__global__ void k_divide_vanishing(const F* c, size_t len, size_t n, F* q, F* rem) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t m = (len + n - 1) / n;
    F acc = F::zero();
    for (size_t j = m; j-- > 1;) {
        size_t idx = j * n + i;
        if (idx < len) acc = acc + ld_fr(c + idx);
        if (idx - n < len - n) st_fr(q + (idx - n), acc);
    }
    st_fr(rem + i, i < len ? acc + ld_fr(c + i) : acc);
}
__device__ __forceinline__ size_t reindex_by_subdomain(size_t j, size_t period, size_t x) {  // EvaluationDomain::reindex_by_subdomain
    if (j < x) return j * period;
    size_t i = j - x;
    return i + i / (period - 1) + 1;
}
__device__ __forceinline__ F t_term(const F& t, int v) {
    if (v == 1) return t;
    if (v == -1) return t.neg();
    return mul_small(t, v);
}
// light columns: one thread per variable
__global__ void __launch_bounds__(128) k_t_evals(F* out, CscView a, CscView b, CscView c, F eta_a, F eta_b, F eta_c, const F* r_alpha,
                                                 const uint8_t* __restrict__ heavy, size_t nvar, size_t period, size_t x) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nvar || heavy[j]) return;
    F tot = F::zero();
    const CscView* ms[3] = {&a, &b, &c};
    const F etas[3] = {eta_a, eta_b, eta_c};
#pragma unroll
    for (int mi = 0; mi < 3; ++mi) {
        const CscView& m = *ms[mi];
        const uint32_t lo = m.ptr[j], hi = m.ptr[j + 1];
        if (hi == lo) continue;
        F s = F::zero();
        for (uint32_t e = lo; e < hi; ++e) s = s + t_term(ld_fr(r_alpha + m.row[e]), m.coeff[e]);
        tot = tot + s * etas[mi];
    }
    st_fr(out + reindex_by_subdomain(j, period, x), tot);
}
// heavy columns (key and round-key bits, the wires every ECB block touches: 165,249 columns of ~450 entries at 4 KiB): one CTA per variable
__global__ void __launch_bounds__(256) k_t_evals_heavy(F* out, CscView a, CscView b, CscView c, F eta_a, F eta_b, F eta_c, const F* r_alpha,
                                                       const uint32_t* __restrict__ heavy_cols, size_t period, size_t x) {
    __shared__ F sh[256];
    const size_t j = heavy_cols[blockIdx.x];
    F tot = F::zero();
    const CscView* ms[3] = {&a, &b, &c};
    const F etas[3] = {eta_a, eta_b, eta_c};
#pragma unroll
    for (int mi = 0; mi < 3; ++mi) {
        const CscView& m = *ms[mi];
        const uint32_t lo = m.ptr[j], hi = m.ptr[j + 1];
        if (hi == lo) continue;
        F s = F::zero();
        for (uint32_t e = lo + threadIdx.x; e < hi; e += blockDim.x) s = s + t_term(ld_fr(r_alpha + m.row[e]), m.coeff[e]);
        tot = tot + s * etas[mi];
    }
    sh[threadIdx.x] = tot;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fr(out + reindex_by_subdomain(j, period, x), sh[0]);
}
// giant columns (the constant one: 25.6 M entries over A, B, C at 4 KiB): with one CTA the column alone took 110 ms per proof
// (profiles/r2_launches_bench_4k.txt, k_t_evals_heavy).  A giant column is cut into chunks of T_GIANT_CHUNK entries per matrix: CTA
// (column, chunk) sums its chunk, a second kernel adds the column's partial sums.
static constexpr uint32_t T_GIANT_CHUNK = 8192;
__global__ void __launch_bounds__(256) k_t_evals_giant(F* partials, CscView a, CscView b, CscView c, F eta_a, F eta_b, F eta_c, const F* r_alpha,
                                                       const uint32_t* __restrict__ giant_cols) {
    __shared__ F sh[256];
    const size_t j = giant_cols[blockIdx.x];
    const uint32_t s0 = blockIdx.y * T_GIANT_CHUNK;
    F tot = F::zero();
    const CscView* ms[3] = {&a, &b, &c};
    const F etas[3] = {eta_a, eta_b, eta_c};
    bool any = false;
#pragma unroll
    for (int mi = 0; mi < 3; ++mi) {
        const CscView& m = *ms[mi];
        const uint32_t end = m.ptr[j + 1];
        if (end - m.ptr[j] <= s0) continue;
        const uint32_t lo = m.ptr[j] + s0, hi = end - lo > T_GIANT_CHUNK ? lo + T_GIANT_CHUNK : end;
        any = true;
        F s = F::zero();
        for (uint32_t e = lo + threadIdx.x; e < hi; e += blockDim.x) s = s + t_term(ld_fr(r_alpha + m.row[e]), m.coeff[e]);
        tot = tot + s * etas[mi];
    }
    if (!any) return;  // block-uniform: this chunk lies past the end of the column in all three matrices
    sh[threadIdx.x] = tot;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fr(partials + (size_t)blockIdx.x * gridDim.y + blockIdx.y, sh[0]);
}
__global__ void __launch_bounds__(256) k_t_evals_giant_sum(F* out, const F* partials, uint32_t chunks_max, CscView a, CscView b, CscView c,
                                                           const uint32_t* __restrict__ giant_cols, size_t period, size_t x) {
    __shared__ F sh[256];
    const size_t j = giant_cols[blockIdx.x];
    uint32_t len = a.ptr[j + 1] - a.ptr[j];
    len = max(len, b.ptr[j + 1] - b.ptr[j]);
    len = max(len, c.ptr[j + 1] - c.ptr[j]);
    const uint32_t chunks = (len + T_GIANT_CHUNK - 1) / T_GIANT_CHUNK;  // the chunks k_t_evals_giant wrote
    F tot = F::zero();
    for (uint32_t s = threadIdx.x; s < chunks; s += blockDim.x) tot = tot + ld_fr(partials + (size_t)blockIdx.x * chunks_max + s);
    sh[threadIdx.x] = tot;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fr(out + reindex_by_subdomain(j, period, x), sh[0]);
}
__global__ void k_round2(F* out, const F* ra, const F* za, const F* zb, const F* t, const F* z, F eta_a, F eta_b, F eta_c, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F a = ld_fr(za + i), b = ld_fr(zb + i);
    F s = (eta_a + eta_c * b) * a + eta_b * b;
    st_fr(out + i, ld_fr(ra + i) * s - ld_fr(t + i) * ld_fr(z + i));
}
__global__ void k_den_coset(F* row, const F* col, const F* rc, F alpha, F beta, F ab, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_fr(row + i, ab - alpha * ld_fr(row + i) - beta * ld_fr(col + i) + ld_fr(rc + i));
}
struct R3Args {
    const F* val[3];
    const F* den[3];
    F eta[3];
    F vv;
    F vkinv[4];
};
__global__ void k_round3(F* out, R3Args p, const F* f, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F da = ld_fr(p.den[0] + i), db = ld_fr(p.den[1] + i), dc = ld_fr(p.den[2] + i);
    F bc = db * dc, ac = da * dc, ab = da * db;
    F a = p.eta[0] * ld_fr(p.val[0] + i) * bc + p.eta[1] * ld_fr(p.val[1] + i) * ac + p.eta[2] * ld_fr(p.val[2] + i) * ab;
    F num = a * p.vv - ab * dc * ld_fr(f + i);
    st_fr(out + i, num * p.vkinv[i & 3]);
}
// out[a] = in[a] * lo[a & 1023] * hi[a >> 10]
__global__ void k_scale_powers(F* out, const F* in, const F* __restrict__ lo, const F* __restrict__ hi, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F a = ld_fr(in + i) * ld_fr(lo + (i & 1023));
    if (i >> 10) a = a * ld_fr(hi + (i >> 10));
    st_fr(out + i, a);
}
__global__ void k_lincomb_den(F* out, const F* a, const F* b, const F* c, F sa, F sb, F k0, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F v = ld_fr(a + i) * sa + ld_fr(b + i) * sb + ld_fr(c + i);
    if (i == 0) v = v + k0;
    st_fr(out + i, v);
}
__global__ void k_mul3(F* out, const F* a, const F* b, const F* c, F s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(out + i, ld_fr(a + i) * ld_fr(b + i) * ld_fr(c + i) * s);
}
__global__ void k_fma3(F* acc, const F* a, const F* b, const F* c, F s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(acc + i, ld_fr(acc + i) + ld_fr(a + i) * ld_fr(b + i) * ld_fr(c + i) * s);
}
// four coset interpolants R_j = sum_b c_b i^(j b) (i = a primitive 4th root of unity) -> c_b = 1/4 sum_j w^(j b) R_j, w = 1/i, in place
__global__ void k_coset4_combine(F* v, size_t k, F w, F quarter) {
    size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= k) return;
    F r0 = ld_fr(v + a), r1 = ld_fr(v + k + a), r2 = ld_fr(v + 2 * k + a), r3 = ld_fr(v + 3 * k + a);
    F s02 = r0 + r2, d02 = r0 - r2, s13 = r1 + r3, d13 = (r1 - r3) * w;
    st_fr(v + a, (s02 + s13) * quarter);
    st_fr(v + k + a, (d02 + d13) * quarter);
    st_fr(v + 2 * k + a, (s02 - s13) * quarter);
    st_fr(v + 3 * k + a, (d02 - d13) * quarter);
}
// three coset interpolants p_j = c_0 + u_j c_1 + u_j^2 c_2 (block 3 of the quotient is zero) -> c_b = sum_j m[3 b + j] p_j, in place
struct Mat3 { F m[9]; };
__global__ void k_coset3_combine(F* v, size_t k, Mat3 M) {
    size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= k) return;
    F p0 = ld_fr(v + a), p1 = ld_fr(v + k + a), p2 = ld_fr(v + 2 * k + a);
    st_fr(v + a, p0 * M.m[0] + p1 * M.m[1] + p2 * M.m[2]);
    st_fr(v + k + a, p0 * M.m[3] + p1 * M.m[4] + p2 * M.m[5]);
    st_fr(v + 2 * k + a, p0 * M.m[6] + p1 * M.m[7] + p2 * M.m[8]);
}
// partial[c] = sum_{i < CH} coeffs[c*CH + i] * x^i
constexpr int EV_CH = 256;
__global__ void __launch_bounds__(128) k_eval_partial(const F* coeffs, size_t n, F x, F* partial) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = c * EV_CH;
    if (lo >= n) return;
    size_t hi = lo + EV_CH < n ? lo + EV_CH : n;
    F acc = F::zero();
    for (size_t i = hi; i-- > lo;) acc = acc * x + ld_fr(coeffs + i);
    st_fr(partial + c, acc);
}
constexpr int DL_CH = 2048;
__global__ void __launch_bounds__(128) k_chunk_horner(const F* coeffs, size_t n, F x, F* partial) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = c * DL_CH;
    if (lo >= n) return;
    size_t hi = lo + DL_CH < n ? lo + DL_CH : n;
    F acc = F::zero();
    for (size_t i = hi; i-- > lo;) acc = acc * x + ld_fr(coeffs + i);
    st_fr(partial + c, acc);
}
// q[i-1] = sum_{j >= i} c[j] z^(j-i): per chunk, starting from the carry of everything above it
__global__ void __launch_bounds__(128) k_div_linear(const F* coeffs, size_t n, F z, const F* carry, F* q) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = c * DL_CH;
    if (lo >= n) return;
    size_t hi = lo + DL_CH < n ? lo + DL_CH : n;
    F acc = ld_fr(carry + c);
    for (size_t i = hi; i-- > lo;) {
        acc = acc * z + ld_fr(coeffs + i);
        if (i >= 1) st_fr(q + (i - 1), acc);
    }
}
__global__ void k_z_poly(F* out, const F* w, size_t len_w, const F* xp, size_t x) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len_w + x) return;
    F v = F::zero();
    if (i >= x) v = ld_fr(w + (i - x));
    if (i < len_w) v = v - ld_fr(w + i);
    if (i < x) v = v + ld_fr(xp + i);
    st_fr(out + i, v);
}
__global__ void k_mask_fix(F* c, size_t n) { st_fr(c, (ld_fr(c + n) + ld_fr(c + 2 * n)).neg()); }
__global__ void k_bits_to_fr(F* out, const uint8_t* bits, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(out + i, bits[i] ? F::one() : F::zero());
}

#define LAUNCH(ctx, kernel, n, threads, ...)                                              \
    do {                                                                                  \
        if ((n) > 0) {                                                                    \
            kernel<<<cdiv((n), (threads)), (threads), 0, (ctx)->stream>>>(__VA_ARGS__);   \
            (ctx)->launches++;                                                            \
            ZK_CUDA((ctx), cudaGetLastError());                                           \
        }                                                                                 \
    } while (0)

}  // namespace

int po_powers(zkaes_ctx* ctx, F* out, size_t n, const F& base, const F& c) { LAUNCH(ctx, k_powers, n, TB, out, n, base, c); return ZK_OK; }
int po_fill(zkaes_ctx* ctx, F* out, size_t n, const F& v) { LAUNCH(ctx, k_fill, n, TB, out, n, v); return ZK_OK; }
int po_vec(zkaes_ctx* ctx, int op, F* out, const F* a, const F* b, size_t n) { LAUNCH(ctx, k_vec, n, TB, op, out, a, b, n); return ZK_OK; }
int po_scale(zkaes_ctx* ctx, F* out, const F* a, const F& s, size_t n) { LAUNCH(ctx, k_scale, n, TB, out, a, s, n); return ZK_OK; }
int po_axpy(zkaes_ctx* ctx, F* acc, const F* x, const F& s, size_t n) { LAUNCH(ctx, k_axpy, n, TB, acc, x, s, n); return ZK_OK; }
int po_rsub_scalar(zkaes_ctx* ctx, F* out, const F* a, const F& s, size_t n) { LAUNCH(ctx, k_rsub_scalar, n, TB, out, a, s, n); return ZK_OK; }
int po_batch_inverse(zkaes_ctx* ctx, F* out, const F* in, size_t n) {
    size_t runs = (n + INV_CH - 1) / INV_CH;
    LAUNCH(ctx, k_batch_inverse, runs, 128, out, in, n);
    return ZK_OK;
}
int po_gather(zkaes_ctx* ctx, F* out, const F* table, const uint32_t* idx, size_t n) { LAUNCH(ctx, k_gather, n, TB, out, table, idx, n); return ZK_OK; }
int po_gather_scaled(zkaes_ctx* ctx, F* out, const F* table, const uint32_t* idx, const int8_t* coeff, const F& s, size_t n) {
    LAUNCH(ctx, k_gather_scaled, n, TB, out, table, idx, coeff, s, n);
    return ZK_OK;
}
int po_gather_fma(zkaes_ctx* ctx, F* acc, const F* table, const uint32_t* idx, const int8_t* coeff, const F* w, const F& s, size_t n) {
    LAUNCH(ctx, k_gather_fma, n, TB, acc, table, idx, coeff, w, s, n);
    return ZK_OK;
}
int po_den_k(zkaes_ctx* ctx, F* out, const F* table, const uint32_t* ridx, const uint32_t* cidx, const F& alpha, const F& beta, size_t n) {
    LAUNCH(ctx, k_den_k, n, TB, out, table, ridx, cidx, alpha, beta, n);
    return ZK_OK;
}
int po_spmv_bits(zkaes_ctx* ctx, F* out, const uint32_t* row_ptr, const uint32_t* col, const int8_t* coeff, const uint8_t* z, size_t nrows,
                 size_t n_out) {
    LAUNCH(ctx, k_spmv_bits, n_out, TB, out, row_ptr, col, coeff, z, nrows, n_out);
    return ZK_OK;
}
int po_spmv_bits_i32(zkaes_ctx* ctx, int32_t* out, const uint32_t* row_ptr, const uint32_t* col, const int8_t* coeff, const uint8_t* z, size_t nrows,
                     size_t n_out) {
    LAUNCH(ctx, k_spmv_bits_i32, n_out, TB, out, row_ptr, col, coeff, z, nrows, n_out);
    return ZK_OK;
}
int po_assignment_h_i32(zkaes_ctx* ctx, int32_t* out, const uint8_t* z, size_t h, size_t ratio, size_t num_instance, size_t num_witness) {
    LAUNCH(ctx, k_assignment_h_i32, h, TB, out, z, h, ratio, num_instance, num_witness);
    return ZK_OK;
}
int po_scale_strided(zkaes_ctx* ctx, F* out, const F* a, const F& s, size_t stride, size_t count) {
    LAUNCH(ctx, k_scale_strided, count, TB, out, a, s, stride, count);
    return ZK_OK;
}
int po_w_evals(zkaes_ctx* ctx, F* out, const uint8_t* z, const F* x_evals, size_t h, size_t ratio, size_t num_instance, size_t num_witness) {
    LAUNCH(ctx, k_w_evals, h, TB, out, z, x_evals, h, ratio, num_instance, num_witness);
    return ZK_OK;
}
int po_add_vanishing(zkaes_ctx* ctx, F* c, size_t n, const F& r) {
    k_add_vanishing<<<1, 1, 0, ctx->stream>>>(c, n, r);
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    return ZK_OK;
}
int po_divide_vanishing(zkaes_ctx* ctx, const F* c, size_t len, size_t n, F* q, F* rem) {
    LAUNCH(ctx, k_divide_vanishing, n, TB, c, len, n, q, rem);
    return ZK_OK;
}
int po_t_evals(zkaes_ctx* ctx, F* out, const CscView m[3], const F eta[3], const F* r_alpha, const uint8_t* heavy_flag,
               const uint32_t* heavy_cols, size_t n_heavy, const uint32_t* giant_cols, size_t n_giant, size_t giant_max_len, size_t nvar, size_t h,
               size_t x) {
    cudaStream_t st = ctx->stream;
    ZK_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(F) * h, st));
    LAUNCH(ctx, k_t_evals, nvar, 128, out, m[0], m[1], m[2], eta[0], eta[1], eta[2], r_alpha, heavy_flag, nvar, h / x, x);
    if (n_heavy) {
        k_t_evals_heavy<<<(unsigned)n_heavy, 256, 0, st>>>(out, m[0], m[1], m[2], eta[0], eta[1], eta[2], r_alpha, heavy_cols, h / x, x);
        ctx->launches++;
        ZK_CUDA(ctx, cudaGetLastError());
    }
    if (n_giant) {
        const unsigned chunks = (unsigned)std::max<size_t>((giant_max_len + T_GIANT_CHUNK - 1) / T_GIANT_CHUNK, 1);
        if (chunks > 65535) return fail(ctx, ZK_ERR_UNSUPPORTED, "t_evals: a matrix column with more than 2^29 entries");
        DevBuf partials;
        ZK_CUDA(ctx, partials.alloc(sizeof(F) * n_giant * chunks, st));
        k_t_evals_giant<<<dim3((unsigned)n_giant, chunks), 256, 0, st>>>(partials.as<F>(), m[0], m[1], m[2], eta[0], eta[1], eta[2], r_alpha, giant_cols);
        k_t_evals_giant_sum<<<(unsigned)n_giant, 256, 0, st>>>(out, partials.as<F>(), chunks, m[0], m[1], m[2], giant_cols, h / x, x);
        ctx->launches += 2;
        ZK_CUDA(ctx, cudaGetLastError());
    }
    return ZK_OK;
}
int po_round2(zkaes_ctx* ctx, F* out, const F* ra, const F* za, const F* zb, const F* t, const F* z, const F eta[3], size_t n) {
    LAUNCH(ctx, k_round2, n, TB, out, ra, za, zb, t, z, eta[0], eta[1], eta[2], n);
    return ZK_OK;
}
int po_den_coset(zkaes_ctx* ctx, F* row, const F* col, const F* rc, const F& alpha, const F& beta, const F& ab, size_t n) {
    LAUNCH(ctx, k_den_coset, n, TB, row, col, rc, alpha, beta, ab, n);
    return ZK_OK;
}
int po_round3(zkaes_ctx* ctx, F* out, const F* const val[3], const F* const den[3], const F* f, const F eta[3], const F& vv, const F vkinv[4],
              size_t n) {
    R3Args p;
    for (int i = 0; i < 3; ++i) {
        p.val[i] = val[i];
        p.den[i] = den[i];
        p.eta[i] = eta[i];
    }
    p.vv = vv;
    for (int i = 0; i < 4; ++i) p.vkinv[i] = vkinv[i];
    LAUNCH(ctx, k_round3, n, TB, out, p, f, n);
    return ZK_OK;
}
int po_scale_powers(zkaes_ctx* ctx, F* out, const F* in, const F& base, size_t n) {
    cudaStream_t st = ctx->stream;
    DevBuf lo, hi;
    size_t nhi = (n >> 10) + 1;
    ZK_CUDA(ctx, lo.alloc(sizeof(F) * 1024, st));
    ZK_CUDA(ctx, hi.alloc(sizeof(F) * nhi, st));
    F b1024 = base;
    for (int i = 0; i < 10; ++i) b1024 = b1024.sqr();
    ZK_TRY(po_powers(ctx, lo.as<F>(), 1024, base, F::one()));
    ZK_TRY(po_powers(ctx, hi.as<F>(), nhi, b1024, F::one()));
    LAUNCH(ctx, k_scale_powers, n, TB, out, in, lo.as<F>(), hi.as<F>(), n);
    return ZK_OK;
}
int po_lincomb_den(zkaes_ctx* ctx, F* out, const F* a, const F* b, const F* c, const F& sa, const F& sb, const F& k0, size_t n) {
    LAUNCH(ctx, k_lincomb_den, n, TB, out, a, b, c, sa, sb, k0, n);
    return ZK_OK;
}
int po_mul3(zkaes_ctx* ctx, F* out, const F* a, const F* b, const F* c, const F& s, size_t n) { LAUNCH(ctx, k_mul3, n, TB, out, a, b, c, s, n); return ZK_OK; }
int po_fma3(zkaes_ctx* ctx, F* acc, const F* a, const F* b, const F* c, const F& s, size_t n) { LAUNCH(ctx, k_fma3, n, TB, acc, a, b, c, s, n); return ZK_OK; }
int po_coset4_combine(zkaes_ctx* ctx, F* v, size_t k, const F& i4_inv) {
    LAUNCH(ctx, k_coset4_combine, k, TB, v, k, i4_inv, F::from_u64(4).inverse());
    return ZK_OK;
}
int po_coset3_combine(zkaes_ctx* ctx, F* v, size_t k, const F u[3]) {
    // inverse Vandermonde through the Lagrange basis on the nodes u_j: l_j(y) = (y - u_a)(y - u_b) / ((u_j - u_a)(u_j - u_b))
    Mat3 M;
    for (int j = 0; j < 3; ++j) {
        const F& ua = u[(j + 1) % 3];
        const F& ub = u[(j + 2) % 3];
        F d = ((u[j] - ua) * (u[j] - ub)).inverse();
        M.m[0 + j] = ua * ub * d;
        M.m[3 + j] = (ua + ub).neg() * d;
        M.m[6 + j] = d;
    }
    LAUNCH(ctx, k_coset3_combine, k, TB, v, k, M);
    return ZK_OK;
}
int po_eval(zkaes_ctx* ctx, const F* coeffs, size_t n, const F& x, F* out_host) {
    if (n == 0) {
        *out_host = F::zero();
        return ZK_OK;
    }
    cudaStream_t st = ctx->stream;
    // level 1 on the device, then the (n / 256)-term polynomial in y = x^256 on the host
    size_t np = (n + EV_CH - 1) / EV_CH;
    DevBuf part;
    ZK_CUDA(ctx, part.alloc(sizeof(F) * np, st));
    LAUNCH(ctx, k_eval_partial, np, 128, coeffs, n, x, part.as<F>());
    F y = x;
    for (int i = 0; i < 8; ++i) y = y.sqr();
    const F* cur = part.as<F>();
    size_t cn = np;
    DevBuf part2;
    if (cn > 4096) {  // second device level for the largest polynomials
        size_t np2 = (cn + EV_CH - 1) / EV_CH;
        ZK_CUDA(ctx, part2.alloc(sizeof(F) * np2, st));
        LAUNCH(ctx, k_eval_partial, np2, 128, cur, cn, y, part2.as<F>());
        for (int i = 0; i < 8; ++i) y = y.sqr();
        cur = part2.as<F>();
        cn = np2;
    }
    std::vector<F> hp(cn);
    ZK_CUDA(ctx, cudaMemcpyAsync(hp.data(), cur, sizeof(F) * cn, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    F acc = F::zero();
    for (size_t i = cn; i-- > 0;) acc = acc * y + hp[i];
    *out_host = acc;
    return ZK_OK;
}
int po_div_linear(zkaes_ctx* ctx, const F* c, size_t n, const F& z, F* q) {
    if (n < 2) return ZK_OK;
    cudaStream_t st = ctx->stream;
    size_t nc = (n + DL_CH - 1) / DL_CH;
    DevBuf part, carry;
    ZK_CUDA(ctx, part.alloc(sizeof(F) * nc, st));
    ZK_CUDA(ctx, carry.alloc(sizeof(F) * nc, st));
    LAUNCH(ctx, k_chunk_horner, nc, 128, c, n, z, part.as<F>());
    std::vector<F> hp(nc), hc(nc);
    ZK_CUDA(ctx, cudaMemcpyAsync(hp.data(), part.p, sizeof(F) * nc, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    // carry into chunk c = Horner value of all higher chunks: H_c = P_{c+1} + z^CH * H_{c+1}
    F zc = z;
    for (int i = 0; i < 11; ++i) zc = zc.sqr();  // z^2048
    static_assert(DL_CH == 2048, "update the exponent");
    F acc = F::zero();
    for (size_t k = nc; k-- > 0;) {
        hc[k] = acc;
        acc = acc * zc + hp[k];
    }
    ZK_CUDA(ctx, cudaMemcpyAsync(carry.p, hc.data(), sizeof(F) * nc, cudaMemcpyHostToDevice, st));
    LAUNCH(ctx, k_div_linear, nc, 128, c, n, z, carry.as<F>(), q);
    ZK_CUDA(ctx, cudaStreamSynchronize(st));  // hc must outlive the copy
    return ZK_OK;
}
int po_z_poly(zkaes_ctx* ctx, F* out, const F* w, size_t len_w, const F* x_poly, size_t x) {
    LAUNCH(ctx, k_z_poly, len_w + x, TB, out, w, len_w, x_poly, x);
    return ZK_OK;
}
int po_mask_fix(zkaes_ctx* ctx, F* c, size_t n) {
    k_mask_fix<<<1, 1, 0, ctx->stream>>>(c, n);
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    return ZK_OK;
}
int po_bits_to_fr(zkaes_ctx* ctx, F* out, const uint8_t* bits, size_t n) { LAUNCH(ctx, k_bits_to_fr, n, TB, out, bits, n); return ZK_OK; }

}  // namespace zk
