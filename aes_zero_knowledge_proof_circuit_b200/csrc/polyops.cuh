// Fr vector kernels of the Marlin prover rounds (polyops.cu): the HBM-bound glue around the NTTs and MSMs.
//
// They stand in for the dense-polynomial / evaluation-vector arithmetic ark-marlin 0.3.0's prover performs through
// ark-poly 0.3.0 (DensePolynomial, Evaluations) and ark-ff's batch_inversion -- reached from src/lib.rs:111.
// All vectors are Montgomery Fr elements (32 B) resident in HBM; every kernel is one coalesced pass.
#pragma once
#include "common.cuh"
#include "ff.cuh"

namespace zk {

using FrS = Fp<Fr377Params>;  // the reference's scalar field (src/lib.rs:47)

// out[i] = c * base^i
int po_powers(zkaes_ctx* ctx, FrS* out, size_t n, const FrS& base, const FrS& c);
int po_fill(zkaes_ctx* ctx, FrS* out, size_t n, const FrS& v);
// elementwise: op 0 add, 1 sub, 2 mul  (out may alias a or b)
int po_vec(zkaes_ctx* ctx, int op, FrS* out, const FrS* a, const FrS* b, size_t n);
int po_scale(zkaes_ctx* ctx, FrS* out, const FrS* a, const FrS& s, size_t n);
// acc[i] += s * x[i]
int po_axpy(zkaes_ctx* ctx, FrS* acc, const FrS* x, const FrS& s, size_t n);
// out[i] = s - a[i]
int po_rsub_scalar(zkaes_ctx* ctx, FrS* out, const FrS* a, const FrS& s, size_t n);
// ark-ff batch_inversion (zeros stay zero); out must not alias in
int po_batch_inverse(zkaes_ctx* ctx, FrS* out, const FrS* in, size_t n);
// out[i] = table[idx[i]]
int po_gather(zkaes_ctx* ctx, FrS* out, const FrS* table, const uint32_t* idx, size_t n);
// out[i] = coeff[i] * table[idx[i]] * s      (small integer coefficients)
int po_gather_scaled(zkaes_ctx* ctx, FrS* out, const FrS* table, const uint32_t* idx, const int8_t* coeff, const FrS& s, size_t n);
// acc[i] += s * coeff[i] * table[idx[i]] * w[i]
int po_gather_fma(zkaes_ctx* ctx, FrS* acc, const FrS* table, const uint32_t* idx, const int8_t* coeff, const FrS* w, const FrS& s, size_t n);
// out[i] = (beta - table[ridx[i]]) * (alpha - table[cidx[i]])
int po_den_k(zkaes_ctx* ctx, FrS* out, const FrS* table, const uint32_t* ridx, const uint32_t* cidx, const FrS& alpha, const FrS& beta, size_t n);
// sparse rows (CSR, int8 coefficients) times the bit assignment z: out[r] = sum coeff * z[col]  for r < nrows, 0 for nrows <= r < n_out
int po_spmv_bits(zkaes_ctx* ctx, FrS* out, const uint32_t* row_ptr, const uint32_t* col, const int8_t* coeff, const uint8_t* z, size_t nrows,
                 size_t n_out);
// the same row sums as int32 (no field conversion): the small-scalar MSM's input
int po_spmv_bits_i32(zkaes_ctx* ctx, int32_t* out, const uint32_t* row_ptr, const uint32_t* col, const int8_t* coeff, const uint8_t* z, size_t nrows,
                     size_t n_out);
// the full variable assignment laid out over H as ark-marlin orders it (instance variable j at j * ratio, witness in between, zero padded), as int32
int po_assignment_h_i32(zkaes_ctx* ctx, int32_t* out, const uint8_t* z, size_t h, size_t ratio, size_t num_instance, size_t num_witness);
// out[i * stride] = a[i * stride] * s, i < count
int po_scale_strided(zkaes_ctx* ctx, FrS* out, const FrS* a, const FrS& s, size_t stride, size_t count);
// w_evals over H (ahp/prover.rs first round): 0 on the X-subgroup positions, else w_ext[k - k/ratio - 1] - x_evals[k]
int po_w_evals(zkaes_ctx* ctx, FrS* out, const uint8_t* z, const FrS* x_evals, size_t h, size_t ratio, size_t num_instance, size_t num_witness);
// c[0] -= r ; c[n] += r      (c + r * (X^n - 1); c must have n + 1 entries)
int po_add_vanishing(zkaes_ctx* ctx, FrS* c, size_t n, const FrS& r);
// c (len coefficients) = q * (X^n - 1) + rem:  q gets len - n entries (if len > n), rem gets n entries
int po_divide_vanishing(zkaes_ctx* ctx, const FrS* c, size_t len, size_t n, FrS* q, FrS* rem);
// t_evals over H (ahp/prover.rs calculate_t), gather form over the column-major copies of A, B, C:
//   out[reindex(j)] = sum_M eta_M * sum_{(r, v) in column j of M} v * r_alpha[r]
struct CscView {
    const uint32_t* ptr;
    const uint32_t* row;
    const int8_t* coeff;
};
// Columns with many entries (heavy_flag[j] != 0) get one CTA each (heavy_cols) instead of one thread; the few with more than 2^16 entries
// (giant_cols; giant_max_len = a bound of their length in any one matrix) are cut into chunks, one CTA per (column, chunk).
int po_t_evals(zkaes_ctx* ctx, FrS* out, const CscView m[3], const FrS eta[3], const FrS* r_alpha, const uint8_t* heavy_flag,
               const uint32_t* heavy_cols, size_t n_heavy, const uint32_t* giant_cols, size_t n_giant, size_t giant_max_len, size_t nvar, size_t h,
               size_t x);
// out = ra * (eta_a * za + eta_b * zb + eta_c * za * zb) - t * z
int po_round2(zkaes_ctx* ctx, FrS* out, const FrS* ra, const FrS* za, const FrS* zb, const FrS* t, const FrS* z, const FrS eta[3], size_t n);
// den = ab - alpha * row - beta * col + rc   (in place into row)
int po_den_coset(zkaes_ctx* ctx, FrS* row, const FrS* col, const FrS* rc, const FrS& alpha, const FrS& beta, const FrS& ab, size_t n);
// out = (vv * (eta_a val_a den_b den_c + eta_b val_b den_a den_c + eta_c val_c den_a den_b) - den_a den_b den_c * f) * vkinv[i & 3]
int po_round3(zkaes_ctx* ctx, FrS* out, const FrS* const val[3], const FrS* const den[3], const FrS* f, const FrS eta[3], const FrS& vv,
              const FrS vkinv[4], size_t n);
// out[a] = in[a] * base^a   (coset shift of a coefficient vector; out may alias in)
int po_scale_powers(zkaes_ctx* ctx, FrS* out, const FrS* in, const FrS& base, size_t n);
// out = sa * a + sb * b + c, and out[0] += k0   (one denominator polynomial ab - alpha row - beta col + row_col)
int po_lincomb_den(zkaes_ctx* ctx, FrS* out, const FrS* a, const FrS* b, const FrS* c, const FrS& sa, const FrS& sb, const FrS& k0, size_t n);
// out = a * b * c * s
int po_mul3(zkaes_ctx* ctx, FrS* out, const FrS* a, const FrS* b, const FrS* c, const FrS& s, size_t n);
// acc += s * a * b * c
int po_fma3(zkaes_ctx* ctx, FrS* acc, const FrS* a, const FrS* b, const FrS* c, const FrS& s, size_t n);
// v = four k-blocks: block j holds the coefficients of p mod (X^k - i^j) for a polynomial p of degree < 4k, i = i4 a primitive
// 4th root of unity (i.e. its interpolant on the coset w_4k^j * <w_k>, shift undone); replaces them by p's four coefficient blocks
int po_coset4_combine(zkaes_ctx* ctx, FrS* v, size_t k, const FrS& i4_inv);
// v = three |K|-blocks holding the interpolants of a polynomial of degree < 3|K| on the cosets with u_j = s_j^|K|
// (each with its shift undone); replaces them by the polynomial's three coefficient blocks
int po_coset3_combine(zkaes_ctx* ctx, FrS* v, size_t k, const FrS u[3]);
// polynomial evaluation (Horner), result on the host
int po_eval(zkaes_ctx* ctx, const FrS* coeffs, size_t n, const FrS& x, FrS* out_host);
// q = c / (X - z) (synthetic division, remainder dropped): n coefficients in, n - 1 out; q must not alias c
int po_div_linear(zkaes_ctx* ctx, const FrS* c, size_t n, const FrS& z, FrS* q);
// z(X) = w(X) * (X^x - 1) + x_poly(X): out has len_w + x entries
int po_z_poly(zkaes_ctx* ctx, FrS* out, const FrS* w, size_t len_w, const FrS* x_poly, size_t x);
// c[0] = -(c[n] + c[2n])   (mask polynomial: make the sum over H vanish)
int po_mask_fix(zkaes_ctx* ctx, FrS* c, size_t n);
// bits (one byte each) -> Montgomery 0 / 1
int po_bits_to_fr(zkaes_ctx* ctx, FrS* out, const uint8_t* bits, size_t n);

}  // namespace zk
