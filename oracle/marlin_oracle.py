"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the Marlin prover the reference's `encrypt()` runs
(`simpleworks::marlin::generate_proof`, src/lib.rs:111) plus the indexer (`synthesize_keys`, src/lib.rs:138-174) and a
verifier (`verify_encryption`, src/lib.rs:116-136).

The algorithm lives in crates that are NOT under /root/reference; this file restates their published
algorithms from the pinned versions:
    ark-marlin 0.3.0 (fork Entropy1729/marlin@bde002d)   ahp/{indexer,prover,verifier,constraint_systems}.rs, lib.rs, rng.rs
    ark-poly-commit 0.3.0                                  kzg10/mod.rs, marlin/marlin_pc/mod.rs, marlin/mod.rs, lib.rs
    ark-poly 0.3.0 / ark-ff 0.3.0 / ark-serialize 0.3.0    domains, DensePolynomial, batch_inversion, wire formats
PARITY PINNING: "parity unpinned" -- the reference holds no golden proof bytes (its tests only assert accept / reject,
tests/integration_tests.rs:330-371) and cannot be compiled here (no cargo).  What this oracle is checked against:
the algebraic verifier below accepts its proofs and rejects tampered statements (tests/test_marlin_oracle.py).  The
product's GPU prover is then compared byte-for-byte with this oracle on the same seeds.

The verifier checks the two KZG batch openings with the TEST-SRS TRAPDOOR instead of a pairing:
    e(C - v*G - rv*gamma*G, H) == e(W, (tau - z) H)   <=>   C - v*G - rv*gamma*G == (tau - z) * W   in G1 (prime order),
which is exact for the insecure test SRS both the reference (README.md:26) and this build use.

Heavy lifting (NTT, MSM, vector ops) is done by oracle/liboracle.so; orchestration and scalars are Python ints.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

from .cpu import FQ, FR, Oracle, ints_to_limbs, limbs_to_ints
from .fs_rng import ChaCha20Rng, FiatShamirRng, fr_rand, fr_rand_many

CURVE = 377
P = FR[CURVE]
Q = FQ[CURVE]
PROTOCOL_NAME = b"MARLIN-2019"
TWO_ADICITY = 47
FR_GEN = 22

_orc = None


def orc() -> Oracle:
    global _orc
    if _orc is None:
        _orc = Oracle()
    return _orc


# ---------------------------------------------------------------------------------------------------------------
# field helpers: python int (canonical) <-> (4,) uint64 Montgomery
# ---------------------------------------------------------------------------------------------------------------
R256 = (1 << 256) % P
R256_INV = pow(R256, -1, P)


def to_m(x: int) -> np.ndarray:
    return ints_to_limbs([x % P * R256 % P], 4)[0]


def from_m(a) -> int:
    return limbs_to_ints(np.asarray(a, dtype=np.uint64).reshape(1, 4))[0] * R256_INV % P


def vec_to_m(xs) -> np.ndarray:
    return ints_to_limbs([x % P * R256 % P for x in xs], 4)


def vec_from_m(a) -> list:
    return [v * R256_INV % P for v in limbs_to_ints(np.asarray(a, dtype=np.uint64).reshape(-1, 4))]


def vadd(a, b):
    return orc().fr_vec(CURVE, 0, a, b)


def vsub(a, b):
    return orc().fr_vec(CURVE, 1, a, b)


def vmul(a, b):
    return orc().fr_vec(CURVE, 2, a, b)


def zeros(n):
    return np.zeros((n, 4), dtype=np.uint64)


def is_zero_rows(a):
    return ~a.any(axis=1)


def poly_trim(c):
    """drop trailing zero coefficients (DensePolynomial::from_coefficients_vec)"""
    nz = np.nonzero(c.any(axis=1))[0]
    return c[: (nz[-1] + 1 if len(nz) else 0)]


def poly_eval(c, x: int) -> int:
    if len(c) == 0:
        return 0
    return from_m(orc().fr_poly_eval(CURVE, c, to_m(x)))


def poly_add_scaled(acc, coeff: int, c):
    """acc += coeff * c  (arrays of possibly different length); returns the new array"""
    if len(c) == 0 or coeff % P == 0:
        return acc
    t = vmul(c, to_m(coeff).reshape(1, 4))
    if len(acc) < len(t):
        acc = np.concatenate([acc, zeros(len(t) - len(acc))])
    acc = acc.copy()
    acc[: len(t)] = vadd(acc[: len(t)], t)
    return acc


class Domain:
    """ark-poly 0.3.0 Radix2EvaluationDomain"""

    def __init__(self, min_size: int):
        n = 1
        while n < min_size:
            n <<= 1
        self.size = n
        self.log = n.bit_length() - 1
        root = pow(FR_GEN, (P - 1) >> TWO_ADICITY, P)
        self.gen = pow(root, 1 << (TWO_ADICITY - self.log), P)
        self._elems = None

    def elements(self) -> np.ndarray:
        if self._elems is None:
            self._elems = orc().fr_powers(CURVE, to_m(self.gen), self.size)
        return self._elems

    def element(self, i: int) -> int:
        return pow(self.gen, i, P)

    def vanishing(self, x: int) -> int:
        return (pow(x, self.size, P) - 1) % P

    def _pad(self, c):
        assert len(c) <= self.size, (len(c), self.size)
        if len(c) == self.size:
            return np.ascontiguousarray(c)
        return np.concatenate([c, zeros(self.size - len(c))])

    def fft(self, c):
        return orc().ntt(CURVE, self._pad(c))

    def ifft(self, e):
        assert len(e) == self.size
        return orc().ntt(CURVE, e, inverse=True)

    def coset_fft(self, c):
        return orc().ntt(CURVE, self._pad(c), coset=True)

    def coset_ifft(self, e):
        return orc().ntt(CURVE, e, inverse=True, coset=True)

    def reindex_by_subdomain(self, other: "Domain", index: int) -> int:
        period = self.size // other.size
        if index < other.size:
            return index * period
        i = index - other.size
        x = period - 1
        return i + (i // x) + 1

    def reindex_vec(self, other: "Domain", idx: np.ndarray) -> np.ndarray:
        period = self.size // other.size
        idx = np.asarray(idx, dtype=np.int64)
        i = idx - other.size
        x = max(period - 1, 1)
        return np.where(idx < other.size, idx * period, i + (i // x) + 1)

    # unnormalized bivariate Lagrange poly u_H(x, y) = (v_H(x) - v_H(y)) / (x - y)
    def u(self, x: int, y: int) -> int:
        if x != y:
            return (self.vanishing(x) - self.vanishing(y)) * pow((x - y) % P, -1, P) % P
        return self.size * pow(x, self.size - 1, P) % P

    def u_alpha_on_domain(self, alpha: int) -> np.ndarray:
        """[u_H(alpha, h) for h in H] = v_H(alpha) / (alpha - h)   (alpha outside H)"""
        d = vsub(np.broadcast_to(to_m(alpha), (self.size, 4)).copy(), self.elements())
        inv = orc().fr_batch_inv(CURVE, d)
        return vmul(inv, to_m(self.vanishing(alpha)).reshape(1, 4))


def divide_by_vanishing(c, n):
    """(q, r) with c = q * (X^n - 1) + r, deg r < n   (DensePolynomial::divide_by_vanishing_poly)"""
    if len(c) <= n:
        return zeros(0), c.copy()
    m = (len(c) + n - 1) // n  # number of blocks
    cp = np.concatenate([c, zeros(m * n - len(c))]).reshape(m, n, 4)
    q = np.zeros((m - 1, n, 4), dtype=np.uint64)
    acc = cp[m - 1]
    q[m - 2] = acc
    for j in range(m - 2, 0, -1):
        acc = vadd(cp[j], acc)
        q[j - 1] = acc
    r = vadd(cp[0], acc)
    return poly_trim(q.reshape(-1, 4)), poly_trim(r)


# ---------------------------------------------------------------------------------------------------------------
# G1 / serialisation helpers
# ---------------------------------------------------------------------------------------------------------------
RQ = (1 << 384) % Q
RQ_INV = pow(RQ, -1, Q)
INF = np.zeros(12, dtype=np.uint64)


def g1_xy(pt):
    """affine (12,) uint64 Montgomery -> (x, y) canonical ints, or None for infinity"""
    if not np.asarray(pt).any():
        return None
    x = limbs_to_ints(np.asarray(pt[:6]).reshape(1, 6))[0] * RQ_INV % Q
    y = limbs_to_ints(np.asarray(pt[6:]).reshape(1, 6))[0] * RQ_INV % Q
    return x, y


def g1_to_bytes_uncompressed(pt) -> bytes:
    """ark-ff ToBytes for GroupAffine: x || y (canonical LE, 48 B each) || infinity flag byte"""
    xy = g1_xy(pt)
    if xy is None:
        return (0).to_bytes(48, "little") + (1).to_bytes(48, "little") + b"\x01"
    return xy[0].to_bytes(48, "little") + xy[1].to_bytes(48, "little") + b"\x00"


def g1_serialize_compressed(pt) -> bytes:
    """ark-serialize 0.3.0 CanonicalSerialize for GroupAffine (SWFlags in the top two bits of the last byte)"""
    xy = g1_xy(pt)
    if xy is None:
        b = bytearray(48)
        b[47] |= 1 << 6
        return bytes(b)
    x, y = xy
    b = bytearray(x.to_bytes(48, "little"))
    if y > (Q - y) % Q:
        b[47] |= 1 << 7
    return bytes(b)


def fr_bytes(x: int) -> bytes:
    return (x % P).to_bytes(32, "little")


def g1_add(a, b):
    return orc().g1_add(CURVE, np.asarray(a, dtype=np.uint64), np.asarray(b, dtype=np.uint64))


def g1_neg(a):
    a = np.asarray(a, dtype=np.uint64).copy()
    if not a.any():
        return a
    a[6:] = orc().field_op(CURVE, 1, 6, a[6:].reshape(1, 6))[0]
    return a


def g1_scale(pt, s: int):
    sc = ints_to_limbs([s % P], 4)
    return orc().g1_mul(CURVE, np.asarray(pt, dtype=np.uint64).reshape(1, 12), sc)[0]


def msm(bases, coeffs_mont):
    """VariableBaseMSM::multi_scalar_mul(bases, coeffs.into_repr()) -> affine"""
    n = len(coeffs_mont)
    if n == 0:
        return INF.copy()
    assert len(bases) >= n, (len(bases), n)
    sc = orc().from_mont(CURVE, 0, coeffs_mont)
    return orc().g1_msm(CURVE, np.ascontiguousarray(bases[:n]), sc)


# ---------------------------------------------------------------------------------------------------------------
# SRS (KZG10::setup / MarlinKZG10::trim): test SRS from seeds, trapdoor kept for the algebraic verifier
# ---------------------------------------------------------------------------------------------------------------
def seed_to_scalar(seed32: bytes) -> int:
    """the product's convention (csrc/srs.cu): the seed as a little-endian integer with the top 4 bits cleared"""
    v = int.from_bytes(seed32, "little") & ((1 << 252) - 1)
    return v % P


@dataclass
class SRS:
    max_degree: int
    tau: int
    gamma: int
    powers_of_g: np.ndarray  # (max_degree + 1, 12)
    powers_of_gamma_g: np.ndarray  # (3, 12): gamma * tau^i * G, i <= hiding_bound + 1

    @staticmethod
    def generate(max_degree: int, tau_seed: bytes, gamma_seed: bytes) -> "SRS":
        tau, gamma = seed_to_scalar(tau_seed), seed_to_scalar(gamma_seed)
        pg = orc().g1_srs(CURVE, ints_to_limbs([tau], 4)[0], max_degree + 1)
        gg = orc().g1_mul_gen(CURVE, ints_to_limbs([gamma * pow(tau, i, P) % P for i in range(3)], 4))
        return SRS(max_degree, tau, gamma, pg, gg)


# ---------------------------------------------------------------------------------------------------------------
# Indexer (ahp/indexer.rs + constraint_systems.rs)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class R1CS:
    """Output of circuit synthesis, before Marlin's padding: rows are lists of (column, integer coefficient)."""
    a: list
    b: list
    c: list
    num_instance: int  # including the constant-one variable
    num_witness: int


def pad_r1cs(r: R1CS, instance_vals=None, witness_vals=None):
    """pad_input_for_indexer_and_prover + make_matrices_square: instance count -> power of two (zeros), then dummy
    constraints 0*0=0 or dummy witnesses (value ONE) until #constraints == #variables.  Column indices of witnesses
    move up by the number of padded instance variables."""
    x = 1
    while x < r.num_instance:
        x <<= 1
    shift = x - r.num_instance

    def fix(rows):
        return [[(c + shift if c >= r.num_instance else c, v) for c, v in row] for row in rows]

    a, b, c = fix(r.a), fix(r.b), fix(r.c)
    ncons, nwit = len(a), r.num_witness
    nvar = x + nwit
    inst = None if instance_vals is None else list(instance_vals) + [0] * shift
    wit = None if witness_vals is None else list(witness_vals)
    if nvar > ncons:
        for _ in range(nvar - ncons):
            a.append([]); b.append([]); c.append([])
    else:
        nwit += ncons - nvar
        if wit is not None:
            wit += [1] * (ncons - nvar)
    return R1CS(a, b, c, x, nwit), inst, wit


def balance_matrices(a, b):
    a_d = sum(len(r) for r in a)
    b_d = sum(len(r) for r in b)
    a_is_denser = a_d == max(a_d, b_d)
    for i in range(len(a)):
        if a_is_denser:
            ra, rb = len(a[i]), len(b[i])
            a[i], b[i] = b[i], a[i]
            a_d = a_d - ra + rb
            b_d = b_d - rb + ra
            a_is_denser = a_d == max(a_d, b_d)


@dataclass
class MatrixArith:
    row_idx: np.ndarray  # H-index of row(kappa) (= the VARIABLE's element: arithmetisation of M^*)
    col_idx: np.ndarray  # H-index of col(kappa) (= the CONSTRAINT's element)
    evals: dict  # 'row','col','val' -> (k,4) evaluations on K
    polys: dict  # 'row','col','val','row_col' -> coefficient arrays (length k)


@dataclass
class Index:
    num_variables: int
    num_constraints: int
    num_non_zero: int
    num_instance: int
    a: list
    b: list
    c: list
    arith: dict  # 'a','b','c' -> MatrixArith
    domain_h: Domain
    domain_k: Domain
    domain_x: Domain
    max_degree: int = 0
    comms: list = field(default_factory=list)  # 12 commitments (affine) in INDEXER_POLYNOMIALS order

    def info_bytes(self) -> bytes:
        return struct.pack("<QQQ", self.num_variables, self.num_constraints, self.num_non_zero)

    def vk_bytes(self) -> bytes:
        """ToBytes of IndexVerifierKey: index_info || index_comms"""
        return self.info_bytes() + b"".join(commitment_to_bytes(c, None) for c in self.comms)


INDEXER_POLYNOMIALS = ["a_row", "a_col", "a_val", "a_row_col", "b_row", "b_col", "b_val", "b_row_col", "c_row", "c_col", "c_val",
                       "c_row_col"]


def arithmetize_matrix(m, dk: Domain, dh: Domain, dx: Domain) -> MatrixArith:
    rows, cols, vals = [], [], []
    for r, row in enumerate(m):
        for c, v in sorted(row):
            rows.append(c)
            cols.append(r)
            vals.append(v)
    count = len(rows)
    row_idx = dh.reindex_vec(dx, np.array(rows, dtype=np.int64)) if count else np.zeros(0, dtype=np.int64)
    col_idx = np.array(cols, dtype=np.int64)
    elems = dh.elements()
    # val = M / u_H(row, row), u_H(y, y) = |H| y^(|H|-1) = |H| / y
    hinv = pow(dh.size, -1, P)
    small = {v: to_m(v * hinv) for v in set(vals)}
    val_c = np.stack([small[v] for v in vals]) if count else zeros(0)
    val_e = vmul(val_c, elems[row_idx]) if count else zeros(0)
    pad = dk.size - count
    row_idx = np.concatenate([row_idx, np.zeros(pad, dtype=np.int64)])
    col_idx = np.concatenate([col_idx, np.zeros(pad, dtype=np.int64)])
    row_e, col_e = elems[row_idx], elems[col_idx]
    val_e = np.concatenate([val_e, zeros(pad)])
    rc_e = vmul(row_e, col_e)
    ev = {"row": row_e, "col": col_e, "val": val_e}
    polys = {"row": dk.ifft(row_e), "col": dk.ifft(col_e), "val": dk.ifft(val_e), "row_col": dk.ifft(rc_e)}
    return MatrixArith(row_idx, col_idx, ev, polys)


def ahp_max_degree(h: int, k: int) -> int:
    zk = 1
    return max(2 * h + zk - 2, 3 * h + 2 * zk - 3, h, h, 3 * k - 3)


def index_r1cs(r1cs: R1CS, srs: SRS | None = None) -> Index:
    """AHPForR1CS::index + Marlin::index (commitments to the 12 index polynomials, no hiding)."""
    pr, _, _ = pad_r1cs(r1cs)
    a, b, c = [list(r) for r in pr.a], [list(r) for r in pr.b], [list(r) for r in pr.c]
    nnz = max(sum(len(r) for r in m) for m in (a, b, c))
    balance_matrices(a, b)
    ncons = len(a)
    nvar = pr.num_instance + pr.num_witness
    assert ncons == nvar, "NonSquareMatrix"
    dh, dk, dx = Domain(ncons), Domain(nnz), Domain(pr.num_instance)
    ar = {"a": arithmetize_matrix(a, dk, dh, dx), "b": arithmetize_matrix(b, dk, dh, dx), "c": arithmetize_matrix(c, dk, dh, dx)}
    idx = Index(nvar, ncons, nnz, pr.num_instance, a, b, c, ar, dh, dk, dx, ahp_max_degree(dh.size, dk.size))
    if srs is not None:
        assert srs.max_degree >= idx.max_degree
        for name in INDEXER_POLYNOMIALS:
            m, which = name[0], name[2:]
            idx.comms.append(msm(srs.powers_of_g, poly_trim(ar[m].polys[which])))
    return idx


# ---------------------------------------------------------------------------------------------------------------
# Polynomial commitments (MarlinKZG10)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class Rand:
    blinding: np.ndarray  # coefficient array (empty = not hiding)
    shifted: np.ndarray | None = None


def commitment_to_bytes(comm, shifted) -> bytes:
    """ToBytes of marlin_pc::Commitment: comm || shifted_exists || (shifted or the identity)"""
    return g1_to_bytes_uncompressed(comm) + (b"\x01" if shifted is not None else b"\x00") + g1_to_bytes_uncompressed(
        shifted if shifted is not None else INF)


def kzg_commit(powers, gamma_powers, coeffs, hiding_bound, rng):
    c = poly_trim(coeffs)
    nz = np.nonzero(c.any(axis=1))[0]
    lead = int(nz[0]) if len(nz) else 0  # skip_leading_zeros_and_convert_to_bigints
    comm = msm(powers[lead:], c[lead:])
    blind = zeros(0)
    if hiding_bound is not None:
        blind = fr_rand_many(rng, P, hiding_bound + 2)  # P::rand(hiding_bound + 1): degree + 1 coefficients
        comm = g1_add(comm, msm(gamma_powers, blind))
    return comm, blind


def pc_commit(srs: SRS, max_degree: int, polys, rng):
    """polys: list of (label, coeffs, degree_bound, hiding_bound) -> list of (comm, shifted_comm), list of Rand"""
    comms, rands = [], []
    for _label, coeffs, bound, hiding in polys:
        comm, blind = kzg_commit(srs.powers_of_g, srs.powers_of_gamma_g, coeffs, hiding, rng)
        shifted, sblind = None, None
        if bound is not None:
            shifted, sblind = kzg_commit(srs.powers_of_g[max_degree - bound:], srs.powers_of_gamma_g, coeffs, hiding, rng)
        comms.append((comm, shifted))
        rands.append(Rand(blind, sblind))
    return comms, rands


# ---------------------------------------------------------------------------------------------------------------
# Prover (ahp/prover.rs + lib.rs::prove)
# ---------------------------------------------------------------------------------------------------------------
def sample_outside_domain(rng, d: Domain) -> int:
    while True:
        t = from_m(fr_rand(rng, P))
        if d.vanishing(t) != 0:
            return t


def matvec_small(m, z):
    """rows of small integer coefficients times an integer assignment -> list of ints"""
    return [sum(v * z[c] for c, v in row) for row in m]


def prove(idx: Index, srs: SRS, r1cs: R1CS, instance_vals, witness_vals, zk_seed: bytes):
    """Returns (proof dict, proof bytes).  instance_vals includes the leading 1."""
    pr, inst, wit = pad_r1cs(r1cs, instance_vals, witness_vals)
    dh, dk, dx = idx.domain_h, idx.domain_k, idx.domain_x
    h, k, x = dh.size, dk.size, dx.size
    D = idx.max_degree
    zk = ChaCha20Rng(zk_seed)
    public_input = inst[1:]
    fs = FiatShamirRng(PROTOCOL_NAME + idx.vk_bytes() + b"".join(fr_bytes(v) for v in public_input))

    # ---------------- first round ----------------
    z = inst + wit
    za_i, zb_i = matvec_small(idx.a, z), matvec_small(idx.b, z)
    za_i += [0] * (h - len(za_i))  # Evaluations::interpolate zero-pads to the domain size
    zb_i += [0] * (h - len(zb_i))
    x_poly = dx.ifft(vec_to_m(inst))
    x_evals = dh.fft(x_poly)
    ratio = h // x
    w_ext = wit + [0] * (h - x - len(wit))
    kk = np.arange(h)
    src = kk - kk // ratio - 1
    w_vals = vec_to_m(w_ext + [0])[np.where(kk % ratio == 0, len(w_ext), src)]
    w_evals = vsub(w_vals, x_evals)
    w_evals[kk % ratio == 0] = 0
    v_h_blind = lambda c, rnd: _add_vanishing_multiple(c, rnd, h)
    r_w, r_a, r_b = fr_rand(zk, P), fr_rand(zk, P), fr_rand(zk, P)
    w_full = v_h_blind(dh.ifft(w_evals), r_w)
    w_poly, rem = divide_by_vanishing(w_full, x)
    assert len(rem) == 0, "w is not divisible by v_X"
    z_a_poly = v_h_blind(dh.ifft(vec_to_m(za_i)), r_a)
    z_b_poly = v_h_blind(dh.ifft(vec_to_m(zb_i)), r_b)
    mask = fr_rand_many(zk, P, 3 * h + 2 - 3 + 1)
    sigma = mask[0]
    for j in range(h, len(mask), h):
        sigma = vadd(sigma.reshape(1, 4), mask[j].reshape(1, 4))[0]
    mask[0] = vsub(mask[0].reshape(1, 4), sigma.reshape(1, 4))[0]
    first = [("w", w_poly, None, 1), ("z_a", z_a_poly, None, 1), ("z_b", z_b_poly, None, 1), ("mask_poly", mask, None, None)]
    c1, r1 = pc_commit(srs, D, first, zk)
    fs.absorb(b"".join(commitment_to_bytes(*c) for c in c1))
    alpha = sample_outside_domain(fs.rng, dh)
    eta_a, eta_b, eta_c = (from_m(fr_rand(fs.rng, P)) for _ in range(3))

    # ---------------- second round ----------------
    r_alpha_evals = dh.u_alpha_on_domain(alpha)
    r_alpha_poly = dh.ifft(r_alpha_evals)
    ra_int = vec_from_m(r_alpha_evals)
    t_evals = [0] * h
    for m, eta in ((idx.a, eta_a), (idx.b, eta_b), (idx.c, eta_c)):
        for r, row in enumerate(m):
            if not row:
                continue
            e = eta * ra_int[r] % P
            for c, v in row:
                j = dh.reindex_by_subdomain(dx, c)
                t_evals[j] = (t_evals[j] + e * v) % P
    t_poly = dh.ifft(vec_to_m(t_evals))
    # z(X) = w(X) v_X(X) + x(X)
    z_poly = np.concatenate([zeros(x), w_poly])
    z_poly[: len(w_poly)] = vsub(z_poly[: len(w_poly)], w_poly)
    z_poly[:x] = vadd(z_poly[:x], x_poly)
    dm = Domain(4 * h)
    ea, eb = dm.fft(z_a_poly), dm.fft(z_b_poly)
    summed = vadd(vmul(vadd(vmul(eb, to_m(eta_c).reshape(1, 4)), np.broadcast_to(to_m(eta_a), (dm.size, 4)).copy()), ea),
                  vmul(eb, to_m(eta_b).reshape(1, 4)))
    rhs_e = vsub(vmul(dm.fft(r_alpha_poly), summed), vmul(dm.fft(t_poly), dm.fft(z_poly)))
    rhs = dm.ifft(rhs_e)
    q1 = mask.copy()
    q1 = np.concatenate([q1, zeros(max(0, len(rhs) - len(q1)))])
    q1[: len(rhs)] = vadd(q1[: len(rhs)], rhs)
    h_1, x_g_1 = divide_by_vanishing(poly_trim(q1), h)
    assert len(x_g_1) == 0 or not x_g_1[0].any(), "outer sumcheck: sum over H is not zero"
    g_1 = poly_trim(x_g_1[1:])
    second = [("t", t_poly, None, None), ("g_1", g_1, h - 2, 1), ("h_1", h_1, None, 1)]
    c2, r2 = pc_commit(srs, D, second, zk)
    fs.absorb(b"".join(commitment_to_bytes(*c) for c in c2))
    beta = sample_outside_domain(fs.rng, dh)

    # ---------------- third round ----------------
    vh_a, vh_b = dh.vanishing(alpha), dh.vanishing(beta)
    am, bm = to_m(alpha).reshape(1, 4), to_m(beta).reshape(1, 4)
    f_e = zeros(k)
    for mname, eta in (("a", eta_a), ("b", eta_b), ("c", eta_c)):
        ev = idx.arith[mname].evals
        den = vmul(vsub(np.broadcast_to(bm, (k, 4)).copy(), ev["row"]), vsub(np.broadcast_to(am, (k, 4)).copy(), ev["col"]))
        f_e = vadd(f_e, vmul(vmul(ev["val"], orc().fr_batch_inv(CURVE, den)), to_m(eta).reshape(1, 4)))
    f_e = vmul(f_e, to_m(vh_a * vh_b).reshape(1, 4))
    f_poly = dk.ifft(f_e)
    g_2 = poly_trim(f_poly[1:])
    # h_2 = (a - b f) / v_K, computed on the coset g*B (|B| = 4k) where v_K does not vanish (same polynomial as the
    # interpolate-on-B-then-divide route of ahp/prover.rs)
    db = Domain(4 * k)
    den_e, val_e = {}, {}
    ab = to_m(alpha * beta)
    for mname in "abc":
        pl = idx.arith[mname].polys
        row, col, rc = db.coset_fft(pl["row"]), db.coset_fft(pl["col"]), db.coset_fft(pl["row_col"])
        den_e[mname] = vadd(vsub(vsub(np.broadcast_to(ab, (db.size, 4)).copy(), vmul(row, am)), vmul(col, bm)), rc)
        val_e[mname] = db.coset_fft(pl["val"])
    a_e = vadd(vadd(vmul(vmul(val_e["a"], vmul(den_e["b"], den_e["c"])), to_m(eta_a).reshape(1, 4)),
                    vmul(vmul(val_e["b"], vmul(den_e["a"], den_e["c"])), to_m(eta_b).reshape(1, 4))),
               vmul(vmul(val_e["c"], vmul(den_e["a"], den_e["b"])), to_m(eta_c).reshape(1, 4)))
    a_e = vmul(a_e, to_m(vh_a * vh_b).reshape(1, 4))
    b_e = vmul(vmul(den_e["a"], den_e["b"]), den_e["c"])
    num = vsub(a_e, vmul(b_e, db.coset_fft(f_poly)))
    # v_K on the coset: (g w^i)^k - 1, period 4 in i
    gk = pow(FR_GEN, k, P)
    w4 = pow(db.gen, k, P)
    vk_inv4 = vec_to_m([pow((gk * pow(w4, i, P) - 1) % P, -1, P) for i in range(4)])
    num = vmul(num, np.tile(vk_inv4, (db.size // 4, 1)))
    h_2 = poly_trim(db.coset_ifft(num))
    assert len(h_2) <= 3 * k - 3, "inner sumcheck quotient has the wrong degree"
    third = [("g_2", g_2, k - 2, None), ("h_2", h_2, None, None)]
    c3, r3 = pc_commit(srs, D, third, zk)
    fs.absorb(b"".join(commitment_to_bytes(*c) for c in c3))
    gamma = from_m(fr_rand(fs.rng, P))

    # ---------------- evaluations + openings ----------------
    polys = {}
    rands = {}
    bounds = {}
    for name in INDEXER_POLYNOMIALS:
        polys[name] = idx.arith[name[0]].polys[name[2:]]
        rands[name] = Rand(zeros(0))
        bounds[name] = None
    for (label, coeffs, bound, _hid), rnd in zip(first + second + third, r1 + r2 + r3):
        polys[label], rands[label], bounds[label] = coeffs, rnd, bound
    ev = {"g_1": poly_eval(g_1, beta), "g_2": poly_eval(g_2, gamma), "t": poly_eval(t_poly, beta), "z_b": poly_eval(z_b_poly, beta)}
    ab = alpha * beta % P
    for m in "abc":
        pl = idx.arith[m].polys
        ev[m + "_denom"] = (ab - alpha * poly_eval(pl["row"], gamma) - beta * poly_eval(pl["col"], gamma) + poly_eval(pl["row_col"], gamma)) % P
    evaluations = [ev[lab] for lab in sorted(ev)]  # a_denom, b_denom, c_denom, g_1, g_2, t, z_b
    fs.absorb(b"".join(fr_bytes(v) for v in evaluations))
    opening_challenge = fs.rng.next_u128() % P
    lcs = construct_lcs(idx, public_input, ev, alpha, eta_a, eta_b, eta_c, beta, gamma)
    pc_proof = []
    for point, labels in query_points(beta, gamma):
        pc_proof.append(open_at_point(srs, D, lcs, polys, rands, bounds, labels, point, opening_challenge))
    proof = {"commitments": [c1, c2, c3], "evaluations": evaluations, "pc_proof": pc_proof}
    return proof, serialize_proof(proof)


def query_points(beta, gamma):
    """verifier_query_set grouped by point label ("beta" < "gamma"), LC labels in BTreeSet order"""
    return ((beta, ["g_1", "outer_sumcheck", "t", "z_b"]), (gamma, ["a_denom", "b_denom", "c_denom", "g_2", "inner_sumcheck"]))


EVAL_LABELS = ["a_denom", "b_denom", "c_denom", "g_1", "g_2", "t", "z_b"]


def _add_vanishing_multiple(c, rnd, n):
    """c + rnd * (X^n - 1)"""
    out = np.concatenate([c, zeros(n + 1 - len(c))])
    out[0] = vsub(out[0].reshape(1, 4), rnd.reshape(1, 4))[0]
    out[n] = vadd(out[n].reshape(1, 4), rnd.reshape(1, 4))[0]
    return out


def construct_lcs(idx: Index, public_input, ev, alpha, eta_a, eta_b, eta_c, beta, gamma):
    """ahp/mod.rs construct_linear_combinations.  An LC is a list of (coeff, label | None for the constant one).
    `ev` holds the seven evaluations the proof carries (EVAL_LABELS)."""
    dh, dk, dx = idx.domain_h, idx.domain_k, idx.domain_x
    r_alpha_at_beta = dh.u(alpha, beta)
    v_h_alpha, v_h_beta, v_x_beta = dh.vanishing(alpha), dh.vanishing(beta), dx.vanishing(beta)
    z_b_b, t_b, g_1_b, g_2_g = ev["z_b"], ev["t"], ev["g_1"], ev["g_2"]
    x_poly = dx.ifft(vec_to_m([1] + list(public_input)))
    x_at_beta = poly_eval(x_poly, beta)
    lcs = {
        "z_b": [(1, "z_b")], "g_1": [(1, "g_1")], "t": [(1, "t")], "g_2": [(1, "g_2")],
        "outer_sumcheck": [(1, "mask_poly"), (r_alpha_at_beta * (eta_a + eta_c * z_b_b) % P, "z_a"), (r_alpha_at_beta * eta_b * z_b_b % P, None),
                           (-t_b * v_x_beta % P, "w"), (-t_b * x_at_beta % P, None), (-v_h_beta % P, "h_1"), (-beta * g_1_b % P, None)],
    }
    ab = alpha * beta % P
    for m in "abc":
        lcs[m + "_denom"] = [(ab, None), (-alpha % P, m + "_row"), (-beta % P, m + "_col"), (1, m + "_row_col")]
    dn = {m: ev[m + "_denom"] for m in "abc"}
    vv = v_h_alpha * v_h_beta % P
    b_at_gamma = dn["a"] * dn["b"] * dn["c"] % P
    b_expr = b_at_gamma * (gamma * g_2_g + t_b * pow(dk.size, -1, P)) % P
    lcs["inner_sumcheck"] = [(eta_a * dn["b"] * dn["c"] * vv % P, "a_val"), (eta_b * dn["a"] * dn["c"] * vv % P, "b_val"),
                             (eta_c * dn["b"] * dn["a"] * vv % P, "c_val"), (-b_expr % P, None), (-dk.vanishing(gamma) % P, "h_2")]
    return lcs


def open_at_point(srs, D, lcs, polys, rands, bounds, labels, point, ch):
    """marlin_pc open_individual_opening_challenges over the LC polynomials queried at `point` (labels sorted)."""
    p = zeros(0)
    r = zeros(0)
    shifted_terms = []  # (challenge, witness coeffs, bound, shifted blinding)
    j = 0
    for lab in labels:
        terms = lcs[lab]
        poly, rnd = zeros(0), zeros(0)
        for coeff, name in terms:
            if name is None:
                continue
            poly = poly_add_scaled(poly, coeff, polys[name])
            rnd = poly_add_scaled(rnd, coeff, rands[name].blinding)
        named = [t for t in terms if t[1] is not None]
        bound = bounds[named[0][1]] if len(named) == 1 and len(terms) == 1 else None
        cj = pow(ch, j, P)
        j += 1
        p = poly_add_scaled(p, cj, poly)
        r = poly_add_scaled(r, cj, rnd)
        if bound is not None:
            cj1 = pow(ch, j, P)
            j += 1
            name = terms[0][1]
            sb = rands[name].shifted
            shifted_terms.append((cj1, orc().fr_div_linear(CURVE, polys[name], to_m(point)), bound, sb))
    zm = to_m(point)
    w = msm(srs.powers_of_g, poly_trim(orc().fr_div_linear(CURVE, p, zm))) if len(p) > 1 else INF.copy()
    random_v = None
    if len(poly_trim(r)):
        w = g1_add(w, msm(srs.powers_of_gamma_g, orc().fr_div_linear(CURVE, r, zm)))
        random_v = poly_eval(r, point)
    for cj1, wit, bound, sb in shifted_terms:
        # shift_polynomial by (largest_bound - bound) on shifted_powers(None) = powers[D - largest_bound..]
        sw = msm(srs.powers_of_g[D - bound:], poly_trim(wit))
        w = g1_add(w, g1_scale(sw, cj1))
        if sb is not None and len(poly_trim(sb)):
            w = g1_add(w, g1_scale(msm(srs.powers_of_gamma_g, orc().fr_div_linear(CURVE, sb, zm)), cj1))
            random_v = ((random_v or 0) + cj1 * poly_eval(sb, point)) % P
    return {"w": w, "random_v": random_v}


def serialize_proof(proof) -> bytes:
    """ark-serialize 0.3.0 CanonicalSerialize of ark_marlin::Proof (compressed points)."""
    out = bytearray()
    out += struct.pack("<Q", len(proof["commitments"]))
    for rnd in proof["commitments"]:
        out += struct.pack("<Q", len(rnd))
        for comm, shifted in rnd:
            out += g1_serialize_compressed(comm)
            if shifted is None:
                out += b"\x00"
            else:
                out += b"\x01" + g1_serialize_compressed(shifted)
    out += struct.pack("<Q", len(proof["evaluations"]))
    for v in proof["evaluations"]:
        out += fr_bytes(v)
    out += struct.pack("<Q", 3) + b"\x00\x00\x00"  # three ProverMsg::EmptyMessage -> Option::None
    out += struct.pack("<Q", len(proof["pc_proof"]))
    for pr in proof["pc_proof"]:
        out += g1_serialize_compressed(pr["w"])
        out += b"\x00" if pr["random_v"] is None else b"\x01" + fr_bytes(pr["random_v"])
    out += b"\x00"  # BatchLCProof.evals = None
    return bytes(out)


# ---------------------------------------------------------------------------------------------------------------
# Verifier (lib.rs::verify + ahp/verifier.rs + marlin_pc check_combinations), KZG check through the test trapdoor
# ---------------------------------------------------------------------------------------------------------------
def verify(idx: Index, srs: SRS, public_input, proof) -> bool:
    """public_input: the statement WITHOUT the leading one (the reference passes 8 bits per ciphertext byte,
    src/lib.rs:121-128); padded here with zeros to |X| - 1 like lib.rs::verify."""
    dh, dk, dx = idx.domain_h, idx.domain_k, idx.domain_x
    D = idx.max_degree
    public_input = list(public_input) + [0] * (dx.size - 1 - len(public_input))
    if len(public_input) != dx.size - 1:
        return False
    c1, c2, c3 = proof["commitments"]
    fs = FiatShamirRng(PROTOCOL_NAME + idx.vk_bytes() + b"".join(fr_bytes(v) for v in public_input))
    fs.absorb(b"".join(commitment_to_bytes(*c) for c in c1))
    alpha = sample_outside_domain(fs.rng, dh)
    eta_a, eta_b, eta_c = (from_m(fr_rand(fs.rng, P)) for _ in range(3))
    fs.absorb(b"".join(commitment_to_bytes(*c) for c in c2))
    beta = sample_outside_domain(fs.rng, dh)
    fs.absorb(b"".join(commitment_to_bytes(*c) for c in c3))
    gamma = from_m(fr_rand(fs.rng, P))
    if len(proof["evaluations"]) != len(EVAL_LABELS):
        return False
    ev = dict(zip(EVAL_LABELS, proof["evaluations"]))
    fs.absorb(b"".join(fr_bytes(v) for v in proof["evaluations"]))
    ch = fs.rng.next_u128() % P
    lcs = construct_lcs(idx, public_input, ev, alpha, eta_a, eta_b, eta_c, beta, gamma)
    comms = {}
    for name, c in zip(INDEXER_POLYNOMIALS, idx.comms):
        comms[name] = (c, None, None)
    h, k = dh.size, dk.size
    for (label, bound), c in zip([("w", None), ("z_a", None), ("z_b", None), ("mask_poly", None), ("t", None), ("g_1", h - 2),
                                  ("h_1", None), ("g_2", k - 2), ("h_2", None)], c1 + c2 + c3):
        if (bound is None) != (c[1] is None):
            return False
        comms[label] = (c[0], c[1], bound)
    G, gamma_G = srs.powers_of_g[0], srs.powers_of_gamma_g[0]
    for (point, labels), pr in zip(query_points(beta, gamma), proof["pc_proof"]):
        comb = INF.copy()
        comb_v = 0
        j = 0
        for lab in labels:
            terms = lcs[lab]
            value = ev.get(lab, 0)  # the two sumcheck LCs evaluate to zero
            c_lc, shifted, bound = INF.copy(), None, None
            for coeff, name in terms:
                if name is None:
                    value = (value - coeff) % P
                    continue
                cm, sh, bd = comms[name]
                if bd is not None:
                    if len(terms) != 1 or coeff != 1:
                        return False
                    shifted, bound = sh, bd
                c_lc = g1_add(c_lc, g1_scale(cm, coeff))
            cj = pow(ch, j, P)
            j += 1
            comb = g1_add(comb, g1_scale(c_lc, cj))
            comb_v = (comb_v + cj * value) % P
            if bound is not None:
                cj1 = pow(ch, j, P)
                j += 1
                adj = g1_add(shifted, g1_neg(g1_scale(srs.powers_of_g[D - bound], value)))
                comb = g1_add(comb, g1_scale(adj, cj1))
        lhs = g1_add(comb, g1_neg(g1_scale(G, comb_v)))
        if pr["random_v"] is not None:
            lhs = g1_add(lhs, g1_neg(g1_scale(gamma_G, pr["random_v"])))
        rhs = g1_scale(pr["w"], (srs.tau - point) % P)
        if not (lhs == rhs).all():
            return False
    return True


# ---------------------------------------------------------------------------------------------------------------
# Wire formats read back (ark-serialize 0.3.0 CanonicalDeserialize) + a verifier set-up that needs no indexer run
# ---------------------------------------------------------------------------------------------------------------
def _fq_sqrt(a: int):
    """Tonelli-Shanks in Fq (q = 1 mod 2^46)"""
    a %= Q
    if a == 0:
        return 0
    if pow(a, (Q - 1) // 2, Q) != 1:
        return None
    s, t = 0, Q - 1
    while t % 2 == 0:
        s += 1
        t //= 2
    z = 2
    while pow(z, (Q - 1) // 2, Q) != Q - 1:
        z += 1
    m, c, tt, r = s, pow(z, t, Q), pow(a, t, Q), pow(a, (t + 1) // 2, Q)
    while tt != 1:
        i, t2 = 0, tt
        while t2 != 1:
            t2 = t2 * t2 % Q
            i += 1
        b = pow(c, 1 << (m - i - 1), Q)
        m, c = i, b * b % Q
        tt, r = tt * c % Q, r * b % Q
    return r


def _xy_to_affine(x: int, y: int):
    return np.concatenate([ints_to_limbs([x * RQ % Q], 6)[0], ints_to_limbs([y * RQ % Q], 6)[0]])


def g1_deserialize_compressed(b: bytes):
    assert len(b) == 48
    flags = b[47] >> 6
    if flags & 1:
        return INF.copy()
    x = int.from_bytes(b[:47] + bytes([b[47] & 0x3F]), "little")
    y = _fq_sqrt((x * x * x + 1) % Q)  # BLS12-377 G1: y^2 = x^3 + 1
    if y is None:
        raise ValueError("not on the curve")
    if (y > (Q - y) % Q) != bool(flags & 2):
        y = (Q - y) % Q
    return _xy_to_affine(x, y)


def g1_from_bytes_uncompressed(b: bytes):
    assert len(b) == 97
    if b[96]:
        return INF.copy()
    return _xy_to_affine(int.from_bytes(b[:48], "little"), int.from_bytes(b[48:96], "little"))


def deserialize_proof(data: bytes):
    pos = 0

    def u64():
        nonlocal pos
        v = struct.unpack_from("<Q", data, pos)[0]
        pos += 8
        return v

    def take(n):
        nonlocal pos
        v = data[pos:pos + n]
        pos += n
        return v

    rounds = []
    for _ in range(u64()):
        rnd = []
        for _ in range(u64()):
            comm = g1_deserialize_compressed(take(48))
            shifted = g1_deserialize_compressed(take(48)) if take(1)[0] else None
            rnd.append((comm, shifted))
        rounds.append(rnd)
    evaluations = [int.from_bytes(take(32), "little") for _ in range(u64())]
    for _ in range(u64()):  # prover messages
        if take(1)[0]:
            for _ in range(u64()):
                take(32)
    pc = []
    for _ in range(u64()):
        w = g1_deserialize_compressed(take(48))
        rv = int.from_bytes(take(32), "little") if take(1)[0] else None
        pc.append({"w": w, "random_v": rv})
    if take(1)[0]:
        raise ValueError("unexpected BatchLCProof.evals")
    assert pos == len(data), "trailing bytes"
    return {"commitments": rounds, "evaluations": evaluations, "pc_proof": pc}


def index_from_vk_bytes(vk: bytes, num_instance_padded: int) -> Index:
    """Verifier-side index: sizes + the 12 index commitments, read from IndexVerifierKey's ToBytes form."""
    nvar, ncons, nnz = struct.unpack_from("<QQQ", vk, 0)
    comms = [g1_from_bytes_uncompressed(vk[24 + 195 * i: 24 + 195 * i + 97]) for i in range(12)]
    dh, dk, dx = Domain(ncons), Domain(nnz), Domain(num_instance_padded)
    idx = Index(nvar, ncons, nnz, num_instance_padded, [], [], [], {}, dh, dk, dx, ahp_max_degree(dh.size, dk.size), comms)
    idx._vk_raw = vk
    idx.vk_bytes = lambda: vk
    return idx


def g2_serialize_compressed(pt) -> bytes:
    """ark-serialize 0.3.0 CanonicalSerialize for a G2 GroupAffine over Fq2: x.c0 || x.c1 canonical LE (96 B), SWFlags in the top two
    bits of the last byte; "positive" means y > -y in ark-ff's order of quadratic extensions (c1 compared first, then c0)."""
    if pt is None:
        b = bytearray(96)
        b[95] |= 1 << 6
        return bytes(b)
    (x0, x1), (y0, y1) = pt
    b = bytearray(x0.to_bytes(48, "little") + x1.to_bytes(48, "little"))
    n0, n1 = (Q - y0) % Q, (Q - y1) % Q
    if (y1, y0) > (n1, n0):
        b[95] |= 1 << 7
    return bytes(b)


def verifying_key_bytes(index_vk: bytes, x_padded: int, max_degree: int, h: int, k: int, tau_seed: bytes, gamma_seed: bytes) -> bytes:
    """The VerifyingKey `verify_encryption` takes: ark-serialize 0.3.0 CanonicalSerialize of ark_marlin::IndexVerifierKey --
    index_info {num_variables, num_constraints, num_non_zero, num_instance_variables} (4 x u64), Vec of 12 marlin_pc::Commitment
    {comm: compressed G1, shifted_comm: None}, marlin_pc::VerifierKey {kzg10 vk {g, gamma_g: G1; h, beta_h = tau h: compressed G2},
    degree_bounds_and_shift_powers: Some([(bound, tau^(D - bound) G)]) ascending, max_degree, supported_degree}.
    `index_vk` is IndexVerifierKey's ToBytes form (Index.vk_bytes()).  Built independently of the product (oracle G1 arithmetic,
    big-integer G2 from tools/pairing_model.py)."""
    from tools import pairing_model as pr

    tau, gamma = seed_to_scalar(tau_seed), seed_to_scalar(gamma_seed)
    g1 = lambda s: g1_serialize_compressed(orc().g1_mul_gen(CURVE, ints_to_limbs([s % P], 4))[0])
    nvar, ncons, nnz = struct.unpack_from("<QQQ", index_vk, 0)
    out = struct.pack("<QQQQ", nvar, ncons, nnz, x_padded) + struct.pack("<Q", 12)
    for i in range(12):
        out += g1_serialize_compressed(g1_from_bytes_uncompressed(index_vk[24 + 195 * i: 24 + 195 * i + 97])) + b"\x00"
    out += g1(1) + g1(gamma) + g2_serialize_compressed(pr.G2) + g2_serialize_compressed(pr.e2mul(pr.G2, tau))
    bounds = sorted({h - 2, k - 2})
    out += b"\x01" + struct.pack("<Q", len(bounds))
    for b in bounds:
        out += struct.pack("<Q", b) + g1(pow(tau, max_degree - b, P))
    out += struct.pack("<QQ", max_degree, max_degree)
    return out


class SparseSRS:
    """The handful of SRS elements the verifier touches, derived from the trapdoor (test SRS)."""

    def __init__(self, max_degree: int, tau_seed: bytes, gamma_seed: bytes):
        self.max_degree = max_degree
        self.tau, self.gamma = seed_to_scalar(tau_seed), seed_to_scalar(gamma_seed)
        self.powers_of_gamma_g = orc().g1_mul_gen(CURVE, ints_to_limbs([self.gamma], 4))
        self._cache = {}
        self.powers_of_g = self

    def __getitem__(self, i: int):
        if i not in self._cache:
            self._cache[i] = orc().g1_mul_gen(CURVE, ints_to_limbs([pow(self.tau, i, P)], 4))[0]
        return self._cache[i]
