"""GPU parity tests (run with -m gpu on a B200): every call goes through the C ABI (ctypes) and is compared
bit-exactly with the CPU oracle on the same seeded inputs; large sizes use size-independent properties."""
import numpy as np
import pytest

import aes_zero_knowledge_proof_circuit_b200 as zk
from oracle.cpu import FQ, FR, ints_to_limbs, limbs_to_ints, rand_fr

pytestmark = pytest.mark.gpu
CURVES = [377, 381]
TUNING_DEFAULTS = {"msm_acc_blocks": 3, "msm_window_max": 23, "msm_pair_round": 0, "msm_madd_call": 1, "msm_prefetch": 0}  # csrc/common.cuh zkaes_ctx


def rand_fq(rng, curve, n):
    q = FQ[curve]
    a = rng.integers(0, 1 << 63, size=(n, 6), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 6), dtype=np.uint64)
    a[:, 5] %= np.uint64(q >> 320)
    return a


# ---------------- arithmetic ----------------
@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("variant", [0, 1])
def test_device_field_ops(ctx, oracle, curve, field, variant):
    rng = np.random.default_rng(curve * 10 + field)
    p = (FR if field == 0 else FQ)[curve]
    nl = 4 if field == 0 else 6
    n = 1 << 14
    a = rand_fr(rng, curve, n) if field == 0 else rand_fq(rng, curve, n)
    b = rand_fr(rng, curve, n) if field == 0 else rand_fq(rng, curve, n)
    edge = ints_to_limbs([0, 1, p - 1, p - 2, (1 << (64 * nl - 8)) % p, p >> 1], nl)
    a[: len(edge)] = edge
    b[: len(edge)] = edge[::-1]
    a[len(edge): 2 * len(edge)] = edge
    b[len(edge): 2 * len(edge)] = edge
    for op in (0, 1, 2, 3):  # 3 = square of a (Fq: the dedicated squaring of tools/gen_mont_asm.py) against the oracle's product a * a
        got = ctx.selftest_field(curve, field, op, variant, a, b)
        exp = oracle.field_op(curve, field, 2 if op == 3 else op, a, a if op == 3 else b)
        bad = np.nonzero((got != exp).any(axis=1))[0]
        assert bad.size == 0, f"curve {curve} field {field} op {op} variant {variant}: {bad.size} mismatches, first at {bad[:4]}"


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("field", [0, 1])
def test_device_out_of_line_multiplier(ctx, oracle, curve, field):
    """Fp::mul_call -- the one out-of-line copy of the Montgomery multiplier that the MSM inner loop and the cold curve formulas
    call -- gives the oracle's products (selftest variant 2)"""
    rng = np.random.default_rng(curve + 29 + field)
    p = (FR if field == 0 else FQ)[curve]
    nl = 4 if field == 0 else 6
    n = 1 << 14
    a, b = (rand_fr if field == 0 else rand_fq)(rng, curve, n), (rand_fr if field == 0 else rand_fq)(rng, curve, n)
    edge = ints_to_limbs([0, 1, p - 1, p - 2, (1 << (64 * nl - 8)) % p, p >> 1], nl)
    a[: len(edge)] = edge
    b[: len(edge)] = edge[::-1]
    a[len(edge): 2 * len(edge)] = edge
    b[len(edge): 2 * len(edge)] = edge
    for op in (0, 1, 2, 3):
        got = ctx.selftest_field(curve, field, op, 2, a, b)
        exp = oracle.field_op(curve, field, 2 if op == 3 else op, a, a if op == 3 else b)
        bad = np.nonzero((got != exp).any(axis=1))[0]
        assert bad.size == 0, f"curve {curve} field {field} op {op}: {bad.size} mismatches, first at {bad[:4]}"


@pytest.mark.parametrize("curve", CURVES)
def test_device_g1_formulas(ctx, oracle, curve):
    rng = np.random.default_rng(curve)
    n = 64
    ka, kb = rand_fr(rng, curve, n), rand_fr(rng, curve, n)
    kb[0] = ka[0]
    kb[1] = ints_to_limbs([FR[curve] - limbs_to_ints(ka[1:2])[0]], 4)[0]
    A, B = oracle.g1_mul_gen(curve, ka), oracle.g1_mul_gen(curve, kb)
    B[2] = 0
    A[3] = 0
    exp = np.stack([oracle.g1_add(curve, A[i], B[i]) for i in range(n)])
    for op in (0, 1):
        assert (ctx.selftest_g1(curve, op, A, B) == exp).all(), (curve, op)
    exp2 = np.stack([oracle.g1_add(curve, A[i], A[i]) for i in range(n)])
    assert (ctx.selftest_g1(curve, 2, A, A) == exp2).all()


# ---------------- NTT ----------------
@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 7, 10, 11, 12, 13, 16, 17])
def test_ntt_matches_oracle(ctx, oracle, curve, log_n):
    rng = np.random.default_rng(curve * 100 + log_n)
    x = rand_fr(rng, curve, 1 << log_n)
    for inverse in (False, True):
        for coset in (False, True):
            got = ctx.ntt_fr(curve, x, inverse=inverse, coset=coset)
            exp = oracle.ntt(curve, x, inverse=inverse, coset=coset)
            assert (got == exp).all(), (curve, log_n, inverse, coset)


@pytest.mark.parametrize("curve", CURVES)
def test_ntt_small_matches_dft_definition(ctx, oracle, curve):
    rng = np.random.default_rng(9)
    x = rand_fr(rng, curve, 64)
    assert (ctx.ntt_fr(curve, x) == oracle.dft_naive(curve, x)).all()
    assert (ctx.ntt_fr(curve, x, coset=True) == oracle.dft_naive(curve, x, coset=True)).all()


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_large_roundtrip_and_linearity(ctx, log_n):
    curve = 377
    rng = np.random.default_rng(log_n)
    n = 1 << log_n
    x = rand_fr(rng, curve, n)
    y = ctx.ntt_fr(curve, x)
    assert (ctx.ntt_fr(curve, y, inverse=True) == x).all()
    yc = ctx.ntt_fr(curve, x, coset=True)
    assert (ctx.ntt_fr(curve, yc, inverse=True, coset=True) == x).all()
    # delta at index 1 -> evaluations are the powers of w: y[k] = w^k ; check y[k]^n == 1 spot-wise via another transform
    d = np.zeros((n, 4), dtype=np.uint64)
    d[0] = x[0]
    c = ctx.ntt_fr(curve, d)  # constant polynomial -> all evaluations equal
    assert (c == x[0]).all()


# ---------------- MSM ----------------
@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("n", [1, 2, 33, 1000, 4096])
def test_msm_matches_oracle(ctx, oracle, curve, n):
    rng = np.random.default_rng(curve * 7 + n)
    bases = oracle.g1_walk(curve, 12345 + n, 7, n)
    sc = rand_fr(rng, curve, n)
    if n > 8:
        sc[1] = 0
        sc[2] = ints_to_limbs([1], 4)[0]
        sc[3] = ints_to_limbs([FR[curve] - 1], 4)[0]
        sc[4] = sc[5]          # equal scalars, distinct points
        bases[6] = bases[7]    # equal points (doubling inside a bucket when digits collide)
        sc[6] = sc[7]
        bases[8] = 0           # a point at infinity
    got = ctx.msm_g1(curve, bases, sc)
    exp = oracle.g1_msm(curve, bases, sc, algo=0)
    assert (got == exp).all(), (curve, n)


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("n,bits", [(0, 1), (1, 1), (5, 2), (1000, 1), (1000, 2), (4097, 5), (70001, 13), (200000, 1), (200000, 2)])
def test_small_scalar_msm_matches_oracle(ctx, oracle, curve, n, bits):
    """zkaes_msm_g1_small (one signed digit per term, 32 pseudo-windows whose sums are added without doublings: the kernel path of the
    Lagrange-basis commitments) against the CPU oracle's MSM on the same terms with the values taken mod r -- bits, sums of a few bits,
    the widest digits allowed, equal and opposite points inside one bucket, zeros, a point at infinity, all-equal values (one bucket
    run per pseudo-window)."""
    rng = np.random.default_rng(curve * 131 + n * 7 + bits)
    lim = 1 << (bits - 1)
    vals = rng.integers(-lim, lim + 1, size=n).astype(np.int32) if bits > 1 else rng.integers(0, 2, size=n).astype(np.int32)
    bases = oracle.g1_walk(curve, 4242 + n, 3, n)
    if n > 8:
        vals[0] = lim
        vals[1] = -lim
        vals[2] = 0
        bases[4] = bases[3]
        vals[3] = vals[4] = 1   # equal points in one bucket: a doubling
        bases[6] = bases[5]
        vals[5], vals[6] = 1, -1  # P + (-P)
        bases[7] = 0            # infinity
        vals[7] = 1
    if n == 200000 and bits == 2:
        vals[:] = 2             # every term in the same bucket of its pseudo-window
    sc = ints_to_limbs([int(v) % FR[curve] for v in vals], 4) if n else np.zeros((0, 4), dtype=np.uint64)
    got = ctx.msm_g1_small(curve, bases, vals, bits)
    exp = oracle.g1_msm(curve, bases, sc, algo=0)
    assert (got == exp).all(), (curve, n, bits)


def test_small_scalar_msm_rejects_wide_values(ctx, oracle):
    bases = oracle.g1_walk(377, 1, 3, 4)
    with pytest.raises(zk.ZkAesError):
        ctx.msm_g1_small(377, bases, np.array([0, 1, 3, 0], dtype=np.int32), 2)   # |3| > 2^(2-1)
    with pytest.raises(zk.ZkAesError):
        ctx.msm_g1_small(377, bases, np.array([0, 1, 1, 0], dtype=np.int32), 14)  # digit width out of range


@pytest.mark.parametrize("window", [3, 5, 8, 11, 13, 16])
def test_msm_window_sizes(ctx, oracle, window):
    curve, n = 377, 3000
    rng = np.random.default_rng(window)
    bases = oracle.g1_walk(curve, 99, 5, n)
    sc = rand_fr(rng, curve, n)
    exp = oracle.g1_msm(curve, bases, sc)
    ctx.set_msm_window(window)
    try:
        assert (ctx.msm_g1(curve, bases, sc) == exp).all()
    finally:
        ctx.set_msm_window(0)


@pytest.mark.parametrize("tuning", [{"msm_acc_blocks": 4}, {"msm_window_max": 12}, {"msm_window_max": 22, "msm_acc_blocks": 3}, {"msm_pair_round": 0},
                                    {"msm_pair_round": 1, "msm_window_max": 10}, {"msm_pair_round": 2, "msm_window_max": 10},
                                    {"msm_pair_round": 3, "msm_window_max": 9}, {"msm_madd_call": 0}, {"msm_madd_call": 1},
                                    {"msm_madd_call": 1, "msm_acc_blocks": 4}, {"msm_madd_call": 0, "msm_acc_blocks": 4},
                                    {"msm_prefetch": 1}, {"msm_prefetch": 2}, {"msm_prefetch": 1, "msm_window_max": 10}, {"msm_prefetch": 2, "msm_window_max": 10}])
def test_msm_tuning_knobs_do_not_change_results(ctx, oracle, tuning):
    """70,000 terms: large enough for the batched-affine pair round (on by default) -- checked against the oracle with the
    round on and off, with degenerate pairs in the buckets (equal points, opposite points, infinity, repeated scalars)."""
    curve, n = 377, 70000
    rng = np.random.default_rng(5)
    bases = oracle.g1_walk(curve, 17, 3, n)
    sc = rand_fr(rng, curve, n)
    sc[:64] = sc[64]            # a bucket heavy enough to straddle several accumulation slices
    bases[100:164] = bases[100]  # the same point 64 times with the same scalar: doublings inside pairs
    sc[100:164] = sc[100]
    bases[201] = bases[200]      # P and -P in the same buckets: s P + (r - s) P
    sc[201] = ints_to_limbs([FR[curve] - limbs_to_ints(sc[200:201])[0]], 4)[0]
    bases[300] = 0               # infinity
    sc[301] = 0
    exp = oracle.g1_msm(curve, bases, sc)
    try:
        for k, v in tuning.items():
            ctx.set_tuning(k, v)
        assert (ctx.msm_g1(curve, bases, sc) == exp).all()
    finally:
        for k, v in TUNING_DEFAULTS.items():
            ctx.set_tuning(k, v)
    with pytest.raises(Exception):
        ctx.set_tuning("no_such_knob", 1)


def test_msm_edge_cases(ctx, oracle):
    curve = 377
    bases = oracle.g1_walk(curve, 3, 2, 16)
    zero = np.zeros((16, 4), dtype=np.uint64)
    assert (ctx.msm_g1(curve, bases, zero) == 0).all()
    assert (ctx.msm_g1(curve, bases[:0], zero[:0]) == 0).all()
    ones = np.zeros((16, 4), dtype=np.uint64)
    ones[:, 0] = 1
    assert (ctx.msm_g1(curve, bases, ones) == oracle.g1_msm(curve, bases, ones, algo=1)).all()
    # all scalars equal and maximal: every term lands in the same buckets
    big = np.tile(ints_to_limbs([FR[curve] - 1], 4), (16, 1))
    assert (ctx.msm_g1(curve, bases, big) == oracle.g1_msm(curve, bases, big)).all()


def test_srs_powers_across_normalisation_chunks(ctx, oracle):
    """The fixed-base kernels leave XYZZ sums in a chunk buffer of 2^22 points and a second kernel normalises runs of 32 points with one
    inversion each: points on both sides of the run and chunk boundaries against the oracle's tau^i G, and every point on the curve."""
    import torch

    curve, n = 377, (1 << 22) + 37
    r = FR[curve]
    seed = bytes(range(7, 39))
    tau = int.from_bytes(seed, "little") & ((1 << 252) - 1)
    d_bases = torch.empty(n * 96, dtype=torch.uint8, device="cuda")
    ctx.srs_powers_device(curve, seed, n, d_bases)
    ctx.sync()
    bases = d_bases.cpu().numpy().view(np.uint64).reshape(n, 12)
    idx = [0, 1, 31, 32, 33, 63, 64, (1 << 22) - 33, (1 << 22) - 1, 1 << 22, (1 << 22) + 1, (1 << 22) + 31, (1 << 22) + 32, n - 1]
    exp = oracle.g1_mul_gen(curve, ints_to_limbs([pow(tau, i, r) for i in idx], 4))
    assert (bases[idx] == exp).all()
    assert oracle.g1_on_curve(curve, bases[:: 257])
    assert oracle.g1_on_curve(curve, bases[(1 << 22) - 64:])


def test_srs_powers_and_large_msm_properties(ctx, oracle):
    """2^18 terms: SRS generated on the device, checked on-curve + first points against the oracle; MSM checked by
    linearity (msm(a)+msm(b) == msm(a+b mod r)) and by the trapdoor identity msm(s, tau^i G) == (sum s_i tau^i) G."""
    import torch

    curve, n = 377, 1 << 18
    r = FR[curve]
    seed = bytes(range(32))
    tau = int.from_bytes(seed, "little") & ((1 << 252) - 1)
    d_bases = torch.empty(n * 96, dtype=torch.uint8, device="cuda")
    ctx.srs_powers_device(curve, seed, n, d_bases)
    ctx.sync()
    bases = d_bases.cpu().numpy().view(np.uint64).reshape(n, 12)
    assert oracle.g1_on_curve(curve, bases)
    first = oracle.g1_mul_gen(curve, ints_to_limbs([pow(tau, i, r) for i in range(6)], 4))
    assert (bases[:6] == first).all()
    last = oracle.g1_mul_gen(curve, ints_to_limbs([pow(tau, n - 1, r)], 4))
    assert (bases[n - 1] == last[0]).all()

    rng = np.random.default_rng(1)
    a, b = rand_fr(rng, curve, n), rand_fr(rng, curve, n)
    ai, bi = limbs_to_ints(a), limbs_to_ints(b)
    s = ints_to_limbs([(x + y) % r for x, y in zip(ai, bi)], 4)
    d_a, d_b, d_s = (torch.from_numpy(v.view(np.int64)).cuda() for v in (a, b, s))
    ma = ctx.msm_g1_device(curve, d_bases, d_a, n)
    mb = ctx.msm_g1_device(curve, d_bases, d_b, n)
    ms = ctx.msm_g1_device(curve, d_bases, d_s, n)
    assert (oracle.g1_add(curve, ma, mb) == ms).all()
    # trapdoor identity (Horner in python ints)
    acc = 0
    for x in reversed(ai):
        acc = (acc * tau + x) % r
    assert (ma == oracle.g1_mul_gen(curve, ints_to_limbs([acc], 4))[0]).all()
    # Montgomery-form scalars give the same result
    a_m = oracle.to_mont(curve, 0, a)
    d_am = torch.from_numpy(a_m.view(np.int64)).cuda()
    assert (ctx.msm_g1_device(curve, d_bases, d_am, n, scalars_montgomery=True) == ma).all()
    # split across two "ranks" by point range and fold
    wb = ctx.msm_g1_windows_bytes(curve, n)
    win = torch.zeros(2 * wb, dtype=torch.uint8, device="cuda")
    h = n // 2
    ctx.msm_g1_windows(curve, d_bases, d_a, h, n, win[:wb])
    ctx.msm_g1_windows(curve, d_bases[h * 96:], d_a[h:], n - h, n, win[wb:])
    assert (ctx.msm_g1_fold(curve, win, 2, n) == ma).all()
