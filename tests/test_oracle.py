"""CPU tests that PIN THE ORACLE: against every golden vector the reference's own tests hold for the path
(tests/golden/*.json, extracted by tools/extract_reference_vectors.py) and against independent definitions
(big-int arithmetic, double-and-add, O(n^2) DFT) where the reference holds none."""
import json
import os
import random

import numpy as np
import pytest

from oracle.cpu import FQ, FR, ints_to_limbs, limbs_to_ints, rand_fr

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return json.load(open(os.path.join(GOLD, name)))


# ---------------- AES bytes: reference tests/integration_tests.rs:50-310, src/aes_circuit.rs:704-846 ----------------
def test_fips197_round_trace(oracle):
    t = load("fips197_round_trace.json")
    pt, key = bytes.fromhex(t["plaintext"]), bytes.fromhex(t["key"])
    ct, trace = oracle.aes128_ecb(pt, key, trace=True)
    assert ct.hex() == t["ciphertext"]
    states = [trace[16 * i:16 * i + 16].hex() for i in range(40)]
    assert states[0] == t["start_of_round"][0]
    k = 1
    for rnd in range(1, 10):
        assert states[k] == t["after_sub_bytes"][rnd - 1]
        assert states[k + 1] == t["after_shift_rows"][rnd - 1]
        assert states[k + 2] == t["after_mix_columns"][rnd - 1]
        assert states[k + 3] == t["start_of_round"][rnd]
        k += 4
    assert states[k] == t["after_sub_bytes"][9]
    assert states[k + 1] == t["after_shift_rows"][9]
    assert states[k + 2] == t["start_of_round"][10] == t["ciphertext"]


def test_gadget_step_vectors(oracle):
    s = load("gadget_steps.json")
    a = s["add_round_key"]
    assert oracle.aes_step("add_round_key", bytes.fromhex(a["input"]), bytes.fromhex(a["key"])).hex() == a["output"]
    m = s["mix_columns"]
    assert oracle.aes_step("mix_columns", bytes.fromhex(m["input"])).hex() == m["output"]
    b = s["sub_bytes"]
    assert oracle.aes_step("sub_bytes", bytes.fromhex(b["input"])).hex() == b["output"]
    k = s["key_expansion"]
    rk = oracle.aes_step("derive_keys", bytes.fromhex(k["key"]))
    assert rk[:16].hex() == k["key"] and rk[160:176].hex() == k["round_key_10"]
    # the algorithmic S-box (src/aes.rs:24-62) equals the constant table of lookup_table (src/aes_circuit.rs:433-694)
    import ctypes
    sb = ctypes.create_string_buffer(256)
    oracle.lib.orc_aes_sbox(sb)
    assert sb.raw.hex() == s["lookup_table"]


def test_shift_rows_permutation(oracle):
    # src/aes_circuit.rs:762-796 / src/aes.rs:289-303: explicit index map on random bytes
    perm = [0, 5, 10, 15, 4, 9, 14, 3, 8, 13, 2, 7, 12, 1, 6, 11]
    rnd = random.Random(7)
    for _ in range(8):
        x = bytes(rnd.randrange(256) for _ in range(16))
        assert oracle.aes_step("shift_rows", x) == bytes(x[p] for p in perm)


def test_encrypt_e2e_ciphertexts(oracle):
    for case in load("encrypt_e2e.json"):
        ct = oracle.aes128_ecb(bytes.fromhex(case["plaintext"]), bytes.fromhex(case["key"]))
        assert ct.hex() == case["ciphertext"]
        if case["wrong_ciphertext"]:
            assert ct.hex() != case["wrong_ciphertext"]


def test_ragged_message_rejected(oracle):
    # `chunks(16)` would hand add_round_key a short block -> ensure!(len == 16) (src/aes_circuit.rs:218-221)
    with pytest.raises(ValueError):
        oracle.aes128_ecb(bytes(17), bytes(16))
    assert oracle.aes128_ecb(b"", bytes(16)) == b""


def test_against_independent_aes(oracle):
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

    rnd = random.Random(11)
    for nblk in (1, 3, 16):
        key = bytes(rnd.randrange(256) for _ in range(16))
        msg = bytes(rnd.randrange(256) for _ in range(16 * nblk))
        enc = Cipher(algorithms.AES(key), modes.ECB()).encryptor()
        assert oracle.aes128_ecb(msg, key) == enc.update(msg) + enc.finalize()


# ---------------- field / curve constants ----------------
@pytest.mark.parametrize("curve", [377, 381])
def test_constants(oracle, curve):
    r, q = FR[curve], FQ[curve]
    assert limbs_to_ints(oracle.constant(curve, 0)[None])[0] == r
    assert limbs_to_ints(oracle.constant(curve, 1)[None])[0] == q
    assert limbs_to_ints(oracle.constant(curve, 2)[None])[0] == (1 << 256) % r
    assert limbs_to_ints(oracle.constant(curve, 3)[None])[0] == (1 << 384) % q
    s, g = (47, 22) if curve == 377 else (32, 7)
    root = limbs_to_ints(oracle.constant(curve, 4)[None])[0] * pow(1 << 256, -1, r) % r
    assert root == pow(g, (r - 1) >> s, r)
    expect = {377: 8065159656716812877374967518403273466521432693661810619979959746626482506078,
              381: 10238227357739495823651030575849232062558860180284477541189508159991286009131}[curve]
    assert root == expect  # arkworks' TWO_ADIC_ROOT_OF_UNITY
    gen = oracle.constant(curve, 6)[None]
    assert oracle.g1_on_curve(curve, gen)


@pytest.mark.parametrize("curve", [377, 381])
@pytest.mark.parametrize("field", [0, 1])
def test_field_ops_vs_bigint(oracle, curve, field):
    p = (FR if field == 0 else FQ)[curve]
    nl = 4 if field == 0 else 6
    R = 1 << (64 * nl)
    rnd = random.Random(curve + field)
    vals = [0, 1, p - 1, p - 2, 2] + [rnd.randrange(p) for _ in range(200)]
    a = ints_to_limbs(vals, nl)
    b = ints_to_limbs(list(reversed(vals)), nl)
    bv = list(reversed(vals))
    Ri = pow(R, -1, p)
    assert limbs_to_ints(oracle.field_op(curve, field, 0, a, b)) == [(x + y) % p for x, y in zip(vals, bv)]
    assert limbs_to_ints(oracle.field_op(curve, field, 1, a, b)) == [(x - y) % p for x, y in zip(vals, bv)]
    assert limbs_to_ints(oracle.field_op(curve, field, 2, a, b)) == [x * y * Ri % p for x, y in zip(vals, bv)]
    inv = limbs_to_ints(oracle.field_op(curve, field, 3, a))
    assert all((x * y - R * R) % p == 0 for x, y in zip(vals, inv) if x)
    assert limbs_to_ints(oracle.to_mont(curve, field, a)) == [x * R % p for x in vals]


# ---------------- MSM: ark-ec Pippenger restatement == definition ----------------
@pytest.mark.parametrize("curve", [377, 381])
@pytest.mark.parametrize("n", [1, 2, 31, 32, 257])
def test_msm_pippenger_equals_definition(oracle, curve, n):
    rng = np.random.default_rng(curve * 1000 + n)
    r = FR[curve]
    ks = rand_fr(rng, curve, n)
    pts = oracle.g1_mul_gen(curve, ks)
    assert oracle.g1_on_curve(curve, pts)
    sc = rand_fr(rng, curve, n)
    if n > 4:
        sc[1] = 0
        sc[2] = ints_to_limbs([1], 4)[0]
        sc[3] = ints_to_limbs([r - 1], 4)[0]
    fast = oracle.g1_msm(curve, pts, sc, algo=0)
    slow = oracle.g1_msm(curve, pts, sc, algo=1)
    assert (fast == slow).all()
    # discrete-log check: sum k_i s_i * G
    tot = sum(k * s for k, s in zip(limbs_to_ints(ks), limbs_to_ints(sc))) % r
    assert (fast == oracle.g1_mul_gen(curve, ints_to_limbs([tot], 4))[0]).all()


def test_msm_empty_and_zero(oracle):
    pts = oracle.g1_walk(377, 5, 3, 8)
    z = np.zeros((8, 4), dtype=np.uint64)
    assert (oracle.g1_msm(377, pts, z) == 0).all()  # infinity encoded as zeros
    assert (oracle.g1_msm(377, pts[:0], z[:0]) == 0).all()


def test_walk_points(oracle):
    pts = oracle.g1_walk(377, 7, 3, 50)
    assert oracle.g1_on_curve(377, pts)
    exp = oracle.g1_mul_gen(377, ints_to_limbs([7 + 3 * i for i in range(50)], 4))
    assert (pts == exp).all()


# ---------------- NTT: ark-poly restatement == O(n^2) DFT ----------------
@pytest.mark.parametrize("curve", [377, 381])
@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 9])
@pytest.mark.parametrize("coset", [False, True])
def test_ntt_equals_dft(oracle, curve, log_n, coset):
    rng = np.random.default_rng(curve + log_n)
    x = rand_fr(rng, curve, 1 << log_n)
    y = oracle.ntt(curve, x, coset=coset)
    assert (y == oracle.dft_naive(curve, x, coset=coset)).all()
    assert (oracle.ntt(curve, y, inverse=True, coset=coset) == x).all()


def test_ntt_polynomial_identity(oracle):
    # evaluations of c0 + c1 X over the size-4 domain, against python big ints
    curve, r = 377, FR[377]
    R = 1 << 256
    c = [5, 7, 0, 0]
    x = oracle.to_mont(curve, 0, ints_to_limbs(c, 4))
    ev = limbs_to_ints(oracle.from_mont(curve, 0, oracle.ntt(curve, x)))
    w = pow(22, (r - 1) >> 2, r)
    assert ev == [(5 + 7 * pow(w, i, r)) % r for i in range(4)]
