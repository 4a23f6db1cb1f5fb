"""verify_encryption (host verifier, csrc/verifier.cpp + csrc/pairing.h) -- CPU tier, no GPU needed.

  * the C++ pairing against the big-integer model (tools/pairing_model.py): GT values bit for bit, bilinearity;
  * the reference's own verifier assertions (tests/integration_tests.rs:313-372: accept the proof for the right ciphertext,
    reject it for a wrong one) on the golden 16-byte proof, with a verifying key built independently by the oracle."""
import json
import os

import pytest

import aes_zero_knowledge_proof_circuit_b200 as zk
from tools import pairing_model as pr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "marlin_proof_16B.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLD) as f:
        return json.load(f)


def _gt_bytes(e):
    return b"".join(c.to_bytes(48, "little") for f2 in e for c in f2)


def test_pairing_matches_big_integer_model():
    # NOTE on independence: tools/pairing_model.py also GENERATES the product's pairing constants
    # (twist coefficient, G2 generator, final exponent).  This test therefore pins the C++ tower / Miller loop arithmetic, not the constants;
    # the constants are pinned by test_accepts_golden_proof_and_rejects_wrong_ciphertext and the GPU tier, where the pairing verifier must
    # agree with the oracle's TRAPDOOR verifier (G1 only, no G2 constant involved) on every accept / reject.
    e = pr.pairing(pr.G1, pr.G2)
    assert zk.pairing_selftest(1, 1) == _gt_bytes(e)
    a, b = 0x1234567890ABCDEF1234567890ABCDEF, pr.r - 5
    assert zk.pairing_selftest(a, b) == _gt_bytes(pr.f12pow(e, a * b % pr.r))       # bilinear in both arguments
    assert zk.pairing_selftest(0, 7) == _gt_bytes(pr.F12ONE)                         # infinity pairs to one
    assert zk.pairing_selftest(7, 0) == _gt_bytes(pr.F12ONE)


def test_structured_final_exponentiation_is_the_cube_of_the_plain_power():
    """csrc/pairing.h checks pairing equations with the BLS12 hard-part chain (five powers by x, Frobenius maps) instead of the 4,300-bit
    power: tests/pairing_check.cpp requires it to equal the plain power cubed bit for bit, the Frobenius maps to be the q-th / q^2-th powers,
    the inverse to invert, and e(aP, bQ) e(-abP, Q) == 1 (and != 1 one step off) through the equation form the verifier calls."""
    import ctypes
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "pairing_check.cpp")
    out = os.path.join(root, "tests", "_pairing_check.so")
    hdr = os.path.join(root, "aes_zero_knowledge_proof_circuit_b200", "csrc", "pairing.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    lib = ctypes.CDLL(out)
    lib.pairing_fast_check.restype = ctypes.c_int
    lib.pairing_fast_check.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
    for a, b in ((1, 1), (0x1234567, 0xFEDCBA9876543), (2**63 + 5, 3)):
        assert lib.pairing_fast_check(a, b) == 0, (a, b)


def test_accepts_golden_proof_and_rejects_wrong_ciphertext(golden):
    vk, proof, ct = bytes.fromhex(golden["verifying_key"]), bytes.fromhex(golden["proof"]), bytes.fromhex(golden["ciphertext"])
    assert zk.verify_encryption(vk, proof, ct) is True
    bad = bytearray(ct)
    bad[5] ^= 0x10
    assert zk.verify_encryption(vk, proof, bytes(bad)) is False                      # tests/integration_tests.rs:332-336
    assert zk.verify_encryption(vk, proof, ct[:15]) is False                         # a shorter statement is a different statement


def test_statement_length_follows_ark_marlins_domain_rule(golden):
    """ark-marlin 0.3.0 takes domain_x from public_input.len() + 1 and zero-pads the input to |X| - 1 itself (verify()), so the
    statement is bound to the key through the power-of-two bracket of its bit count, exactly as in the reference: lengths from
    another bracket are rejected whatever their bytes, also when the dropped / added bytes are zero."""
    vk, proof, ct = bytes.fromhex(golden["verifying_key"]), bytes.fromhex(golden["proof"]), bytes.fromhex(golden["ciphertext"])
    assert zk.verify_encryption(vk, proof, ct + bytes(16)) is False                  # 257 inputs: domain 512, not this key's 256
    assert zk.verify_encryption(vk, proof, bytes(15)) is False and zk.verify_encryption(vk, proof, b"") is False
    zero_tail = ct[:15] + b"\x00"
    assert zk.verify_encryption(vk, proof, zero_tail[:15]) is False                  # truncation below the bracket is rejected even
    assert zk.verify_encryption(vk, proof, zero_tail) is (zero_tail == ct)           # ... when the dropped byte is zero
    # inside the bracket the reference itself pads with zeros: ct || 00 is the same padded statement for ark-marlin (and here)
    assert zk.verify_encryption(vk, proof, ct + b"\x00") is True
    assert zk.verify_encryption(vk, proof, ct + b"\x01") is False


def test_rejects_tampered_proofs(golden):
    vk, proof, ct = bytes.fromhex(golden["verifying_key"]), bytearray.fromhex(golden["proof"]), bytes.fromhex(golden["ciphertext"])
    # one evaluation changed (offset: 3 rounds of commitments = 8 + (8 + 4*49) + (8 + 3*49 + 48) + (8 + 2*49 + 48), then the count)
    ev0 = 8 + (8 + 4 * 49) + (8 + 3 * 49 + 48) + (8 + 2 * 49 + 48) + 8
    t = bytearray(proof)
    t[ev0 + 3] ^= 1
    assert zk.verify_encryption(vk, bytes(t), ct) is False
    # the opening witness at gamma replaced by the one at beta
    w_beta = len(proof) - (48 + 33) - (48 + 1) - 1
    t = bytearray(proof)
    t[w_beta + 81: w_beta + 81 + 48] = proof[w_beta: w_beta + 48]
    assert zk.verify_encryption(vk, bytes(t), ct) is False
    # the hiding evaluation random_v changed
    t = bytearray(proof)
    t[w_beta + 49 + 2] ^= 4
    assert zk.verify_encryption(vk, bytes(t), ct) is False
    # both opening witnesses exchanged: the two KZG equations are checked as ONE randomised product (KZG10::batch_check); two wrong
    # equations must not cancel
    t = bytearray(proof)
    t[w_beta: w_beta + 48], t[w_beta + 81: w_beta + 81 + 48] = proof[w_beta + 81: w_beta + 81 + 48], proof[w_beta: w_beta + 48]
    assert zk.verify_encryption(vk, bytes(t), ct) is False


def test_malformed_inputs_are_errors(golden):
    vk, proof, ct = bytes.fromhex(golden["verifying_key"]), bytes.fromhex(golden["proof"]), bytes.fromhex(golden["ciphertext"])
    with pytest.raises(zk.ZkAesError):
        zk.verify_encryption(vk, proof[:-7], ct)            # truncated proof
    with pytest.raises(zk.ZkAesError):
        zk.verify_encryption(b"not a key" + vk, proof, ct)
    t = bytearray(proof)
    t[16 + 47] |= 0x3F                                     # first commitment: x >= q
    t[16 + 46] = 0xFF
    with pytest.raises(zk.ZkAesError):
        zk.verify_encryption(vk, bytes(t), ct)
    # a proof for another key: same shape, a commitment swapped -> parses, does not verify
    t = bytearray(proof)
    t[16:16 + 48], t[16 + 49:16 + 49 + 48] = proof[16 + 49:16 + 49 + 48], proof[16:16 + 48]
    assert zk.verify_encryption(vk, bytes(t), ct) is False


def _proof_offsets(proof):
    """byte offsets inside the golden proof: first commitment, its has_shifted byte, prover-message count, first opening"""
    msgs = 8 + (8 + 4 * 49) + (8 + 3 * 49 + 48) + (8 + 2 * 49 + 48) + 8 + 7 * 32
    return {"comm0": 16, "shifted0": 16 + 48, "msg_count": msgs, "msg0": msgs + 8, "open_count": msgs + 8 + 3}


def test_non_canonical_encodings_are_refused(golden):
    """the verifier and zkaes_proof_deserialize share ONE strict reader -- a proof has exactly one accepted encoding"""
    vk, proof, ct = bytes.fromhex(golden["verifying_key"]), bytes.fromhex(golden["proof"]), bytes.fromhex(golden["ciphertext"])
    o = _proof_offsets(proof)
    assert int.from_bytes(proof[o["msg_count"]:o["msg_count"] + 8], "little") == 3 and proof[o["msg0"]:o["msg0"] + 3] == b"\0\0\0"
    cases = {}
    # (a) an empty prover message replaced by FieldElements([42])
    cases["message with elements"] = proof[:o["msg0"]] + b"\x01" + (1).to_bytes(8, "little") + (42).to_bytes(32, "little") + proof[o["msg0"] + 1:]
    # (a') ... or by FieldElements([]): absorbs nothing, ark-marlin would accept it as a second encoding of the same proof
    cases["message with empty vector"] = proof[:o["msg0"]] + b"\x01" + (0).to_bytes(8, "little") + proof[o["msg0"] + 1:]
    # (b) five prover messages
    cases["five messages"] = proof[:o["msg_count"]] + (5).to_bytes(8, "little") + b"\0" * 5 + proof[o["msg0"] + 3:]
    # (c) Option / bool tags other than 0 / 1
    t = bytearray(proof); t[o["shifted0"]] = 2; cases["has_shifted = 2"] = bytes(t)
    t = bytearray(proof); t[o["msg0"]] = 2; cases["message tag = 2"] = bytes(t)
    t = bytearray(proof); t[-1] = 2; cases["BatchLCProof tag = 2"] = bytes(t)
    # (d) point encodings ark rejects or never emits
    t = bytearray(proof); t[o["comm0"] + 47] |= 0xC0; cases["infinity and sign flags"] = bytes(t)
    t = bytearray(proof); t[o["comm0"] + 47] = (t[o["comm0"] + 47] & 0x3F) | 0x40; cases["infinity with non-zero x"] = bytes(t)
    # (e) a point of the curve outside the prime-order subgroup (BLS12-377 G1 has a cofactor): smallest x with x^3 + 1 a square
    from oracle.cpu import FQ
    q = FQ[377]
    x = next(x for x in range(2, 100) if pow(x ** 3 + 1, (q - 1) // 2, q) == 1)
    t = bytearray(proof); t[o["comm0"]:o["comm0"] + 48] = x.to_bytes(48, "little"); cases["outside the subgroup"] = bytes(t)
    for what, data in cases.items():
        with pytest.raises(zk.ZkAesError):
            zk.verify_encryption(vk, data, ct)
        with pytest.raises(zk.ZkAesError):
            zk.deserialize_proof(data)
    assert zk.verify_encryption(vk, proof, ct) is True


def test_implausible_verifying_keys_are_errors_not_hangs(golden):
    vk, proof, ct = bytearray.fromhex(golden["verifying_key"]), bytes.fromhex(golden["proof"]), bytes.fromhex(golden["ciphertext"])
    for off, val in ((8, (1 << 63) + 1), (16, (1 << 63) + 1), (24, 1 << 62), (8, 0)):   # num_constraints, num_non_zero, |X|
        t = bytearray(vk)
        t[off:off + 8] = val.to_bytes(8, "little")
        with pytest.raises(zk.ZkAesError):
            zk.verify_encryption(bytes(t), proof, ct)
    t = bytearray(vk)
    t[-16:-8] = (7).to_bytes(8, "little")                                              # SRS degree too small for the index
    with pytest.raises(zk.ZkAesError):
        zk.verify_encryption(bytes(t), proof, ct)
    with pytest.raises(zk.ZkAesError):
        zk.verify_encryption(bytes(vk) + b"\0", proof, ct)


def test_second_golden_proof_and_cross_statements(golden):
    """FIPS-197 Appendix C.1 under the same keys: accepted for its own ciphertext, and neither proof verifies the other statement"""
    with open(os.path.join(os.path.dirname(GOLD), "marlin_proof_16B_fips_c1.json")) as f:
        g2 = json.load(f)
    assert g2["ciphertext"] == "69c4e0d86a7b0430d8cdb78070b4c55a" and g2["verifying_key"] == golden["verifying_key"]
    vk = bytes.fromhex(golden["verifying_key"])
    p1, c1 = bytes.fromhex(golden["proof"]), bytes.fromhex(golden["ciphertext"])
    p2, c2 = bytes.fromhex(g2["proof"]), bytes.fromhex(g2["ciphertext"])
    assert zk.verify_encryption(vk, p2, c2) is True
    assert zk.verify_encryption(vk, p2, c1) is False
    assert zk.verify_encryption(vk, p1, c2) is False


def test_two_block_golden_proof(golden):
    """32-byte message (two ECB blocks; |H| = 2^19, |X| = 512): the product's host verifier accepts the oracle prover's proof under the oracle's
    key bytes, rejects a flipped ciphertext bit and a truncated statement, and a one-block key does not verify it"""
    with open(os.path.join(os.path.dirname(GOLD), "marlin_proof_32B.json")) as f:
        g = json.load(f)
    vk, proof, ct = bytes.fromhex(g["verifying_key"]), bytes.fromhex(g["proof"]), bytes.fromhex(g["ciphertext"])
    assert (g["h"], g["k"], g["x"]) == (1 << 19, 1 << 20, 512) and len(ct) == 32
    assert ct[:16].hex() == golden["ciphertext"]  # ECB: the first block is the one-block fixture's
    assert zk.verify_encryption(vk, proof, ct) is True
    bad = bytearray(ct)
    bad[17] ^= 0x20
    assert zk.verify_encryption(vk, proof, bytes(bad)) is False
    assert zk.verify_encryption(vk, proof, ct[:16]) is False
    assert zk.verify_encryption(bytes.fromhex(golden["verifying_key"]), proof, ct[:16]) is False


@pytest.mark.parametrize("name", ["marlin_proof_16B.json", "marlin_proof_16B_fips_c1.json", "marlin_proof_32B.json"])
def test_proof_wire_format_round_trip(name):
    """deserialize_proof / serialize_proof (src/lib.rs:52): fields equal the oracle's reading of the same bytes, and packing
    them again reproduces the bytes"""
    from oracle import marlin_oracle as mo

    with open(os.path.join(os.path.dirname(GOLD), name)) as f:
        proof = bytes.fromhex(json.load(f)["proof"])
    fields = zk.deserialize_proof(proof)
    assert (fields.n_rounds, list(fields.round_sizes), fields.n_evaluations, fields.n_openings) == (3, [4, 3, 2], 7, 2)
    ref = mo.deserialize_proof(proof)
    flat = [c for rnd in ref["commitments"] for c in rnd]
    xy = lambda pt: b"".join(int(v).to_bytes(48, "little") for v in mo.g1_xy(pt))
    for got, (comm, shifted) in zip(fields.commitments, flat):
        assert bytes(got.comm) == xy(comm)
        assert bool(got.has_shifted) == (shifted is not None)
        if shifted is not None:
            assert bytes(got.shifted) == xy(shifted)
    assert [bool(c.has_shifted) for c in fields.commitments] == [False] * 5 + [True, False, True, False]   # g_1 and g_2 are degree-bounded
    assert [int.from_bytes(bytes(e), "little") for e in fields.evaluations] == ref["evaluations"]
    for got, o in zip(fields.openings, ref["pc_proof"]):
        assert bytes(got.w) == xy(o["w"])
        assert bool(got.has_random_v) == (o["random_v"] is not None)
        if o["random_v"] is not None:
            assert int.from_bytes(bytes(got.random_v), "little") == o["random_v"]
    assert zk.serialize_proof(fields) == proof
    # a point moved off the curve is refused when packing; a truncated proof when unpacking
    fields.commitments[0].comm[3] ^= 1
    with pytest.raises(zk.ZkAesError):
        zk.serialize_proof(fields)
    with pytest.raises(zk.ZkAesError):
        zk.deserialize_proof(proof[:-1])
