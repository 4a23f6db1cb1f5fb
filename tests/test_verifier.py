"""verify_encryption (host verifier, csrc/verifier.cpp + csrc/pairing.h) -- CPU tier, no GPU needed.

  * the C++ pairing against the big-integer model (oracle/pairing_ref.py): GT values bit for bit, bilinearity;
  * the reference's own verifier assertions (tests/integration_tests.rs:313-372: accept the proof for the right ciphertext,
    reject it for a wrong one) on the golden 16-byte proof, with a verifying key built independently by the oracle."""
import json
import os

import pytest

import aes_zero_knowledge_proof_circuit_b200 as zk
from oracle import pairing_ref as pr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "marlin_proof_16B.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLD) as f:
        return json.load(f)


def _gt_bytes(e):
    return b"".join(c.to_bytes(48, "little") for f2 in e for c in f2)


def test_pairing_matches_big_integer_model():
    e = pr.pairing(pr.G1, pr.G2)
    assert zk.pairing_selftest(1, 1) == _gt_bytes(e)
    a, b = 0x1234567890ABCDEF1234567890ABCDEF, pr.r - 5
    assert zk.pairing_selftest(a, b) == _gt_bytes(pr.f12pow(e, a * b % pr.r))       # bilinear in both arguments
    assert zk.pairing_selftest(0, 7) == _gt_bytes(pr.F12ONE)                         # infinity pairs to one
    assert zk.pairing_selftest(7, 0) == _gt_bytes(pr.F12ONE)


def test_accepts_golden_proof_and_rejects_wrong_ciphertext(golden):
    vk, proof, ct = bytes.fromhex(golden["verifying_key"]), bytes.fromhex(golden["proof"]), bytes.fromhex(golden["ciphertext"])
    assert zk.verify_encryption(vk, proof, ct) is True
    bad = bytearray(ct)
    bad[5] ^= 0x10
    assert zk.verify_encryption(vk, proof, bytes(bad)) is False                      # tests/integration_tests.rs:332-336
    assert zk.verify_encryption(vk, proof, ct[:15]) is False                         # a shorter statement is a different statement


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


@pytest.mark.parametrize("name", ["marlin_proof_16B.json", "marlin_proof_16B_fips_c1.json"])
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
