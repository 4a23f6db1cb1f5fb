"""encrypt() parity (GPU): the device-resident Marlin prover through the C ABI against
  * the golden proof produced by the CPU oracle (tools/gen_golden_proof.py) -- byte for byte, vk included;
  * the oracle's verifier (accept on the right ciphertext, reject on a flipped one) at 16 / 64 / 256 bytes, mirroring
    the reference's only prover-boundary assertions (tests/integration_tests.rs:313-372)."""
import hashlib
import json
import os

import numpy as np
import pytest

import aes_zero_knowledge_proof_circuit_b200 as zk
from oracle import marlin_oracle as mo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TAU, GAMMA = bytes(range(32)), bytes(range(1, 33))


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLD, "marlin_proof_16B.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def pk16(ctx):
    pk = ctx.synthesize_keys(16, TAU, GAMMA)
    yield pk
    pk.close()


def test_key_matches_oracle_golden(pk16, golden):
    assert (pk16.info["h"], pk16.info["k"], pk16.info["x"], pk16.info["max_degree"]) == (golden["h"], golden["k"], golden["x"], golden["max_degree"])
    vk = pk16.vk_bytes()
    assert len(vk) == golden["vk_len"]
    assert hashlib.sha256(vk).hexdigest() == golden["vk_sha256"]


def test_verifying_key_matches_oracle_golden(pk16, golden):
    """the VerifyingKey half of synthesize_keys: product bytes == the blob the oracle builds independently"""
    assert pk16.verifying_key().hex() == golden["verifying_key"]


def test_proof_bytes_match_oracle_golden(ctx, pk16, golden):
    ct, proof = ctx.encrypt(pk16, bytes.fromhex(golden["message"]), bytes.fromhex(golden["key"]), bytes.fromhex(golden["zk_seed"]))
    assert ct.hex() == golden["ciphertext"]
    assert proof.hex() == golden["proof"]
    # deterministic given the seed; a different zk seed gives a different (still valid) proof
    ct2, proof2 = ctx.encrypt(pk16, bytes.fromhex(golden["message"]), bytes.fromhex(golden["key"]), bytes([9] * 32))
    assert ct2 == ct and proof2 != proof


def test_lagrange_basis_round1_gives_the_same_proof_bytes(ctx, pk16, golden):
    """Round 1 commits to w, z_A, z_B in the Lagrange basis (one small digit per element of H) when the key holds those points, and
    through the SRS powers (253-bit coefficients) otherwise: same group elements, hence the same golden proof bytes on both paths."""
    assert pk16.info["lagrange_points"] == pk16.info["h"]  # single GPU: every L_k(tau) G lives on this rank
    msg, key, seed = (bytes.fromhex(golden[k]) for k in ("message", "key", "zk_seed"))
    try:
        ctx.set_tuning("r1_lagrange", 0)
        _, powers = ctx.encrypt(pk16, msg, key, seed)
        ctx.set_tuning("r1_lagrange", 1)
        _, lagrange = ctx.encrypt(pk16, msg, key, seed)
    finally:
        ctx.set_tuning("r1_lagrange", 1)
    assert powers.hex() == golden["proof"]
    assert lagrange.hex() == golden["proof"]


def test_scratch_arena_follows_the_key_size(ctx, pk16, golden):
    """The scratch arena is carved for one key's peak (second proof on a context): a key of another size must release it and measure
    again -- a larger circuit must not be squeezed into the small key's arena, and going back must still give the golden bytes."""
    msg, key, seed = (bytes.fromhex(golden[k]) for k in ("message", "key", "zk_seed"))
    for _ in range(3):  # measure, carve, use
        assert ctx.encrypt(pk16, msg, key, seed)[1].hex() == golden["proof"]
    pk64 = ctx.synthesize_keys(64, TAU, GAMMA)
    try:
        m64 = bytes(range(64))
        for _ in range(3):
            ct, proof = ctx.encrypt(pk64, m64, key, seed)
            assert zk.verify_encryption(pk64.verifying_key(), proof, ct)
    finally:
        pk64.close()
    for _ in range(3):
        assert ctx.encrypt(pk16, msg, key, seed)[1].hex() == golden["proof"]


def test_proof_bytes_match_two_block_golden(ctx):
    """32-byte message (two ECB blocks sharing one key schedule; |H| = 2^19, |X| = 512 interleaved with period 1024): key bytes, ciphertext
    and proof bytes equal the oracle prover's, on both round-1 paths"""
    with open(os.path.join(GOLD, "marlin_proof_32B.json")) as f:
        g = json.load(f)
    pk = ctx.synthesize_keys(32, bytes.fromhex(g["tau_seed"]), bytes.fromhex(g["gamma_seed"]))
    try:
        assert (pk.info["h"], pk.info["k"], pk.info["x"], pk.info["max_degree"]) == (g["h"], g["k"], g["x"], g["max_degree"])
        assert hashlib.sha256(pk.vk_bytes()).hexdigest() == g["vk_sha256"]
        assert pk.verifying_key().hex() == g["verifying_key"]
        msg, key, seed = (bytes.fromhex(g[k]) for k in ("message", "key", "zk_seed"))
        for lagrange in (1, 0):
            ctx.set_tuning("r1_lagrange", lagrange)
            ct, proof = ctx.encrypt(pk, msg, key, seed)
            assert ct.hex() == g["ciphertext"]
            assert proof.hex() == g["proof"], f"r1_lagrange={lagrange}"
    finally:
        ctx.set_tuning("r1_lagrange", 1)
        pk.close()


def test_proof_bytes_match_second_golden(ctx, pk16):
    """another message, key and zk seed (FIPS-197 Appendix C.1) under the same proving key: byte-identical to the oracle prover again"""
    with open(os.path.join(GOLD, "marlin_proof_16B_fips_c1.json")) as f:
        g2 = json.load(f)
    ct, proof = ctx.encrypt(pk16, bytes.fromhex(g2["message"]), bytes.fromhex(g2["key"]), bytes.fromhex(g2["zk_seed"]))
    assert ct.hex() == g2["ciphertext"] == "69c4e0d86a7b0430d8cdb78070b4c55a"
    assert proof.hex() == g2["proof"]


def _verify(pk, ct, proof_bytes):
    """Both verifiers must agree: the oracle's (trapdoor check in G1) and the product's verify_encryption (pairing check)."""
    try:
        product = zk.verify_encryption(pk.verifying_key(), proof_bytes, ct)
    except zk.ZkAesError:
        product = None  # unparseable proof
    idx = mo.index_from_vk_bytes(pk.vk_bytes(), pk.info["x"])
    srs = mo.SparseSRS(pk.info["max_degree"], TAU, GAMMA)
    bits = [(b >> i) & 1 for b in ct for i in range(8)]  # src/helpers/mod.rs:84-93
    try:
        oracle = mo.verify(idx, srs, bits, mo.deserialize_proof(proof_bytes))
    except (ValueError, AssertionError):
        oracle = None
    assert product == oracle, (product, oracle)
    return bool(oracle)


@pytest.mark.parametrize("msg_len", [16, 48, 64, 256])  # 48 bytes: three blocks, 385 instance variables padded to 512
def test_verifier_accepts_and_rejects(ctx, msg_len):
    rng = np.random.default_rng(msg_len)
    msg = rng.integers(0, 256, msg_len, dtype=np.uint8).tobytes()
    key = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
    pk = ctx.synthesize_keys(msg_len, TAU, GAMMA)
    try:
        ct, proof = ctx.encrypt(pk, msg, key, bytes([3] * 32))
        assert _verify(pk, ct, proof)
        wrong = bytearray(ct)
        wrong[msg_len // 2] ^= 0x10
        assert not _verify(pk, bytes(wrong), proof)
        tampered = bytearray(proof)
        tampered[-60] ^= 1  # inside the last opening proof
        assert not _verify(pk, ct, bytes(tampered))
    finally:
        pk.close()


def test_full_size_4kib_proof_verifies(ctx):
    """BASELINE.json's headline configuration: 256 ECB blocks, 37,994,400 constraints, |H| = 2^26, |K| = 2^27 on one GPU"""
    msg = bytes((i * 131 + 7) & 0xFF for i in range(4096))
    key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
    pk = ctx.synthesize_keys(4096, TAU, GAMMA)
    try:
        assert (pk.info["h"], pk.info["k"], pk.info["num_constraints"]) == (1 << 26, 1 << 27, 37994400)
        ct, proof = ctx.encrypt(pk, msg, key, bytes([5] * 32))
        assert _verify(pk, ct, proof)
        wrong = bytearray(ct)
        wrong[4095] ^= 1
        assert not _verify(pk, bytes(wrong), proof)
    finally:
        pk.close()


@pytest.mark.parametrize("srs,polys", [(True, True), (False, False), (True, False)])
def test_key_file_round_trip_proves_the_golden_bytes(ctx, pk16, golden, tmp_path, srs, polys):
    """zkaes_pk_save / zkaes_pk_load: a key rebuilt from its file -- with the bulk sections stored, or recomputed from the seeds
    and matrices -- has the same verifying key and produces the golden proof bytes; corrupt files are errors"""
    path = str(tmp_path / "key16.zkpk")
    pk16.save(path, srs=srs, index_polys=polys)
    size = os.path.getsize(path)
    assert (size < 8192) == (not srs and not polys)
    pk2 = ctx.load_keys(path)
    try:
        assert pk2.info == pk16.info and pk2.vk_bytes() == pk16.vk_bytes() and pk2.verifying_key() == pk16.verifying_key()
        ct, proof = ctx.encrypt(pk2, bytes.fromhex(golden["message"]), bytes.fromhex(golden["key"]), bytes.fromhex(golden["zk_seed"]))
        assert ct.hex() == golden["ciphertext"] and proof.hex() == golden["proof"]
    finally:
        pk2.close()
    raw = bytearray(open(path, "rb").read())
    for what, data in (("truncated", raw[:-5]), ("trailing", raw + b"\0"), ("magic", b"X" + raw[1:]),
                       ("commitment", raw[:8 + 64 + 64 + 5] + bytes([raw[8 + 64 + 64 + 5] ^ 1]) + raw[8 + 64 + 64 + 6:])):
        bad = str(tmp_path / f"bad_{what}.zkpk")
        open(bad, "wb").write(bytes(data))
        with pytest.raises(zk.ZkAesError):
            ctx.load_keys(bad).close()
    with pytest.raises(zk.ZkAesError):
        ctx.load_keys(str(tmp_path / "does_not_exist.zkpk"))


def test_wrong_length_rejected(ctx, pk16):
    with pytest.raises(zk.ZkAesError):
        ctx.encrypt(pk16, b"\x00" * 32, b"\x00" * 16, bytes(32))
