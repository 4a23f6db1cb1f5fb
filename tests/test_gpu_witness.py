"""K1 parity (GPU): the AES witness kernel against the reference's FIPS-197 vectors, the oracle's byte-level AES
and the oracle's gadget model (every R1CS variable, bit-exact); R1CS satisfaction at the full 4 KiB size."""
import json
import os

import numpy as np
import pytest

import aes_zero_knowledge_proof_circuit_b200 as zk
from oracle import r1cs_model as model
from tests.test_circuit import KEY, model_r1cs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fips197_and_main_rs_vectors(ctx):
    with open(os.path.join(GOLD, "encrypt_e2e.json")) as f:
        gold = json.load(f)
    for case in gold:
        msg, key = bytes.fromhex(case["plaintext"]), bytes.fromhex(case["key"])
        c = zk.Circuit(len(msg))
        ct, _ = ctx.witness_aes128_ecb(c, msg, key, want_assignment=False)
        assert ct.hex() == case["ciphertext"], case["source"]


@pytest.mark.parametrize("n_blocks", [1, 2])
def test_assignment_matches_gadget_model(ctx, n_blocks):
    msg = bytes((i * 131 + 7) & 0xFF for i in range(16 * n_blocks))
    padded, inst, wit, ct_model = model_r1cs(msg)
    c = zk.Circuit(len(msg))
    ct, z = ctx.witness_aes128_ecb(c, msg, KEY)
    assert ct == ct_model
    exp = np.array(inst + wit, dtype=np.uint8)
    bad = np.nonzero(z != exp)[0]
    assert bad.size == 0, f"{bad.size} variables differ, first at column {bad[:5]}"


def _check_r1cs(c, z):
    """(A z) o (B z) == C z over the integers reduced mod r -- all values here are tiny, so int64 arithmetic is exact"""
    zz = z.astype(np.int64)
    prods = []
    for which in range(3):
        row_ptr, col, coeff = c.matrix(which)
        terms = coeff.astype(np.int64) * zz[col]
        acc = np.concatenate([[0], np.cumsum(terms)])
        prods.append(acc[row_ptr[1:]] - acc[row_ptr[:-1]])
    return np.nonzero(prods[0] * prods[1] != prods[2])[0]


@pytest.mark.parametrize("msg_len", [64, 256, 4096])
def test_r1cs_satisfied_and_ciphertext_at_size(ctx, oracle, msg_len):
    rng = np.random.default_rng(msg_len)
    msg = rng.integers(0, 256, msg_len, dtype=np.uint8).tobytes()
    key = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
    c = zk.Circuit(msg_len)
    ct, z = ctx.witness_aes128_ecb(c, msg, key)
    assert ct == oracle.aes128_ecb(msg, key)  # src/aes.rs mirror
    assert _check_r1cs(c, z).size == 0
    # public inputs = 8 LSB-first bits per ciphertext byte (src/helpers/mod.rs:84-93)
    bits = np.unpackbits(np.frombuffer(ct, dtype=np.uint8), bitorder="little")
    assert z[0] == 1 and (z[1:1 + 8 * msg_len] == bits).all() and not z[1 + 8 * msg_len:c.info["num_instance"]].any()
    # flipping one wire must break at least one constraint
    z2 = z.copy()
    z2[c.info["num_instance"] + c.info["wit_block0"] + 1000] ^= 1
    assert _check_r1cs(c, z2).size > 0


def test_ragged_message_rejected(ctx):
    c = zk.Circuit(32)
    with pytest.raises(zk.ZkAesError):
        ctx.witness_aes128_ecb(c, b"\x00" * 16, KEY)
