"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/zkaes_b200.h declares,
fails loudly without a GPU, and its host-side arithmetic (used for the O(W*c) MSM fold) matches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import aes_zero_knowledge_proof_circuit_b200 as zk
from oracle.cpu import FQ, FR, ints_to_limbs, rand_fr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "zkaes_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zkaes_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(zk._native.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/zkaes_b200.h but not exported"
    # and the python binding covers all of them
    assert set(syms) == set(zk.lib()._zk_symbols)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zk.ZkAesError):
        zk.Context(0)


def test_null_context_is_an_error_not_a_crash():
    lib = zk.lib()
    assert lib.zkaes_ctx_sync(None) < 0
    assert lib.zkaes_msm_g1(None, 377, None, None, 0, None) < 0
    assert lib.zkaes_last_error(None) == b"null context"


@pytest.mark.parametrize("curve", [377, 381])
@pytest.mark.parametrize("field", [0, 1])
def test_host_field_arithmetic_matches_oracle(oracle, curve, field):
    import random

    p = (FR if field == 0 else FQ)[curve]
    nl = 4 if field == 0 else 6
    rnd = random.Random(curve * 2 + field)
    vals = [0, 1, p - 1, p - 2] + [rnd.randrange(p) for _ in range(300)]
    a = ints_to_limbs(vals, nl)
    b = ints_to_limbs(list(reversed(vals)), nl)
    for op in (0, 1, 2, 3):
        out = np.zeros_like(a)
        rc = zk.lib().zkaes_selftest_host_field(curve, field, op, a.ctypes.data, b.ctypes.data, out.ctypes.data, a.shape[0])
        assert rc == 0
        assert (out == oracle.field_op(curve, field, op, a, b)).all(), (curve, field, op)


@pytest.mark.parametrize("curve", [377, 381])
def test_host_g1_formulas_match_oracle(oracle, curve):
    rng = np.random.default_rng(curve)
    n = 48
    ka, kb = rand_fr(rng, curve, n), rand_fr(rng, curve, n)
    kb[0] = ka[0]  # P + P  -> doubling branch
    kb[1] = ints_to_limbs([FR[curve] - sum(int(x) << (64 * i) for i, x in enumerate(ka[1]))], 4)[0]  # P + (-P) -> infinity
    A, B = oracle.g1_mul_gen(curve, ka), oracle.g1_mul_gen(curve, kb)
    B[2] = 0  # + infinity
    A[3] = 0  # infinity + Q
    exp = np.stack([oracle.g1_add(curve, A[i], B[i]) for i in range(n)])
    assert (exp[1] == 0).all()
    for op in (0, 1):
        out = np.zeros_like(A)
        assert zk.lib().zkaes_selftest_host_g1(curve, op, A.ctypes.data, B.ctypes.data, out.ctypes.data, n) == 0
        assert (out == exp).all(), (curve, op)
    out = np.zeros_like(A)
    assert zk.lib().zkaes_selftest_host_g1(curve, 2, A.ctypes.data, A.ctypes.data, out.ctypes.data, n) == 0
    assert (out == np.stack([oracle.g1_add(curve, A[i], A[i]) for i in range(n)])).all()


def test_header_is_plain_c_and_verifier_links_from_c(tmp_path):
    """include/zkaes_b200.h compiles as C99 and a C program verifies the golden proof through the shared library -- the view
    a cgo / Rust -sys / JNI binding has of the boundary."""
    import json
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "aes_zero_knowledge_proof_circuit_b200")
    exe = str(tmp_path / "c_abi_verify")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", os.path.join(root, "tests", "c_abi_verify.c"), "-o", exe,
                           "-L" + libdir, "-lzkaes_b200", "-Wl,-rpath," + libdir])
    with open(os.path.join(root, "tests", "golden", "marlin_proof_16B.json")) as f:
        gold = json.load(f)
    ct = bytes.fromhex(gold["ciphertext"])
    for name, data in (("vk", bytes.fromhex(gold["verifying_key"])), ("proof", bytes.fromhex(gold["proof"])), ("ct", ct),
                       ("ct_bad", bytes([ct[0] ^ 1]) + ct[1:])):
        (tmp_path / name).write_bytes(data)
    run = lambda c: subprocess.run([exe, str(tmp_path / "vk"), str(tmp_path / "proof"), str(tmp_path / c)], capture_output=True, text=True)
    ok, bad = run("ct"), run("ct_bad")
    assert (ok.returncode, ok.stdout.strip()) == (0, "accepted"), ok.stderr
    assert (bad.returncode, bad.stdout.strip()) == (1, "rejected"), bad.stderr


def test_rust_sys_crate_declares_exactly_the_header_exports():
    """rust/zk-aes-b200-sys/src/lib.rs (the -sys shim a Rust caller binds) declares one `pub fn` per export of include/zkaes_b200.h, no
    more and no fewer, and the shared library exports every one of them.  (No Rust toolchain in this image: the crate is source only.)"""
    import re

    hdr = open(os.path.join(ROOT, "include", "zkaes_b200.h")).read()
    exported = set(re.findall(r"^(?:int|void|size_t|uint64_t|const char\*|void\*) ?\*? ?(zkaes_[a-z0-9_]+)\(", hdr, re.M))
    rust = open(os.path.join(ROOT, "rust", "zk-aes-b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (zkaes_[a-z0-9_]+)\(", rust))
    assert declared == exported, (sorted(declared - exported), sorted(exported - declared))
    assert exported == set(zk.lib()._zk_symbols), (sorted(exported - set(zk.lib()._zk_symbols)), sorted(set(zk.lib()._zk_symbols) - exported))
    # the facade keeps the reference's three signatures (src/lib.rs:60-64, 116-120, 138)
    facade = open(os.path.join(ROOT, "rust", "zk-aes-b200", "src", "lib.rs")).read()
    for sig in ("pub fn synthesize_keys(plaintext_length: usize) -> Result<(ProvingKey, VerifyingKey)>",
                "pub fn encrypt(message: &[u8], secret_key: &[u8; 16], proving_key: ProvingKey) -> Result<MarlinProof>",
                "pub fn verify_encryption(verifying_key: VerifyingKey, proof: &MarlinProof, ciphertext: &[u8]) -> Result<bool>"):
        assert sig in facade, sig
