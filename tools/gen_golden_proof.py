#!/usr/bin/env python3
"""Generate tests/golden/marlin_proof_16B.json with the CPU oracle (oracle/marlin_oracle.py): index + proof for the
reference's own end-to-end case (tests/integration_tests.rs:313-337: FIPS-197 key / plaintext) under fixed seeds.
Takes ~2 minutes on 8 cores.  The GPU parity test compares the product's proof bytes with this fixture."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import marlin_oracle as mo  # noqa: E402
from oracle import r1cs_model as model  # noqa: E402

KEY = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
MSG = bytes.fromhex("3243f6a8885a308d313198a2e0370734")
TAU_SEED, GAMMA_SEED, ZK_SEED = bytes(range(32)), bytes(range(1, 33)), bytes([7] * 32)


# second fixture: FIPS-197 Appendix C.1 (key 00..0f, plaintext 00 11 .. ff -> 69c4e0d8 6a7b0430 d8cdb780 70b4c55a), another zk seed, same SRS
VARIANTS = {
    "marlin_proof_16B.json": (MSG, KEY, ZK_SEED),
    "marlin_proof_16B_fips_c1.json": (bytes.fromhex("00112233445566778899aabbccddeeff"), bytes(range(16)), bytes([0xA5] * 32)),
    # third fixture: TWO ECB blocks (|H| = 2^19, |K| = 2^20, |X| = 512): the key schedule is shared, the X-subdomain interleaving and the
    # padding of the instance differ from the one-block case
    "marlin_proof_32B.json": (MSG + bytes.fromhex("00112233445566778899aabbccddeeff"), KEY, bytes([0x3C] * 32)),
}


def main():
    for name, (msg, key, zk_seed) in VARIANTS.items():
        if len(sys.argv) > 1 and sys.argv[1] != name:
            continue
        generate(name, msg, key, zk_seed)


def generate(name, MSG, KEY, ZK_SEED):
    t0 = time.time()
    cs, ct = model.synthesize(MSG, KEY)
    A, B, C = cs.matrices()
    r1cs = mo.R1CS(A, B, C, len(cs.inst_vals), len(cs.wit_vals))
    idx0 = mo.index_r1cs(r1cs)
    srs = mo.SRS.generate(idx0.max_degree, TAU_SEED, GAMMA_SEED)
    idx = mo.index_r1cs(r1cs, srs)
    proof, pb = mo.prove(idx, srs, r1cs, cs.inst_vals, cs.wit_vals, ZK_SEED)
    pub = [(b >> i) & 1 for b in ct for i in range(8)]
    assert mo.verify(idx, srs, pub, proof), "oracle verifier rejected the oracle proof"
    bad = list(pub)
    bad[9] ^= 1
    assert not mo.verify(idx, srs, bad, proof), "oracle verifier accepted a wrong ciphertext"
    out = {
        "generator": "tools/gen_golden_proof.py (oracle/marlin_oracle.py; CPU restatement -- NOT output of the Rust reference)",
        "message": MSG.hex(), "key": KEY.hex(), "ciphertext": ct.hex(),
        "tau_seed": TAU_SEED.hex(), "gamma_seed": GAMMA_SEED.hex(), "zk_seed": ZK_SEED.hex(),
        "h": idx.domain_h.size, "k": idx.domain_k.size, "x": idx.domain_x.size, "max_degree": idx.max_degree,
        "vk_sha256": hashlib.sha256(idx.vk_bytes()).hexdigest(), "vk_len": len(idx.vk_bytes()), "index_vk": idx.vk_bytes().hex(),
        "verifying_key": mo.verifying_key_bytes(idx.vk_bytes(), idx.domain_x.size, idx.max_degree, idx.domain_h.size, idx.domain_k.size,
                                                TAU_SEED, GAMMA_SEED).hex(),
        "proof": pb.hex(),
    }
    path = os.path.join(ROOT, "tests", "golden", name)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, "in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
