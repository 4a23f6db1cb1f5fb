"""CPU checks of the oracle's Marlin restatement (oracle/marlin_oracle.py, oracle/fs_rng.py) on a toy R1CS: the prover's
output is accepted by the algebraic verifier, every kind of tampering is rejected, and the wire format round-trips.
(The reference's own tests assert exactly accept / reject at this boundary: tests/integration_tests.rs:330-371.)"""
import random

import numpy as np
import pytest

from oracle import marlin_oracle as mo
from oracle.cpu import FR
from oracle.fs_rng import ChaCha20Rng, FiatShamirRng, chacha20_blocks, fr_rand, fr_rand_many

TAU, GAMMA = bytes(range(32)), bytes(range(1, 33))


def toy_r1cs(n_pub=3, n_cons=20, seed=1):
    rnd = random.Random(seed)
    inst = [1] + [rnd.randrange(0, 5) for _ in range(n_pub)]
    wit = [rnd.randrange(0, 3) for _ in range(4)]
    a, b, c = [], [], []
    for _ in range(n_cons):
        nv = len(inst) + len(wit)

        def row():
            cols = rnd.sample(range(nv), rnd.randrange(1, 4))
            return sorted((cc, rnd.choice([1, -1, 2])) for cc in cols)

        ra, rb = row(), row()
        z = inst + wit
        va = sum(v * z[cc] for cc, v in ra)
        vb = sum(v * z[cc] for cc, v in rb)
        wit.append(va * vb % mo.P)
        a.append(ra)
        b.append(rb)
        c.append([(nv, 1)])
    return mo.R1CS(a, b, c, len(inst), len(wit)), inst, wit


@pytest.fixture(scope="module")
def toy():
    r1cs, inst, wit = toy_r1cs()
    idx0 = mo.index_r1cs(r1cs)
    srs = mo.SRS.generate(idx0.max_degree, TAU, GAMMA)
    idx = mo.index_r1cs(r1cs, srs)
    proof, pb = mo.prove(idx, srs, r1cs, inst, wit, bytes([7] * 32))
    return r1cs, inst, wit, idx, srs, proof, pb


def test_chacha20_rfc7539_keystream():
    b = chacha20_blocks(bytes(32), 0, 2)
    assert b[0].astype("<u4").tobytes().hex().startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")
    assert b[1].astype("<u4").tobytes().hex().startswith("9f07e7be5551387a98ba977c732d080dcb0f29a048e3656912c6533e32ee7aed")


def test_fr_rand_vectorised_equals_sequential():
    r1, r2 = ChaCha20Rng(bytes(range(32))), ChaCha20Rng(bytes(range(32)))
    a = fr_rand_many(r1, FR[377], 50)
    b = np.stack([fr_rand(r2, FR[377]) for _ in range(50)])
    assert (a == b).all() and r1.pos == r2.pos
    vals = [sum(int(x) << (64 * k) for k, x in enumerate(row)) for row in a]
    assert all(v < FR[377] for v in vals)


def test_fiat_shamir_absorb_reseeds():
    f = FiatShamirRng(b"seed")
    s0 = f.seed
    x0 = f.rng.next_u64()
    f.absorb(b"data")
    assert f.seed != s0 and f.rng.pos == 0 and f.rng.next_u64() != x0


def test_accepts_honest_proof(toy):
    _, inst, _, idx, srs, proof, pb = toy
    assert mo.verify(idx, srs, inst[1:], proof)
    assert len(pb) == 951  # 9 commitments (2 with a shifted part), 7 evaluations, 2 opening proofs


def test_rejects_wrong_statement_and_tampering(toy):
    _, inst, _, idx, srs, proof, _ = toy
    bad = list(inst[1:])
    bad[0] += 1
    assert not mo.verify(idx, srs, bad, proof)
    for i in range(len(proof["evaluations"])):
        ev = list(proof["evaluations"])
        ev[i] = (ev[i] + 1) % mo.P
        assert not mo.verify(idx, srs, inst[1:], {**proof, "evaluations": ev})
    for which in (0, 1):
        pc = [dict(p) for p in proof["pc_proof"]]
        pc[which]["w"] = mo.g1_add(pc[which]["w"], srs.powers_of_g[0])
        assert not mo.verify(idx, srs, inst[1:], {**proof, "pc_proof": pc})
    comms = [list(r) for r in proof["commitments"]]
    comms[1][0] = (mo.g1_add(comms[1][0][0], srs.powers_of_g[0]), None)
    assert not mo.verify(idx, srs, inst[1:], {**proof, "commitments": comms})


def test_unsatisfied_witness_cannot_be_proved(toy):
    r1cs, inst, wit, idx, srs, _, _ = toy
    w2 = list(wit)
    w2[-1] = (w2[-1] + 1) % mo.P
    with pytest.raises(AssertionError):
        mo.prove(idx, srs, r1cs, inst, w2, bytes([7] * 32))


def test_wire_format_round_trip_and_light_verifier(toy):
    _, inst, _, idx, srs, proof, pb = toy
    p2 = mo.deserialize_proof(pb)
    assert mo.serialize_proof(p2) == pb
    idx2 = mo.index_from_vk_bytes(idx.vk_bytes(), idx.domain_x.size)
    srs2 = mo.SparseSRS(idx.max_degree, TAU, GAMMA)
    assert mo.verify(idx2, srs2, inst[1:], p2)
    bad = list(inst[1:])
    bad[1] += 1
    assert not mo.verify(idx2, srs2, bad, p2)


def test_proof_is_deterministic_in_the_seed(toy):
    r1cs, inst, wit, idx, srs, _, pb = toy
    _, pb2 = mo.prove(idx, srs, r1cs, inst, wit, bytes([7] * 32))
    _, pb3 = mo.prove(idx, srs, r1cs, inst, wit, bytes([8] * 32))
    assert pb2 == pb and pb3 != pb
