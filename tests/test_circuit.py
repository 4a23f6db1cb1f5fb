"""Shape of the product's host-side AES R1CS (csrc/circuit.cpp, through the C ABI) against the oracle's gadget model
(oracle/r1cs_model.py): same variable numbering, same rows of A / B / C, entry by entry.  CPU only."""
import numpy as np
import pytest

import aes_zero_knowledge_proof_circuit_b200 as zk
from oracle import marlin_oracle as mo
from oracle import r1cs_model as model

KEY = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")


def model_r1cs(msg):
    cs, ct = model.synthesize(msg, KEY)
    A, B, C = cs.matrices()
    r = mo.R1CS(A, B, C, len(cs.inst_vals), len(cs.wit_vals))
    padded, inst, wit = mo.pad_r1cs(r, cs.inst_vals, cs.wit_vals)
    return padded, inst, wit, ct


def csr_rows(row_ptr, col, coeff):
    return [list(zip(col[row_ptr[r]:row_ptr[r + 1]].tolist(), coeff[row_ptr[r]:row_ptr[r + 1]].tolist())) for r in range(len(row_ptr) - 1)]


@pytest.mark.parametrize("n_blocks", [1, 2, 3])  # 3 blocks: 385 instance variables padded to 512 (not a power of two of blocks)
def test_matrices_match_gadget_model(n_blocks):
    msg = bytes((i * 131 + 7) & 0xFF for i in range(16 * n_blocks))
    padded, inst, wit, _ = model_r1cs(msg)
    c = zk.Circuit(len(msg))
    info = c.info
    assert info["num_instance"] == padded.num_instance
    assert info["num_witness"] == padded.num_witness
    assert info["num_constraints"] == len(padded.a) == info["num_instance"] + info["num_witness"]
    assert info["num_instance_used"] == 1 + 8 * len(msg)  # the reference's 513 at 64 bytes (src/lib.rs:141)
    for which, ref in enumerate((padded.a, padded.b, padded.c)):
        rows = csr_rows(*c.matrix(which))
        assert len(rows) == len(ref)
        for r, (got, exp) in enumerate(zip(rows, ref)):
            assert got == [(cc, v) for cc, v in exp], (which, r, got, exp)


def test_known_sizes():
    c = zk.Circuit(16)
    assert c.info["num_constraints"] == 185040 and c.info["num_witness_real"] == 184784
    assert (c.info["nnz_a"], c.info["nnz_b"], c.info["nnz_c"]) == (200215, 337744, 344043)  # with AllocatedBool::or for (Is, Is) (round 1, NOR lowering everywhere: 200229, 343012, 339683)
    assert c.info["wit_block_stride"] == c.info["block_instrs"]  # every block witness is the output of one program op
    c4 = zk.Circuit(64)
    assert c4.info["num_instance_used"] == 513 and c4.info["num_instance"] == 1024


@pytest.mark.parametrize("bad", [0, 15, 17, 100])
def test_ragged_lengths_rejected(bad):
    # src/aes_circuit.rs:218-221: add_round_key ensures 16-byte blocks
    with pytest.raises(zk.ZkAesError):
        zk.Circuit(bad)
