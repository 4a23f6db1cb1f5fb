"""B200-native prover path for the AES-128 R1CS circuit of lambdaclass/AES_zero_knowledge_proof_circuit.

Host-side mirror of the reference's public surface (src/lib.rs) over the C ABI in include/zkaes_b200.h.
"""
from ._native import (CURVE_BLS12_377, CURVE_BLS12_381, Circuit, Context, ProofFields, ProvingKey, ZkAesError, comm_unique_id, coset_plan,  # noqa: F401
                      deserialize_proof, lib, pairing_selftest, serialize_proof, shard_range, verify_encryption)
