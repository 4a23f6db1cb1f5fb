// Host-side verifier (verifier.cpp): verifying-key blob and verify_encryption.  No device code.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ec.cuh"

struct zkaes_proof_fields;

namespace zk {

// Verifying key = the VerifyingKey half of `synthesize_keys` (reference src/lib.rs:138,173): the ark-serialize 0.3.0
// CanonicalSerialize bytes of ark_marlin::IndexVerifierKey<Fr, MarlinKZG10<Bls12_377, DensePolynomial<Fr>>> -- index_info (four u64),
// the 12 index commitments (compressed G1 + an empty shifted_comm), marlin_pc::VerifierKey {kzg10 vk: g, gamma_g (G1), h, beta_h
// (compressed G2), degree_bounds_and_shift_powers, max_degree, supported_degree}.  Layout in verifier.cpp.
// tau, gamma: the test-SRS trapdoors (Montgomery form); index_comms: 12 points.
std::vector<uint8_t> build_verifying_key(uint64_t num_variables, uint64_t num_constraints, uint64_t num_non_zero, uint64_t x_padded,
                                         const Affine<G1_377Params>* index_comms, uint64_t max_degree, const Fp<Fr377Params>& tau,
                                         const Fp<Fr377Params>& gamma, std::vector<uint64_t> degree_bounds);

// Returns 0 and sets *accepted to 0/1, or -1 (with *err) when the key or the proof cannot be parsed.
int verify_encryption_host(const uint8_t* vk, size_t vk_len, const uint8_t* proof, size_t proof_len, const uint8_t* ciphertext, size_t ct_len,
                           int* accepted, std::string* err);

// ark_marlin::Proof bytes <-> plain fields (include/zkaes_b200.h: zkaes_proof_fields); -1 with *err on malformed input
int proof_deserialize_host(const uint8_t* proof, size_t len, struct zkaes_proof_fields* out, std::string* err);
int proof_serialize_host(const struct zkaes_proof_fields* in, std::vector<uint8_t>& out, std::string* err);

// e(a G1, b G2) as 12 x 48 canonical LE bytes (c[0].c0, c[0].c1, c[1].c0, ...): test hook against tools/pairing_model.py
void pairing_selftest(const uint8_t a32[32], const uint8_t b32[32], uint8_t out576[576]);

}  // namespace zk
