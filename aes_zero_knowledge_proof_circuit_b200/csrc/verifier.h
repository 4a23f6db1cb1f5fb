// Host-side verifier (verifier.cpp): verifying-key blob and verify_encryption.  No device code.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ff.cuh"

struct zkaes_proof_fields;

namespace zk {

// Verifying key = everything `verify_encryption` (reference src/lib.rs:116-136) needs, as one byte string:
//   "ZKAESVK1" | |X| u64 | SRS max degree u64 | len u64 | IndexVerifierKey ToBytes (index info + 12 commitments) |
//   g, gamma_g (G1, ark-ff ToBytes, 97 B each) | h, beta_h (G2: x.c0 x.c1 y.c0 y.c1 canonical LE, 192 B each) |
//   count u64 | (degree bound u64, shift power tau^(D - bound) G, 97 B) ...
// i.e. ark-marlin's IndexVerifierKey {index_info, index_comms, verifier_key} with ark-poly-commit's marlin_pc::VerifierKey
// {vk: {g, gamma_g, h, beta_h}, degree_bounds_and_shift_powers}.  tau, gamma: the test SRS trapdoors (Montgomery form).
std::vector<uint8_t> build_verifying_key(const std::vector<uint8_t>& index_vk, uint64_t x_padded, uint64_t max_degree, const Fp<Fr377Params>& tau,
                                         const Fp<Fr377Params>& gamma, const std::vector<uint64_t>& degree_bounds);

// Returns 0 and sets *accepted to 0/1, or -1 (with *err) when the key or the proof cannot be parsed.
int verify_encryption_host(const uint8_t* vk, size_t vk_len, const uint8_t* proof, size_t proof_len, const uint8_t* ciphertext, size_t ct_len,
                           int* accepted, std::string* err);

// ark_marlin::Proof bytes <-> plain fields (include/zkaes_b200.h: zkaes_proof_fields); -1 with *err on malformed input
int proof_deserialize_host(const uint8_t* proof, size_t len, struct zkaes_proof_fields* out, std::string* err);
int proof_serialize_host(const struct zkaes_proof_fields* in, std::vector<uint8_t>& out, std::string* err);

// e(a G1, b G2) as 12 x 48 canonical LE bytes (c[0].c0, c[0].c1, c[1].c0, ...): test hook against oracle/pairing_ref.py
void pairing_selftest(const uint8_t a32[32], const uint8_t b32[32], uint8_t out576[576]);

}  // namespace zk
