// Declarations for the device-resident Marlin prover (prover.cu).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "circuit.h"
#include "common.cuh"
#include "ec.cuh"
#include "polyops.cuh"
#include "witness.cuh"

namespace zk {
struct zkaes_pk_impl;
constexpr int ZK_PK_INFO_WORDS = 12;
int pk_synthesize(zkaes_ctx* ctx, size_t msg_len, const uint8_t tau_seed[32], const uint8_t gamma_seed[32], zkaes_pk_impl** out);
void pk_free(zkaes_pk_impl* pk);
// key files: flags bit 0 = include this rank's SRS share, bit 1 = include the 12 index polynomials (prover.cu, "key files")
int pk_save(zkaes_ctx* ctx, const zkaes_pk_impl* pk, const char* path, int flags);
int pk_load(zkaes_ctx* ctx, const char* path, zkaes_pk_impl** out);
const std::vector<uint8_t>& pk_vk_bytes(const zkaes_pk_impl* pk);
const std::vector<uint8_t>& pk_verifying_key(const zkaes_pk_impl* pk);
// 0 msg_len, 1 num_constraints, 2 num_variables, 3-5 nnz(A,B,C), 6 |H|, 7 |K|, 8 |X|, 9 SRS max degree, 10 instance variables used, 11 Lagrange points per basis
void pk_info(const zkaes_pk_impl* pk, uint64_t info[ZK_PK_INFO_WORDS]);
int pk_encrypt(zkaes_ctx* ctx, const zkaes_pk_impl* pk, const uint8_t* msg, size_t msg_len, const uint8_t key[16], const uint8_t zk_seed[32],
               uint8_t* ct_out, std::vector<uint8_t>& proof);
}  // namespace zk
