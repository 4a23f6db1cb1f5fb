// Declarations for the AES witness kernels (witness.cu).
#pragma once
#include "circuit.h"
#include "common.cuh"

namespace zk {
struct WitnessDev {  // device copies of the two witness programs
    WitInstr* fixed = nullptr;
    uint32_t* fixed_lvl = nullptr;
    int fixed_nlvl = 0;
    WitInstr* block = nullptr;
    uint32_t* block_lvl = nullptr;
    int block_nlvl = 0;
    uint32_t* ct_refs = nullptr;
};
int witness_upload(zkaes_ctx* ctx, const AesCircuit& c, WitnessDev& w);
void witness_free(WitnessDev& w);
// d_msg (msg_len bytes), d_key (16 bytes): device.  d_z: num_instance + num_witness bytes.  d_ct: msg_len bytes.
int witness_generate(zkaes_ctx* ctx, const AesCircuit& c, const WitnessDev& w, const uint8_t* d_msg, const uint8_t* d_key, uint8_t* d_z,
                     uint8_t* d_ct);
}  // namespace zk
