// Host-side synthesis of the reference's AES-128-ECB R1CS: constraint matrices (CSR, small integer coefficients) and the
// straight-line Boolean program that the GPU witness kernel (witness.cu) evaluates per ECB block.
//
// Mirrors, for this one circuit, what the reference obtains by running its gadget code against arkworks' constraint
// system:  src/lib.rs:60-98 (message/key allocation), src/lib.rs:176-293 (encrypt_and_generate_constraints),
// src/aes_circuit.rs:20-427 (derive_keys, add_round_key, substitute_byte(s), shift_rows, mix_columns),
// src/helpers/mod.rs:11-64 (add / multiply), with ark-r1cs-std 0.3.1 Boolean/UInt8 expansion rules and ark-relations
// 0.3.0 variable numbering.  Structure only: no witness VALUES are computed on the host (they come from the GPU kernel).
#pragma once
#include <cstdint>
#include <vector>

namespace zk {

// One matrix in CSR: row r owns entries [row_ptr[r], row_ptr[r+1]); columns ascending within a row.
struct CsrMatrix {
    std::vector<uint32_t> row_ptr;  // num_rows + 1
    std::vector<uint32_t> col;
    std::vector<int8_t> coeff;
    size_t nnz() const { return col.size(); }
};

// Witness program.  Every witness variable that is not a primary input (message / key bit) is the output of one op.
//   operand reference: bits 31..30 = space, bit 29 = negate, bits 28..0 = index
//   space 0: constant (index = 0/1)   1: global witness (absolute witness index)
//         2: block-local witness (index relative to the block's first witness)   3: message bit of this block (0..127)
enum WitOp : uint8_t { WOP_XOR = 0, WOP_AND = 1, WOP_SEL = 2 };
struct WitInstr {
    uint32_t dst;      // witness index: absolute (fixed part) or block-relative (block program)
    uint32_t a, b, c;  // XOR: a ^ b; AND: a & b; SEL: a ? b : c   (negations folded into the references)
    uint32_t op;       // WitOp
};
struct WitProgram {
    std::vector<WitInstr> instrs;       // sorted by level
    std::vector<uint32_t> level_start;  // instrs of level l: [level_start[l], level_start[l+1])
};
static constexpr uint32_t REF_CONST = 0u << 30, REF_GLOBAL = 1u << 30, REF_LOCAL = 2u << 30, REF_MSG = 3u << 30, REF_NEG = 1u << 29;
static constexpr uint32_t REF_INDEX_MASK = (1u << 29) - 1;

struct AesCircuit {
    size_t msg_len = 0, n_blocks = 0;
    // ark-relations numbering after Marlin's padding (pad_input_for_indexer_and_prover + make_matrices_square):
    uint32_t num_instance = 0;         // padded to a power of two; [0] is the constant one, then 8 bits per ciphertext byte
    uint32_t num_instance_used = 0;    // 1 + 8 * msg_len
    uint32_t num_witness = 0;          // including dummy padding witnesses (value one)
    uint32_t num_witness_real = 0;
    uint32_t num_constraints = 0;      // including dummy 0*0=0 rows; == num_instance + num_witness
    CsrMatrix a, b, c;
    // witness layout: [message bits 8*len][key bits 128][key schedule][block 0][block 1]...[dummy ones]
    uint32_t wit_key0 = 0, wit_fixed0 = 0, wit_fixed_end = 0, wit_block0 = 0, wit_block_stride = 0;
    WitProgram fixed_prog;   // key schedule; dst/operands absolute (REF_GLOBAL)
    WitProgram block_prog;   // one ECB block; dst block-relative
    // ciphertext bit j (0..127, byte-major, LSB first) of a block = reference into the block program's spaces
    std::vector<uint32_t> ct_refs;
};

// Throws std::runtime_error on a malformed request (msg_len == 0 or not a multiple of 16).
void build_aes_circuit(size_t msg_len, AesCircuit& out);

}  // namespace zk
