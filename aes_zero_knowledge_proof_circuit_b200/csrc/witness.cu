// K1: AES-128-ECB witness generation on the device.
//
// Replaces the value side of the reference's circuit synthesis in encrypt() (src/lib.rs:66-98 -> :176-293): every
// Boolean wire the gadgets allocate -- AddRoundKey xors, the 8-level S-box select trees, the MixColumns xtime/xor
// chains, the key schedule (src/aes_circuit.rs:20-427, src/helpers/mod.rs:11-64) -- gets its value here, in exactly
// the variable order circuit.cpp derives for the constraint matrices.  The reference walks the ECB blocks one after
// the other on one CPU thread (src/lib.rs:194); here every block is one CTA.
//
// Layout: z is one byte per R1CS variable (all variables of this circuit are bits), in column order
//   [ one | 8 ciphertext bits per byte | 0-padding to a power of two ][ message bits | key bits | key schedule | block 0 | ... | dummy ones ]
// Each CTA evaluates the block's straight-line program level by level with the block's wires in shared memory
// (one byte per wire, 148 KB), then streams them to HBM with coalesced 16-byte stores.  HBM traffic is the floor:
// one byte written per witness wire (the field-element expansion happens in the consumers).
#include "witness.cuh"

namespace zk {

// zwit[i - gbias] holds global witness i (gbias != 0 when the globals live in shared memory: key schedule kernel)
__device__ __forceinline__ uint32_t fetch_ref(uint32_t ref, const uint8_t* __restrict__ local, const uint8_t* __restrict__ zwit,
                                              uint32_t msg_base, uint32_t gbias) {
    const uint32_t idx = ref & REF_INDEX_MASK;
    uint32_t v;
    switch (ref >> 30) {
        case 0: v = idx & 1; break;
        case 1: v = zwit[idx - gbias]; break;
        case 2: v = local[idx]; break;
        default: v = zwit[msg_base + idx]; break;
    }
    return v ^ ((ref >> 29) & 1);
}

__device__ __forceinline__ void run_levels(const WitInstr* __restrict__ prog, const uint32_t* __restrict__ lvl, int nlvl, uint8_t* local,
                                           uint32_t local_base, const uint8_t* __restrict__ zwit, uint32_t msg_base, uint32_t gbias) {
    for (int l = 0; l < nlvl; ++l) {
        const uint32_t lo = lvl[l], hi = lvl[l + 1];
        for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const WitInstr in = prog[i];
            const uint32_t a = fetch_ref(in.a, local, zwit, msg_base, gbias);
            const uint32_t b = fetch_ref(in.b, local, zwit, msg_base, gbias);
            uint32_t r;
            if (in.op == WOP_XOR)
                r = a ^ b;
            else if (in.op == WOP_AND)
                r = a & b;
            else
                r = a ? b : fetch_ref(in.c, local, zwit, msg_base, gbias);
            local[in.dst - local_base] = (uint8_t)r;
        }
        __syncthreads();
    }
}

// instance section (one, padding), message + key bits, dummy ones
__global__ void k_wit_inputs(const uint8_t* __restrict__ msg, const uint8_t* __restrict__ key, uint32_t msg_len, uint32_t num_instance,
                             uint32_t wit_key0, uint32_t num_witness_real, uint32_t num_witness, uint8_t* __restrict__ z) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t total = num_instance + num_witness;
    if (i >= total) return;
    if (i < num_instance) {
        if (i == 0) z[0] = 1;
        else if (i > 8 * msg_len) z[i] = 0;  // padding inputs (ciphertext bits are written by k_wit_blocks)
        return;
    }
    uint32_t w = i - num_instance;
    if (w < 8 * msg_len)
        z[i] = (msg[w >> 3] >> (w & 7)) & 1;  // UInt8::new_witness: LSB first
    else if (w < wit_key0 + 128)
        z[i] = (key[(w - wit_key0) >> 3] >> ((w - wit_key0) & 7)) & 1;
    else if (w >= num_witness_real)
        z[i] = 1;  // make_matrices_square's dummy variables
}

// key schedule: one CTA; local space = witnesses [wit_key0, wit_fixed_end)
__global__ void __launch_bounds__(1024) k_wit_fixed(const WitInstr* __restrict__ prog, const uint32_t* __restrict__ lvl, int nlvl,
                                                    uint32_t wit_key0, uint32_t wit_fixed0, uint32_t wit_fixed_end, uint8_t* __restrict__ zwit) {
    extern __shared__ uint8_t sm[];
    // operands are REF_GLOBAL with absolute witness indices: gbias redirects them into shared memory
    for (uint32_t i = threadIdx.x; i < wit_fixed0 - wit_key0; i += blockDim.x) sm[i] = zwit[wit_key0 + i];
    __syncthreads();
    run_levels(prog, lvl, nlvl, sm, wit_key0, sm, 0, wit_key0);
    for (uint32_t i = wit_fixed0 - wit_key0 + threadIdx.x; i < wit_fixed_end - wit_key0; i += blockDim.x) zwit[wit_key0 + i] = sm[i];
}

// one CTA per ECB block
__global__ void __launch_bounds__(1024) k_wit_blocks(const WitInstr* __restrict__ prog, const uint32_t* __restrict__ lvl, int nlvl,
                                                     const uint32_t* __restrict__ ct_refs, uint32_t wit_block0, uint32_t stride,
                                                     uint32_t num_instance, uint8_t* __restrict__ z, uint8_t* __restrict__ ct_out) {
    extern __shared__ uint8_t sm[];
    __shared__ uint8_t ct_bits[128];
    const uint32_t blk = blockIdx.x;
    const uint8_t* zwit = z + num_instance;
    const uint32_t msg_base = 128 * blk;
    run_levels(prog, lvl, nlvl, sm, 0, zwit, msg_base, 0);
    // wires -> HBM (block segments start at arbitrary byte offsets: align the bulk to 16 B)
    uint8_t* dst = z + num_instance + wit_block0 + (size_t)blk * stride;
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > stride) head = stride;
    for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = sm[i];
    uint32_t nvec = (stride - head) / 16;
    if ((head & 3) == 0) {
        // shared-memory side is 4-byte aligned: move 16 B per thread
        for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
            const uint32_t* s = reinterpret_cast<const uint32_t*>(sm + head + 16 * v);
            reinterpret_cast<uint4*>(dst + head)[v] = make_uint4(s[0], s[1], s[2], s[3]);
        }
    } else {
        for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
            const uint8_t* s = sm + head + 16 * v;
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w[k] = s[4 * k] | (s[4 * k + 1] << 8) | (s[4 * k + 2] << 16) | ((uint32_t)s[4 * k + 3] << 24);
            reinterpret_cast<uint4*>(dst + head)[v] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    for (uint32_t i = head + 16 * nvec + threadIdx.x; i < stride; i += blockDim.x) dst[i] = sm[i];
    // ciphertext: public-input bits (src/lib.rs:282-286) + packed bytes for the caller
    if (threadIdx.x < 128) {
        uint32_t v = fetch_ref(ct_refs[threadIdx.x], sm, zwit, msg_base, 0);
        ct_bits[threadIdx.x] = (uint8_t)v;
        z[1 + 128 * blk + threadIdx.x] = (uint8_t)v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        uint32_t b = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) b |= (uint32_t)ct_bits[8 * threadIdx.x + j] << j;
        ct_out[16 * blk + threadIdx.x] = (uint8_t)b;
    }
}

int witness_upload(zkaes_ctx* ctx, const AesCircuit& c, WitnessDev& w) {
    cudaStream_t st = ctx->stream;
    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes ? bytes : 1);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st);
    };
    ZK_CUDA(ctx, up((void**)&w.fixed, c.fixed_prog.instrs.data(), c.fixed_prog.instrs.size() * sizeof(WitInstr)));
    ZK_CUDA(ctx, up((void**)&w.fixed_lvl, c.fixed_prog.level_start.data(), c.fixed_prog.level_start.size() * 4));
    ZK_CUDA(ctx, up((void**)&w.block, c.block_prog.instrs.data(), c.block_prog.instrs.size() * sizeof(WitInstr)));
    ZK_CUDA(ctx, up((void**)&w.block_lvl, c.block_prog.level_start.data(), c.block_prog.level_start.size() * 4));
    ZK_CUDA(ctx, up((void**)&w.ct_refs, c.ct_refs.data(), c.ct_refs.size() * 4));
    w.fixed_nlvl = (int)c.fixed_prog.level_start.size() - 1;
    w.block_nlvl = (int)c.block_prog.level_start.size() - 1;
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    size_t fixed_sm = c.wit_fixed_end - c.wit_key0, block_sm = c.wit_block_stride;
    if (fixed_sm > 227 * 1024 || block_sm > 227 * 1024) return fail(ctx, ZK_ERR_UNSUPPORTED, "witness: program does not fit shared memory");
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_wit_fixed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fixed_sm));
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_wit_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)block_sm));
    return ZK_OK;
}

void witness_free(WitnessDev& w) {
    cudaFree(w.fixed);
    cudaFree(w.fixed_lvl);
    cudaFree(w.block);
    cudaFree(w.block_lvl);
    cudaFree(w.ct_refs);
    w = WitnessDev();
}

int witness_generate(zkaes_ctx* ctx, const AesCircuit& c, const WitnessDev& w, const uint8_t* d_msg, const uint8_t* d_key, uint8_t* d_z,
                     uint8_t* d_ct) {
    cudaStream_t st = ctx->stream;
    uint32_t total = c.num_instance + c.num_witness;
    k_wit_inputs<<<cdiv(total, 256), 256, 0, st>>>(d_msg, d_key, (uint32_t)c.msg_len, c.num_instance, c.wit_key0, c.num_witness_real,
                                                  c.num_witness, d_z);
    k_wit_fixed<<<1, 1024, c.wit_fixed_end - c.wit_key0, st>>>(w.fixed, w.fixed_lvl, w.fixed_nlvl, c.wit_key0, c.wit_fixed0, c.wit_fixed_end,
                                                              d_z + c.num_instance);
    k_wit_blocks<<<(unsigned)c.n_blocks, 1024, c.wit_block_stride, st>>>(w.block, w.block_lvl, w.block_nlvl, w.ct_refs, c.wit_block0,
                                                                       c.wit_block_stride, c.num_instance, d_z, d_ct);
    ctx->launches += 3;
    ZK_CUDA(ctx, cudaGetLastError());
    return ZK_OK;
}

}  // namespace zk
