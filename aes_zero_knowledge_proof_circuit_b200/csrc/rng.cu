// Device-side bulk field sampling: `count` consecutive ark-ff `Fr::rand` draws from a ChaCha20 stream.
//
// Stands in for `DensePolynomial::rand(3|H| + 2 zk - 3, zk_rng)` in ark-marlin 0.3.0's first prover round (mask
// polynomial; reached from src/lib.rs:111) when the zk rng is rand_chacha's ChaCha20Rng: the draws are sequential in
// the reference, but ChaCha is counter based, so candidate j of the stream can be produced by any thread; rejection
// sampling keeps stream order through a prefix sum over the accept flags.  Output is bit-identical to sequential draws.
#include "msm.cuh"
#include "polyops.cuh"
#include "transcript.h"

namespace zk {
namespace {

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
#define ZK_QR(a, b, c, d)                      \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);

struct ChaChaKey {
    uint32_t k[8];
};
// one thread = one ChaCha20 block = 8 u64 = two Fr candidates
__global__ void __launch_bounds__(256) k_chacha_candidates(ChaChaKey key, uint64_t block0, size_t nblocks, uint32_t top_mask, FrS mod,
                                                           uint32_t* __restrict__ cand, uint32_t* __restrict__ flags) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    uint64_t ctr = block0 + b;
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3], key.k[4], key.k[5], key.k[6],
                      key.k[7], (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = s[i];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        ZK_QR(0, 4, 8, 12) ZK_QR(1, 5, 9, 13) ZK_QR(2, 6, 10, 14) ZK_QR(3, 7, 11, 15)
        ZK_QR(0, 5, 10, 15) ZK_QR(1, 6, 11, 12) ZK_QR(2, 7, 8, 13) ZK_QR(3, 4, 9, 14)
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] += s[i];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t* w = x + 8 * c;
        w[7] &= top_mask;
        bool lt = false, decided = false;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            if (!decided && w[i] != mod.v[i]) {
                lt = w[i] < mod.v[i];
                decided = true;
            }
        }
        uint4* o = reinterpret_cast<uint4*>(cand + (2 * b + c) * 8);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
        flags[2 * b + c] = lt ? 1u : 0u;
    }
}
// accepted candidate c goes to out[have + rank(c)] while that is < count; the candidate that completes the request
// reports how many candidates were consumed
__global__ void __launch_bounds__(256) k_compact(const uint32_t* __restrict__ cand, const uint32_t* __restrict__ flags,
                                                 const uint32_t* __restrict__ rank, size_t ncand, size_t have, size_t count, FrS* __restrict__ out,
                                                 unsigned long long* __restrict__ used) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand || !flags[c]) return;
    size_t pos = have + rank[c];
    if (pos >= count) return;
    const uint4* i4 = reinterpret_cast<const uint4*>(cand + c * 8);
    uint4* o = reinterpret_cast<uint4*>(out + pos);
    o[0] = i4[0];
    o[1] = i4[1];
    if (pos == count - 1) *used = c + 1;
}

}  // namespace

// Fills out[0..count) (device) with the next `count` Fr::rand draws of `rng` and advances it.
int fr_rand_device(zkaes_ctx* ctx, ChaCha20Rng& rng, FrS* out, size_t count) {
    cudaStream_t st = ctx->stream;
    constexpr int shave = 256 - Fr377Params::BITS;
    const uint32_t top_mask = 0xffffffffu >> shave;
    FrS mod;
    for (int i = 0; i < 8; ++i) mod.v[i] = Fr377Params::MOD(i);
    ChaChaKey key;
    for (int i = 0; i < 8; ++i) key.k[i] = rng.key[i];
    size_t have = 0;
    // candidates are 4 u64; the kernels work on whole ChaCha blocks (8 u64): align with host draws first
    std::vector<FrS> head;
    while (have < count && (rng.pos & 7)) {
        uint64_t w[4];
        fr_rand_raw<Fr377Params>(rng, w);
        FrS f;
        memcpy(f.v, w, 32);
        head.push_back(f);
        ++have;
    }
    if (!head.empty()) ZK_CUDA(ctx, cudaMemcpyAsync(out, head.data(), sizeof(FrS) * head.size(), cudaMemcpyHostToDevice, st));
    const size_t CHUNK = (size_t)1 << 24;  // candidates per pass
    DevBuf cand, flags, rank, used;
    ZK_CUDA(ctx, used.alloc(sizeof(unsigned long long) + 8, st));
    while (have < count) {
        size_t need = count - have;
        size_t ncand = need + need * 3 / 4 + 4096;  // acceptance is r / 2^253 ~ 0.583
        if (ncand > CHUNK) ncand = CHUNK;
        ncand = (ncand + 1) & ~(size_t)1;
        if (!cand.p) {
            size_t cap = ncand;  // the first pass is the largest
            ZK_CUDA(ctx, cand.alloc(32 * cap, st));
            ZK_CUDA(ctx, flags.alloc(4 * cap, st));
            ZK_CUDA(ctx, rank.alloc(4 * cap, st));
        }
        const uint64_t block0 = rng.pos >> 3;
        k_chacha_candidates<<<cdiv(ncand / 2, 256), 256, 0, st>>>(key, block0, ncand / 2, top_mask, mod, cand.as<uint32_t>(), flags.as<uint32_t>());
        ctx->launches++;
        ZK_TRY(exclusive_scan_u32(ctx, flags.as<uint32_t>(), rank.as<uint32_t>(), (uint32_t)ncand));
        ZK_CUDA(ctx, cudaMemsetAsync(used.p, 0, sizeof(unsigned long long), st));
        k_compact<<<cdiv(ncand, 256), 256, 0, st>>>(cand.as<uint32_t>(), flags.as<uint32_t>(), rank.as<uint32_t>(), ncand, have, count, out,
                                                   used.as<unsigned long long>());
        ctx->launches++;
        ZK_CUDA(ctx, cudaGetLastError());
        uint32_t last_rank = 0, last_flag = 0;
        unsigned long long used_h = 0;
        ZK_CUDA(ctx, cudaMemcpyAsync(&last_rank, rank.as<uint32_t>() + (ncand - 1), 4, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(ctx, cudaMemcpyAsync(&last_flag, flags.as<uint32_t>() + (ncand - 1), 4, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(ctx, cudaMemcpyAsync(&used_h, used.p, sizeof(used_h), cudaMemcpyDeviceToHost, st));
        ZK_CUDA(ctx, cudaStreamSynchronize(st));
        size_t accepted = (size_t)last_rank + last_flag;
        if (have + accepted >= count) {
            if (used_h == 0) return fail(ctx, ZK_ERR_STATE, "fr_rand_device: compaction did not report its end");
            rng.pos += 4 * used_h;
            have = count;
        } else {
            rng.pos += 4 * ncand;
            have += accepted;
        }
        rng.cur_block = ~0ull;
    }
    return ZK_OK;
}

}  // namespace zk
