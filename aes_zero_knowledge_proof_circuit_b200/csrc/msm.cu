// K3: G1 multi-scalar multiplication (bucket method) for sm_100a.
//
// Replaces ark-ec 0.3.0 `VariableBaseMSM::multi_scalar_mul(&[G1Affine], &[BigInteger256]) -> G1Projective`
// (reference Cargo.lock:118-120; reached from src/lib.rs:111 via ark-poly-commit's `commit`/`open`).
// Same inputs (affine bases, canonical 256-bit scalars < r), same mathematical result; the schedule is GPU-first
// (per-thread bodies and their rationale: msm_core.cuh):
//
//   1. k_msm_digits_all: s -> min(s, r-s), all W signed c-bit digits in one carry walk, window-major digit array   [HBM: 32 B/term read, 4 W written]
//      k_msm_sort_pass : histogram of (window, |digit|) keys over the digit array, one grid row per window
//   2. k_scan_*        : exclusive scan of the histogram -> bucket offsets (nb + 1 entries)
//   3. k_msm_sort_pass : counting-sort scatter of point indices (sign in bit 31) by bucket key
//   4. k_msm_accumulate: equal-length slices of the sorted entries, one thread each, XYZZ += affine (8M+2S)  [ALU bound]
//   5. k_msm_merge     : buckets cut by slice boundaries = tail + heads (one thread per slice)
//   6. k_msm_reduce / k_msm_window_final: per-window sum_k k*B_k -> W window sums (XYZZ) on device
// The c*W doublings of the final Horner fold and the affine normalisation are O(W*c) work and run on the
// host (portable arithmetic in ff.cuh) because the result is consumed by the host-side transcript anyway.
// Multi-GPU: each rank runs 1-6 on its point range; window sums are all-gathered and folded (see capi.cu).
//
// Small signed scalars (the prover's Lagrange-basis commitments, zkaes_msm_g1_small): msm_small_window_sums below -- one digit per term,
// steps 3-6 on S pseudo-windows whose sums are added without doublings.
//
// Large inputs are processed in chunks of at most MSM_CHUNK points so that the sort scratch stays bounded;
// buckets persist across chunks.  No host synchronisation inside an MSM: grids are sized from upper bounds and the
// kernels read the actual entry count from offsets[nb].
#include "msm.cuh"

namespace zk {

static constexpr size_t MSM_CHUNK = (size_t)1 << 26;  // sort scratch: 4 B x W per point (3.2 GB at W = 12)

template <class FrP>
__device__ __forceinline__ void load_scalar(const uint32_t* scalars, size_t i, size_t stride, int mont, uint32_t* s) {
    const uint4* p = reinterpret_cast<const uint4*>(scalars) + 2 * i * stride;  // one 32-byte sector per scalar at any stride
    uint4 a = __ldg(p), b = __ldg(p + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    if (mont) {
        Fp<FrP> f;
#pragma unroll
        for (int k = 0; k < 8; ++k) f.v[k] = s[k];
        f = f.from_mont();
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = f.v[k];
    }
}

// Digits: one thread per scalar -- load (32 B), leave Montgomery form, fold to min(s, r - s), one carry walk over all W windows
// (msm_digits_all), W coalesced 4-byte stores into the window-major digit array (digits[w * m + i]).
template <class FrP>
__global__ void __launch_bounds__(256) k_msm_digits_all(const uint32_t* __restrict__ scalars, size_t n, size_t stride, int mont, MsmPlan p,
                                                        uint32_t* __restrict__ digits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar<FrP>(scalars, i, stride, mont, s);
    const uint32_t flip = msm_fold_scalar<FrP>(s);
    msm_digits_all(s, flip, p, digits + i, n);
}

// Counting sort by (window, bucket) over the precomputed digits.  pass 1 (sorted == nullptr): histogram.  pass 2: write sorted
// indices.  grid = (scalars / 256, windows): blockIdx.y is the window, so the blocks of one window run together and its 2^(c-1)
// counters (8 MB at c = 22) stay L2-resident under the atomics; each thread reads ONE digit word (4 B, coalesced).
__global__ void __launch_bounds__(256) k_msm_sort_pass(const uint32_t* __restrict__ digits, size_t n, uint32_t idx_base, uint32_t nbw, int w_base,
                                                       uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                                       uint32_t* __restrict__ sorted) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = digits[(size_t)(w_base + blockIdx.y) * n + i];
    if (e == MSM_DIGIT_NONE) return;
    const uint32_t key = (uint32_t)blockIdx.y * nbw + (e & 0x7fffffffu);  // counts / offsets start at window w_base
    const uint32_t pos = atomicAdd(&counts[key], 1u);
    if (sorted) sorted[offsets[key] + pos] = (idx_base + (uint32_t)i) | (e & 0x80000000u);
}

// The partly filled top window (253 = 11 x 22 + 11 bits at c = 22) has only 2^11 possible digits: 2^26 global atomics on
// 2^11 addresses serialise in L2 (9.2 ms per pass against 0.7-2.5 ms for a full window, profiles/
// r1_launches_msm_2p26_pair_round.txt).  For such a window each block histograms a tile of 16,384 digits in shared
// memory and touches every global counter once; positions inside the tile come from shared-memory cursors.
static constexpr int DIG_TILE_ITERS = 64;  // x 256 threads
static constexpr uint32_t DIG_SMALL_KEYS = 4096;
__global__ void __launch_bounds__(256) k_msm_sort_pass_small(const uint32_t* __restrict__ digits_w, size_t n, uint32_t idx_base, uint32_t nkeys,
                                                             uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                                             uint32_t* __restrict__ sorted) {
    __shared__ uint32_t hist[DIG_SMALL_KEYS], cursor[DIG_SMALL_KEYS];
    for (uint32_t k = threadIdx.x; k < nkeys; k += 256) hist[k] = 0;
    __syncthreads();
    const size_t tile = (size_t)blockIdx.x * (256 * DIG_TILE_ITERS);
    for (int it = 0; it < DIG_TILE_ITERS; ++it) {
        const size_t i = tile + (size_t)it * 256 + threadIdx.x;
        if (i >= n) break;
        const uint32_t e = digits_w[i];
        if (e != MSM_DIGIT_NONE) atomicAdd(&hist[e & 0x7fffffffu], 1u);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < nkeys; k += 256) {
        const uint32_t c = hist[k];
        cursor[k] = c ? atomicAdd(&counts[k], c) : 0u;  // this tile's first position inside bucket k
    }
    if (!sorted) return;
    __syncthreads();
    for (int it = 0; it < DIG_TILE_ITERS; ++it) {
        const size_t i = tile + (size_t)it * 256 + threadIdx.x;
        if (i >= n) break;
        const uint32_t e = digits_w[i];
        if (e == MSM_DIGIT_NONE) continue;
        const uint32_t k = e & 0x7fffffffu;
        const uint32_t pos = atomicAdd(&cursor[k], 1u);
        sorted[offsets[k] + pos] = (idx_base + (uint32_t)i) | (e & 0x80000000u);
    }
}
// number of distinct |digits| the top window can take, or 0 if it is a full window / too large for the shared-memory path
static uint32_t msm_small_top_keys(const MsmPlan& p, int fr_bits) {
    const int top_bits = fr_bits - (p.W - 1) * p.c;
    if (top_bits >= p.c || top_bits < 1 || ((uint32_t)1 << top_bits) > DIG_SMALL_KEYS) return 0;
    return (uint32_t)1 << top_bits;
}
template <class FrP>
static void msm_launch_digits_all(zkaes_ctx* ctx, const uint32_t* sc, size_t m, size_t stride, int mont, const MsmPlan& p, uint32_t* digits) {
    k_msm_digits_all<FrP><<<cdiv(m, 256), 256, 0, ctx->stream>>>(sc, m, stride, mont, p, digits);
    ctx->launches++;
}
// histogram (sorted == nullptr) or scatter pass over windows [w0, w0 + nw) of the digit array; counts / offsets point at window w0's
// first bucket
static void msm_launch_sort_pass(zkaes_ctx* ctx, const uint32_t* digits, size_t m, uint32_t idx_base, const MsmPlan& p, int fr_bits, int w0, int nw,
                                 uint32_t* counts, const uint32_t* offsets, uint32_t* sorted) {
    cudaStream_t st = ctx->stream;
    const uint32_t small = msm_small_top_keys(p, fr_bits);
    int n_full = nw;
    if (small && w0 + nw == p.W) {  // the range ends with the partly filled top window
        --n_full;
        const size_t off = (size_t)n_full * p.nbw;
        k_msm_sort_pass_small<<<cdiv(m, 256 * DIG_TILE_ITERS), 256, 0, st>>>(digits + (size_t)(p.W - 1) * m, m, idx_base, small, counts + off,
                                                                             offsets ? offsets + off : nullptr, sorted);
        ctx->launches++;
    }
    if (n_full > 0) {
        k_msm_sort_pass<<<dim3(cdiv(m, 256), n_full), 256, 0, st>>>(digits, m, idx_base, p.nbw, w0, counts, offsets, sorted);
        ctx->launches++;
    }
}

// ---- exclusive scan over <= 2^30 counters: per-block scan, scan of block sums (two levels), add-back -------
static constexpr int SCAN_BS = 1024;
__global__ void __launch_bounds__(SCAN_BS) k_scan_blocks(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                                                         uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t sh[SCAN_BS];
    uint32_t i = blockIdx.x * SCAN_BS + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < SCAN_BS; d <<= 1) {
        uint32_t t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = sh[threadIdx.x] - v;  // exclusive
    if (threadIdx.x == SCAN_BS - 1 && block_sums) block_sums[blockIdx.x] = sh[threadIdx.x];
}
__global__ void k_scan_add(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ block_offsets) {
    uint32_t i = blockIdx.x * SCAN_BS + threadIdx.x;
    if (i < n) out[i] += block_offsets[blockIdx.x];
}

int exclusive_scan_u32(zkaes_ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n) {
    cudaStream_t st = ctx->stream;
    uint32_t nblk = cdiv(n, SCAN_BS);
    if (nblk > SCAN_BS * SCAN_BS) return fail(ctx, ZK_ERR_UNSUPPORTED, "scan too large");
    DevBuf sums, sums2, top;
    ZK_CUDA(ctx, sums.alloc(sizeof(uint32_t) * nblk, st));
    k_scan_blocks<<<nblk, SCAN_BS, 0, st>>>(in, out, n, sums.as<uint32_t>());
    ctx->launches++;
    if (nblk > 1) {
        uint32_t nblk2 = cdiv(nblk, SCAN_BS);
        ZK_CUDA(ctx, sums2.alloc(sizeof(uint32_t) * nblk, st));
        ZK_CUDA(ctx, top.alloc(sizeof(uint32_t) * nblk2, st));
        k_scan_blocks<<<nblk2, SCAN_BS, 0, st>>>(sums.as<uint32_t>(), sums2.as<uint32_t>(), nblk, top.as<uint32_t>());
        ctx->launches++;
        if (nblk2 > 1) {
            DevBuf top2;
            ZK_CUDA(ctx, top2.alloc(sizeof(uint32_t) * nblk2, st));
            k_scan_blocks<<<1, SCAN_BS, 0, st>>>(top.as<uint32_t>(), top2.as<uint32_t>(), nblk2, nullptr);
            k_scan_add<<<nblk2, SCAN_BS, 0, st>>>(sums2.as<uint32_t>(), nblk, top2.as<uint32_t>());
            ctx->launches += 2;
        }
        k_scan_add<<<nblk, SCAN_BS, 0, st>>>(out, n, sums2.as<uint32_t>());
        ctx->launches++;
    }
    ZK_CUDA(ctx, cudaGetLastError());
    return ZK_OK;
}

// ---- bucket accumulation ---------------------------------------------------------------------------------
// Bases are read in the arkworks form (12 x u32 Montgomery limbs per coordinate, 96 B per point, 6 x 128-bit loads).
// BLOCKS = resident blocks per SM the register allocation is held to: 3 (166 registers, no spills) or 4 (128 registers,
// 92 bytes of spills, 16 instead of 12 warps per SM to cover the IMAD dependency waits).
// CALL: the ten field products of a mixed addition go through one out-of-line multiplier (XYZZ::madd_call) instead of ten
// inlined copies -- the loop body then fits the instruction cache.  PF: the next entry's point is staged in shared memory while the
// current one is added (msm_core.cuh: 1 = cp.async, 2 = cp.async.bulk + mbarrier).
template <class C, int BLOCKS, bool CALL, int PF>
__global__ void __launch_bounds__(128, BLOCKS) k_msm_accumulate(const uint32_t* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                                const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t n_slices, uint32_t L,
                                                                XYZZ<C>* __restrict__ buckets, XYZZ<C>* __restrict__ head,
                                                                XYZZ<C>* __restrict__ tail, uint32_t* __restrict__ tail_bucket) {
    __shared__ __align__(128) uint8_t s_pts[PF ? 128 * 192 : 16];
    __shared__ __align__(16) uint64_t s_bars[PF == 2 ? 128 * 2 : 2];
    msm_slice_accumulate<C, CALL, PF>(blockIdx.x * blockDim.x + threadIdx.x, n_slices, L, offsets, nb, sorted, bases, buckets, head, tail, tail_bucket,
                                      s_pts, s_bars);
}
template <class C>
static void msm_launch_accumulate(zkaes_ctx* ctx, size_t slices, const uint32_t* bases, const uint32_t* sorted, const uint32_t* offsets, uint32_t nb,
                                  uint32_t L, XYZZ<C>* buckets, XYZZ<C>* head, XYZZ<C>* tail, uint32_t* tail_bucket) {
    const unsigned grid = cdiv(slices, 128);
    cudaStream_t st = ctx->stream;
#define ZK_ACC(B, CALLV, PFV) k_msm_accumulate<C, B, CALLV, PFV><<<grid, 128, 0, st>>>(bases, sorted, offsets, nb, (uint32_t)slices, L, buckets, head, tail, tail_bucket)
    if (ctx->msm_prefetch == 1) {
        ZK_ACC(3, true, 1);
    } else if (ctx->msm_prefetch == 2) {
        ZK_ACC(3, true, 2);
    } else if (ctx->msm_madd_call) {
        if (ctx->msm_acc_blocks == 4) ZK_ACC(4, true, 0); else ZK_ACC(3, true, 0);
    } else {
        if (ctx->msm_acc_blocks == 4) ZK_ACC(4, false, 0); else ZK_ACC(3, false, 0);
    }
#undef ZK_ACC
    ctx->launches++;
}

template <class C>
__global__ void __launch_bounds__(128) k_msm_merge(const uint32_t* __restrict__ offsets, uint32_t n_slices, uint32_t L, XYZZ<C>* __restrict__ buckets,
                                                   const XYZZ<C>* __restrict__ head, const XYZZ<C>* __restrict__ tail,
                                                   const uint32_t* __restrict__ tail_bucket) {
    msm_merge_slice<C>(blockIdx.x * blockDim.x + threadIdx.x, n_slices, L, offsets, buckets, head, tail, tail_bucket);
}
// 4 warps per block, one slice per warp; only the slices whose bucket runs on for more than MSM_MERGE_SERIAL_MAX slices do work
template <class C>
__global__ void __launch_bounds__(128) k_msm_merge_heavy(const uint32_t* __restrict__ offsets, uint32_t n_slices, uint32_t L, XYZZ<C>* __restrict__ buckets,
                                                   const XYZZ<C>* __restrict__ head, const XYZZ<C>* __restrict__ tail,
                                                   const uint32_t* __restrict__ tail_bucket) {
    __shared__ XYZZ<C> sh[128];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * 4 + warp;
    XYZZ<C> acc;
    uint32_t b;
    if (!msm_merge_lane<C>(t, lane, n_slices, L, offsets, head, tail, tail_bucket, &acc, &b)) return;  // warp-uniform
    sh[threadIdx.x] = acc;
    __syncwarp();
    for (int s = 16; s > 0; s >>= 1) {
        if ((int)lane < s) {
            XYZZ<C> x = sh[threadIdx.x];
            x.add(sh[threadIdx.x + s]);  // most lanes hold infinity: add() returns at once
            sh[threadIdx.x] = x;
        }
        __syncwarp();
    }
    if (lane == 0) msm_store_xyzz<C>(buckets + b, sh[threadIdx.x]);
}

// ---- bucket reduction: S_w = sum_{j<nbw} (j+1) * B[w][j] ---------------------------------------------------
static constexpr int RED_BS = 128;
// grid: (blocks_per_window, W).  Each thread owns `seg` consecutive buckets of its window.
template <class C>
__global__ void __launch_bounds__(RED_BS) k_msm_reduce(const XYZZ<C>* __restrict__ buckets, uint32_t nbw, uint32_t seg,
                                                       XYZZ<C>* __restrict__ partials) {
    __shared__ XYZZ<C> sh[RED_BS];
    uint32_t w = blockIdx.y;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // segment id within window
    sh[threadIdx.x] = msm_reduce_segment<C>(buckets + (size_t)w * nbw, nbw, seg, t);
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            XYZZ<C> a = sh[threadIdx.x];
            a.add(sh[threadIdx.x + s]);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[(size_t)w * gridDim.x + blockIdx.x] = sh[0];
}

// one warp per window: lane-strided partial sums, then a shuffle-free fold through shared memory
template <class C>
__global__ void __launch_bounds__(32) k_msm_window_final(const XYZZ<C>* __restrict__ partials, uint32_t per_window, XYZZ<C>* __restrict__ out) {
    __shared__ XYZZ<C> sh[32];
    const uint32_t w = blockIdx.x;
    XYZZ<C> a = XYZZ<C>::inf();
    for (uint32_t k = threadIdx.x; k < per_window; k += 32) a.add(partials[(size_t)w * per_window + k]);
    sh[threadIdx.x] = a;
    __syncwarp();
    for (int s = 16; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            XYZZ<C> x = sh[threadIdx.x];
            x.add(sh[threadIdx.x + s]);
            sh[threadIdx.x] = x;
        }
        __syncwarp();
    }
    if (threadIdx.x == 0) out[w] = sh[0];
}

// ---- pair rounds (msm_core.cuh): batched-affine first levels of the bucket sums ------------------------------------------------
__global__ void __launch_bounds__(256) k_pad_pow2(const uint32_t* __restrict__ counts, uint32_t* __restrict__ padded, uint32_t n, uint32_t align) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) padded[i] = (counts[i] + align - 1) & ~(align - 1);
}
__global__ void __launch_bounds__(256) k_shift_right(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, int shift) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] >> shift;
}
// `round` = 1, 2, ...: the round's pair count is (padded entry count) >> round, read from off2[nbw] on the device
template <class C>
__global__ void __launch_bounds__(128) k_pair_products(const uint32_t* __restrict__ off2, uint32_t nbw, int round, uint32_t G,
                                                       const uint32_t* __restrict__ idx, const uint32_t* __restrict__ pts,
                                                       typename Affine<C>::Fq* __restrict__ prefix, typename Affine<C>::Fq* __restrict__ tprod) {
    msm_pair_products<C>(blockIdx.x * blockDim.x + threadIdx.x, G, off2[nbw] >> round, idx, pts, prefix, tprod);
}
template <class C>
__global__ void __launch_bounds__(128) k_pair_invert(const uint32_t* __restrict__ off2, uint32_t nbw, int round, uint32_t G, uint32_t G2,
                                                     typename Affine<C>::Fq* __restrict__ tprod, typename Affine<C>::Fq* __restrict__ scratch) {
    const uint32_t n_pairs = off2[nbw] >> round;
    msm_pair_invert<typename Affine<C>::Fq>(blockIdx.x * blockDim.x + threadIdx.x, G2, (n_pairs + G - 1) / G, tprod, scratch);
}
template <class C>
__global__ void __launch_bounds__(128) k_pair_add(const uint32_t* __restrict__ off2, uint32_t nbw, int round, uint32_t G,
                                                  const uint32_t* __restrict__ idx, const uint32_t* __restrict__ pts,
                                                  const typename Affine<C>::Fq* __restrict__ prefix, const typename Affine<C>::Fq* __restrict__ tinv,
                                                  uint32_t* __restrict__ out) {
    msm_pair_add<C>(blockIdx.x * blockDim.x + threadIdx.x, G, off2[nbw] >> round, idx, pts, prefix, tinv, out);
}

static uint32_t msm_slice_len(uint64_t max_entries);

// One window at a time (the pair sums of a window are 96 B per two entries: 3.3 GB for 2^26 points, so they cannot be
// kept for all windows at once): 2^R-aligned sort, R pair rounds, slice accumulation of the last round's sums into the
// window's buckets.
template <class C>
int msm_accumulate_paired(zkaes_ctx* ctx, const uint32_t* bases, const uint32_t* scalars, size_t n, int scalars_mont, const MsmPlan& p,
                          size_t scalar_stride, size_t chunk_max, XYZZ<C>* buckets) {
    using FrP = typename C::FrP;
    using Fq = typename Affine<C>::Fq;
    cudaStream_t st = ctx->stream;
    constexpr uint32_t G = 64, G2 = 128;
    const int R = ctx->msm_pair_round;
    const uint32_t align = 1u << R;
    // entries after padding every non-empty bucket up to a multiple of 2^R
    const size_t max_entries2 = ((chunk_max + (size_t)(align - 1) * p.nbw) + align) & ~(size_t)(align - 1);
    const size_t max_pairs = max_entries2 / 2 + 1;       // round 1; every later round has half of the one before
    const size_t max_groups = (max_pairs + G - 1) / G + 1;
    const size_t max_final = (max_entries2 >> R) + 1;    // sums left for the accumulation
    const uint32_t L = msm_slice_len(max_final);
    const size_t max_slices = (max_final + L - 1) / L + 1;
    DevBuf counts, padded, off2, poff, sorted2, prefix, tprod, scratch, sums[2], head, tail, tail_bucket, digits;
    ZK_CUDA(ctx, digits.alloc(sizeof(uint32_t) * chunk_max * p.W, st));
    ZK_CUDA(ctx, counts.alloc(sizeof(uint32_t) * ((size_t)p.nbw + 1), st));
    ZK_CUDA(ctx, padded.alloc(sizeof(uint32_t) * ((size_t)p.nbw + 1), st));
    ZK_CUDA(ctx, off2.alloc(sizeof(uint32_t) * ((size_t)p.nbw + 1), st));
    ZK_CUDA(ctx, poff.alloc(sizeof(uint32_t) * ((size_t)p.nbw + 1), st));
    ZK_CUDA(ctx, sorted2.alloc(sizeof(uint32_t) * max_entries2, st));
    ZK_CUDA(ctx, prefix.alloc(sizeof(Fq) * max_pairs, st));
    ZK_CUDA(ctx, tprod.alloc(sizeof(Fq) * max_groups, st));
    ZK_CUDA(ctx, scratch.alloc(sizeof(Fq) * max_groups, st));
    ZK_CUDA(ctx, sums[1].alloc(96 * max_pairs, st));                       // rounds 1, 3, ...
    if (R > 1) ZK_CUDA(ctx, sums[0].alloc(96 * (max_pairs / 2 + 1), st));  // rounds 2, 4, ...
    ZK_CUDA(ctx, head.alloc(sizeof(XYZZ<C>) * max_slices, st));
    ZK_CUDA(ctx, tail.alloc(sizeof(XYZZ<C>) * max_slices, st));
    ZK_CUDA(ctx, tail_bucket.alloc(sizeof(uint32_t) * max_slices, st));
    for (size_t base = 0; base < n; base += chunk_max) {
        const size_t m = n - base < chunk_max ? n - base : chunk_max;
        const uint32_t* sc = scalars + 8 * base * scalar_stride;
        // upper bounds of this chunk's counts (the kernels read the exact ones from off2[nbw])
        const size_t entries_ub = ((m + (size_t)(align - 1) * p.nbw) + align) & ~(size_t)(align - 1);
        const size_t slices = ((entries_ub >> R) + L) / L;
        msm_launch_digits_all<FrP>(ctx, sc, m, scalar_stride, scalars_mont, p, digits.as<uint32_t>());
        for (int w = 0; w < p.W; ++w) {
            ZK_CUDA(ctx, cudaMemsetAsync(counts.p, 0, sizeof(uint32_t) * ((size_t)p.nbw + 1), st));
            msm_launch_sort_pass(ctx, digits.as<uint32_t>(), m, (uint32_t)base, p, FrP::BITS, w, 1, counts.as<uint32_t>(), nullptr, nullptr);
            k_pad_pow2<<<cdiv((size_t)p.nbw + 1, 256), 256, 0, st>>>(counts.as<uint32_t>(), padded.as<uint32_t>(), p.nbw + 1, align);
            ctx->launches++;
            ZK_TRY(exclusive_scan_u32(ctx, padded.as<uint32_t>(), off2.as<uint32_t>(), p.nbw + 1));
            k_shift_right<<<cdiv((size_t)p.nbw + 1, 256), 256, 0, st>>>(off2.as<uint32_t>(), poff.as<uint32_t>(), p.nbw + 1, R);
            ZK_CUDA(ctx, cudaMemsetAsync(counts.p, 0, sizeof(uint32_t) * (size_t)p.nbw, st));
            ZK_CUDA(ctx, cudaMemsetAsync(sorted2.p, 0xff, sizeof(uint32_t) * entries_ub, st));  // MSM_NONE in the padding slots
            msm_launch_sort_pass(ctx, digits.as<uint32_t>(), m, (uint32_t)base, p, FrP::BITS, w, 1, counts.as<uint32_t>(), off2.as<uint32_t>(),
                                 sorted2.as<uint32_t>());
            ctx->launches++;
            const uint32_t* idx = sorted2.as<uint32_t>();
            const uint32_t* pts = bases;
            for (int r = 1; r <= R; ++r) {
                const unsigned groups = cdiv((entries_ub >> r) + 1, G);
                uint32_t* out = sums[r & 1].as<uint32_t>();
                k_pair_products<C><<<cdiv(groups, 128), 128, 0, st>>>(off2.as<uint32_t>(), p.nbw, r, G, idx, pts, prefix.as<Fq>(), tprod.as<Fq>());
                k_pair_invert<C><<<cdiv(cdiv(groups, G2), 128), 128, 0, st>>>(off2.as<uint32_t>(), p.nbw, r, G, G2, tprod.as<Fq>(), scratch.as<Fq>());
                k_pair_add<C><<<cdiv(groups, 128), 128, 0, st>>>(off2.as<uint32_t>(), p.nbw, r, G, idx, pts, prefix.as<Fq>(), tprod.as<Fq>(), out);
                ctx->launches += 3;
                idx = nullptr;
                pts = out;
            }
            XYZZ<C>* B = buckets + (size_t)w * p.nbw;
            zkaes_ctx::ProfSpan span{};
            if (ctx->prof) {
                cudaEventCreate(&span.e0);
                cudaEventCreate(&span.e1);
                cudaEventRecord(span.e0, st);
            }
            msm_launch_accumulate<C>(ctx, slices, pts, nullptr, poff.as<uint32_t>(), p.nbw, L, B, head.as<XYZZ<C>>(), tail.as<XYZZ<C>>(),
                                     tail_bucket.as<uint32_t>());
            if (ctx->prof) {
                cudaEventRecord(span.e1, st);
                span.terms = w == 0 ? m : 0;               // each term is counted once per chunk
                span.madds = (uint64_t)(entries_ub >> R);  // upper bound: one mixed addition per remaining sum
                ctx->prof_spans.push_back(span);
            }
            k_msm_merge<C><<<cdiv(slices, 128), 128, 0, st>>>(poff.as<uint32_t>(), (uint32_t)slices, L, B, head.as<XYZZ<C>>(), tail.as<XYZZ<C>>(),
                                                              tail_bucket.as<uint32_t>());
            k_msm_merge_heavy<C><<<cdiv(slices, 4), 128, 0, st>>>(poff.as<uint32_t>(), (uint32_t)slices, L, B, head.as<XYZZ<C>>(), tail.as<XYZZ<C>>(),
                                                                  tail_bucket.as<uint32_t>());
            ctx->launches += 2;
            ZK_CUDA(ctx, cudaGetLastError());
        }
    }
    return ZK_OK;
}

// slice length of the accumulation: long enough to amortise the per-slice bookkeeping and keep the partial arrays
// small, short enough for >= ~8 waves of threads on 148 SMs x 384 resident threads
static uint32_t msm_slice_len(uint64_t max_entries) {
    uint64_t L = max_entries / (148ull * 384 * 8);
    if (L < 32) L = 32;
    if (L > 512) L = 512;
    return (uint32_t)L;
}

template <class C>
int msm_window_sums(zkaes_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n, int scalars_mont, const MsmPlan& p,
                    void* d_window_sums, int bases_internal, size_t scalar_stride) {
    using FrP = typename C::FrP;
    cudaStream_t st = ctx->stream;
    if (n == 0) {
        ZK_CUDA(ctx, cudaMemsetAsync(d_window_sums, 0, sizeof(XYZZ<C>) * p.W, st));
        return ZK_OK;
    }
    // entries per chunk stay below 2^31 (u32 offsets, sign bit in the sorted indices)
    size_t chunk_max = MSM_CHUNK;
    while (chunk_max * (size_t)p.W >= ((size_t)1 << 31)) chunk_max >>= 1;
    chunk_max = (n + (n + chunk_max - 1) / chunk_max - 1) / ((n + chunk_max - 1) / chunk_max);  // equal chunks (n = 2^k + 1 must not leave a 1-point chunk)
    if (n >= ((size_t)1 << 31)) return fail(ctx, ZK_ERR_UNSUPPORTED, "msm: more than 2^31 points");
    const uint64_t max_entries = (uint64_t)chunk_max * p.W;
    const uint32_t L = msm_slice_len(max_entries);
    const size_t max_slices = (size_t)((max_entries + L - 1) / L);
    // the pair round (batched-affine first level) pays off once buckets hold several entries; it needs the field inversion,
    // which only the default (arkworks-limb) form of the curve arithmetic provides
    const bool paired = ctx->msm_pair_round > 0 && n >= ((size_t)1 << 16) && (n / p.nbw) >> ctx->msm_pair_round >= 2;
    DevBuf counts, offsets, sorted, buckets, partials, head, tail, tail_bucket, digits;
    ZK_CUDA(ctx, buckets.alloc(sizeof(XYZZ<C>) * (size_t)p.nb, st));
    ZK_CUDA(ctx, cudaMemsetAsync(buckets.p, 0, sizeof(XYZZ<C>) * (size_t)p.nb, st));  // all-zero XYZZ = infinity
    if (!paired) {
        ZK_CUDA(ctx, counts.alloc(sizeof(uint32_t) * ((size_t)p.nb + 1), st));
        ZK_CUDA(ctx, offsets.alloc(sizeof(uint32_t) * ((size_t)p.nb + 1), st));
        ZK_CUDA(ctx, sorted.alloc(sizeof(uint32_t) * max_entries, st));
        ZK_CUDA(ctx, digits.alloc(sizeof(uint32_t) * max_entries, st));  // window-major signed digits of the chunk, computed once
        ZK_CUDA(ctx, head.alloc(sizeof(XYZZ<C>) * max_slices, st));
        ZK_CUDA(ctx, tail.alloc(sizeof(XYZZ<C>) * max_slices, st));
        ZK_CUDA(ctx, tail_bucket.alloc(sizeof(uint32_t) * max_slices, st));
    }
    const uint32_t* bases = reinterpret_cast<const uint32_t*>(d_bases);
    (void)bases_internal;  // the kernels compute in the arkworks form: prepared and unprepared bases are the same bytes
    const auto* scalars = reinterpret_cast<const uint32_t*>(d_scalars);
    if (paired) {
        ZK_TRY(msm_accumulate_paired<C>(ctx, bases, scalars, n, scalars_mont, p, scalar_stride, chunk_max, buckets.as<XYZZ<C>>()));
    }
    for (size_t base = 0; !paired && base < n; base += chunk_max) {
        size_t m = n - base < chunk_max ? n - base : chunk_max;
        const uint32_t* sc = scalars + 8 * base * scalar_stride;
        ZK_CUDA(ctx, cudaMemsetAsync(counts.p, 0, sizeof(uint32_t) * ((size_t)p.nb + 1), st));
        msm_launch_digits_all<FrP>(ctx, sc, m, scalar_stride, scalars_mont, p, digits.as<uint32_t>());
        msm_launch_sort_pass(ctx, digits.as<uint32_t>(), m, (uint32_t)base, p, FrP::BITS, 0, p.W, counts.as<uint32_t>(), nullptr, nullptr);
        // scanning nb + 1 counters (the last one is zero) leaves the total entry count in offsets[nb]
        ZK_TRY(exclusive_scan_u32(ctx, counts.as<uint32_t>(), offsets.as<uint32_t>(), p.nb + 1));
        ZK_CUDA(ctx, cudaMemsetAsync(counts.p, 0, sizeof(uint32_t) * (size_t)p.nb, st));
        msm_launch_sort_pass(ctx, digits.as<uint32_t>(), m, (uint32_t)base, p, FrP::BITS, 0, p.W, counts.as<uint32_t>(), offsets.as<uint32_t>(),
                             sorted.as<uint32_t>());
        const size_t slices = (size_t)(((uint64_t)m * p.W + L - 1) / L);  // upper bound: zero digits produce no entry
        zkaes_ctx::ProfSpan span{};
        if (ctx->prof) {
            cudaEventCreate(&span.e0);
            cudaEventCreate(&span.e1);
            cudaEventRecord(span.e0, st);
        }
        msm_launch_accumulate<C>(ctx, slices, bases, sorted.as<uint32_t>(), offsets.as<uint32_t>(), p.nb, L, buckets.as<XYZZ<C>>(), head.as<XYZZ<C>>(),
                                 tail.as<XYZZ<C>>(), tail_bucket.as<uint32_t>());
        if (ctx->prof) {
            cudaEventRecord(span.e1, st);
            span.terms = m;
            span.madds = (uint64_t)m * p.W;  // upper bound: zero digits are skipped
            ctx->prof_spans.push_back(span);
        }
        k_msm_merge<C><<<cdiv(slices, 128), 128, 0, st>>>(offsets.as<uint32_t>(), (uint32_t)slices, L, buckets.as<XYZZ<C>>(), head.as<XYZZ<C>>(),
                                                           tail.as<XYZZ<C>>(), tail_bucket.as<uint32_t>());
        k_msm_merge_heavy<C><<<cdiv(slices, 4), 128, 0, st>>>(offsets.as<uint32_t>(), (uint32_t)slices, L, buckets.as<XYZZ<C>>(), head.as<XYZZ<C>>(),
                                                               tail.as<XYZZ<C>>(), tail_bucket.as<uint32_t>());
        ctx->launches++;
        ctx->launches++;
        ZK_CUDA(ctx, cudaGetLastError());
    }
    // reduction: segments of <= 64 buckets (the per-segment lo * running product costs ~2 log2(nbw) group operations)
    uint32_t tpw = p.nbw < 2048 ? p.nbw : 2048;  // threads (segments) per window
    if (p.nbw / 64 > tpw) tpw = p.nbw / 64;
    uint32_t seg = p.nbw / tpw;
    uint32_t bs = tpw < RED_BS ? tpw : RED_BS;
    uint32_t bpw = tpw / bs;
    ZK_CUDA(ctx, partials.alloc(sizeof(XYZZ<C>) * (size_t)bpw * p.W, st));
    k_msm_reduce<C><<<dim3(bpw, p.W), bs, 0, st>>>(buckets.as<XYZZ<C>>(), p.nbw, seg, partials.as<XYZZ<C>>());
    k_msm_window_final<C><<<p.W, 32, 0, st>>>(partials.as<XYZZ<C>>(), bpw, reinterpret_cast<XYZZ<C>*>(d_window_sums));
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaGetLastError());
    return ZK_OK;
}

// ---- small-scalar MSM: sum_j v_j * bases[j] for signed integers |v_j| <= 2^(c-1), c <= 13 -------------------------------------------------
// The prover's Lagrange-basis commitments (prover.cu, round 1): the evaluations of w, z_A, z_B over H are bits or sums of a few bits,
// so one c-bit digit per term is the whole scalar -- n_nonzero mixed additions instead of W * n.  The same sort / slice accumulation /
// merge / reduce kernels run on a plan of S PSEUDO-WINDOWS: window s takes the terms [s m, (s + 1) m), m = ceil(n / S), with its own
// 2^(c-1) buckets (so the few distinct values do not put all terms into one bucket run), and the result is the PLAIN sum of the S window
// sums (no doublings between them: msm_sum_windows_host).  vals[start + j * stride] is term j's value (cyclic sharding reads every N-th).
__global__ void __launch_bounds__(256) k_msm_digits_small(const int32_t* __restrict__ vals, size_t n, size_t start, size_t stride, size_t padded,
                                                          uint32_t* __restrict__ digits) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= padded) return;
    digits[j] = j < n ? msm_small_digit(vals[start + j * stride]) : MSM_DIGIT_NONE;
}

template <class C>
int msm_small_window_sums(zkaes_ctx* ctx, const void* d_bases, const int32_t* d_vals, size_t n, size_t val_start, size_t val_stride, int c, int S,
                          void* d_window_sums) {
    cudaStream_t st = ctx->stream;
    if (c < 1 || ((uint32_t)1 << (c - 1)) > DIG_SMALL_KEYS) return fail(ctx, ZK_ERR_ARG, "small msm: digit width out of range");
    if (S < 1 || n >= ((size_t)1 << 31)) return fail(ctx, ZK_ERR_ARG, "small msm: bad sizes");
    if (n == 0) {
        ZK_CUDA(ctx, cudaMemsetAsync(d_window_sums, 0, sizeof(XYZZ<C>) * S, st));
        return ZK_OK;
    }
    MsmPlan p;
    p.c = c;
    p.W = S;
    p.nbw = (uint32_t)1 << (c - 1);
    p.nb = p.nbw * (uint32_t)S;
    const size_t m = (n + (size_t)S - 1) / (size_t)S, padded = m * (size_t)S;
    const uint32_t L = msm_slice_len(padded);
    const size_t slices = (padded + L - 1) / L;
    DevBuf counts, offsets, sorted, buckets, partials, head, tail, tail_bucket, digits;
    ZK_CUDA(ctx, buckets.alloc(sizeof(XYZZ<C>) * (size_t)p.nb, st));
    ZK_CUDA(ctx, cudaMemsetAsync(buckets.p, 0, sizeof(XYZZ<C>) * (size_t)p.nb, st));
    ZK_CUDA(ctx, counts.alloc(sizeof(uint32_t) * ((size_t)p.nb + 1), st));
    ZK_CUDA(ctx, offsets.alloc(sizeof(uint32_t) * ((size_t)p.nb + 1), st));
    ZK_CUDA(ctx, sorted.alloc(sizeof(uint32_t) * padded, st));
    ZK_CUDA(ctx, digits.alloc(sizeof(uint32_t) * padded, st));
    ZK_CUDA(ctx, head.alloc(sizeof(XYZZ<C>) * slices, st));
    ZK_CUDA(ctx, tail.alloc(sizeof(XYZZ<C>) * slices, st));
    ZK_CUDA(ctx, tail_bucket.alloc(sizeof(uint32_t) * slices, st));
    k_msm_digits_small<<<cdiv(padded, 256), 256, 0, st>>>(d_vals, n, val_start, val_stride, padded, digits.as<uint32_t>());
    ctx->launches++;
    // few keys per window: the shared-memory histogram passes (k_msm_sort_pass_small), one launch per pseudo-window; entry = the term's index
    auto sort_pass = [&](const uint32_t* offs, uint32_t* out) {
        for (int s = 0; s < S; ++s) {
            const size_t off = (size_t)s * p.nbw;
            k_msm_sort_pass_small<<<cdiv(m, 256 * DIG_TILE_ITERS), 256, 0, st>>>(digits.as<uint32_t>() + (size_t)s * m, m, (uint32_t)((size_t)s * m), p.nbw,
                                                                                 counts.as<uint32_t>() + off, offs ? offs + off : nullptr, out);
        }
        ctx->launches += S;
    };
    ZK_CUDA(ctx, cudaMemsetAsync(counts.p, 0, sizeof(uint32_t) * ((size_t)p.nb + 1), st));
    sort_pass(nullptr, nullptr);
    ZK_TRY(exclusive_scan_u32(ctx, counts.as<uint32_t>(), offsets.as<uint32_t>(), p.nb + 1));
    ZK_CUDA(ctx, cudaMemsetAsync(counts.p, 0, sizeof(uint32_t) * (size_t)p.nb, st));
    sort_pass(offsets.as<uint32_t>(), sorted.as<uint32_t>());
    const uint32_t* bases = reinterpret_cast<const uint32_t*>(d_bases);
    // not entered in the profile spans (zkaes_ctx_profile): their mixed-addition count is an upper bound taken from the term count, and here
    // zero values -- half of a bit vector -- produce no entry; bench.py's roofline is about the general-scalar launches
    msm_launch_accumulate<C>(ctx, slices, bases, sorted.as<uint32_t>(), offsets.as<uint32_t>(), p.nb, L, buckets.as<XYZZ<C>>(), head.as<XYZZ<C>>(),
                             tail.as<XYZZ<C>>(), tail_bucket.as<uint32_t>());
    k_msm_merge<C><<<cdiv(slices, 128), 128, 0, st>>>(offsets.as<uint32_t>(), (uint32_t)slices, L, buckets.as<XYZZ<C>>(), head.as<XYZZ<C>>(),
                                                       tail.as<XYZZ<C>>(), tail_bucket.as<uint32_t>());
    k_msm_merge_heavy<C><<<cdiv(slices, 4), 128, 0, st>>>(offsets.as<uint32_t>(), (uint32_t)slices, L, buckets.as<XYZZ<C>>(), head.as<XYZZ<C>>(),
                                                           tail.as<XYZZ<C>>(), tail_bucket.as<uint32_t>());
    const uint32_t bs = p.nbw < RED_BS ? p.nbw : RED_BS;  // one bucket per thread
    const uint32_t bpw = p.nbw / bs;
    ZK_CUDA(ctx, partials.alloc(sizeof(XYZZ<C>) * (size_t)bpw * S, st));
    k_msm_reduce<C><<<dim3(bpw, S), bs, 0, st>>>(buckets.as<XYZZ<C>>(), p.nbw, 1, partials.as<XYZZ<C>>());
    k_msm_window_final<C><<<S, 32, 0, st>>>(partials.as<XYZZ<C>>(), bpw, reinterpret_cast<XYZZ<C>*>(d_window_sums));
    ctx->launches += 4;
    ZK_CUDA(ctx, cudaGetLastError());
    return ZK_OK;
}
template int msm_small_window_sums<G1_377Params>(zkaes_ctx*, const void*, const int32_t*, size_t, size_t, size_t, int, int, void*);
template int msm_small_window_sums<G1_381Params>(zkaes_ctx*, const void*, const int32_t*, size_t, size_t, size_t, int, int, void*);

template <class C>
Affine<C> msm_sum_windows_host(const XYZZ<C>* sums, size_t count) {
    XYZZ<C> total = XYZZ<C>::inf();
    for (size_t i = 0; i < count; ++i) total.add(sums[i]);
    return total.to_affine();
}
template Affine<G1_377Params> msm_sum_windows_host<G1_377Params>(const XYZZ<G1_377Params>*, size_t);
template Affine<G1_381Params> msm_sum_windows_host<G1_381Params>(const XYZZ<G1_381Params>*, size_t);

// Horner fold of window sums (host): total = sum_w 2^(c*w) * S_w, then to affine.
template <class C>
Affine<C> msm_fold_windows_host(const XYZZ<C>* sums, int n_sets, const MsmPlan& p) {
    XYZZ<C> total = XYZZ<C>::inf();
    for (int w = p.W - 1; w >= 0; --w) {
        for (int k = 0; k < p.c; ++k) total = total.dbl();
        for (int s = 0; s < n_sets; ++s) total.add(sums[(size_t)s * p.W + w]);
    }
    return total.to_affine();
}

template <class C>
int msm_to_affine(zkaes_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n, int scalars_mont, Affine<C>* out, int bases_internal) {
    MsmPlan p = msm_make_plan(n ? n : 1, C::FrP::BITS, ctx->msm_window_bits);
    DevBuf win;
    ZK_CUDA(ctx, win.alloc(sizeof(XYZZ<C>) * p.W, ctx->stream));
    ZK_TRY(msm_window_sums<C>(ctx, d_bases, d_scalars, n, scalars_mont, p, win.p, bases_internal));
    std::vector<XYZZ<C>> h(p.W);
    ZK_CUDA(ctx, cudaMemcpyAsync(h.data(), win.p, sizeof(XYZZ<C>) * p.W, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = msm_fold_windows_host<C>(h.data(), 1, p);
    return ZK_OK;
}
template int msm_to_affine<G1_377Params>(zkaes_ctx*, const void*, const void*, size_t, int, Affine<G1_377Params>*, int);
template int msm_to_affine<G1_381Params>(zkaes_ctx*, const void*, const void*, size_t, int, Affine<G1_381Params>*, int);

// Kept for the ABI (zkaes_msm_g1_prepare_bases): the kernels read the arkworks form directly, so preparing is the identity.
template <class C>
int msm_bases_to_internal(zkaes_ctx* ctx, const void* src, void* dst, size_t n) {
    if (n && dst != src) ZK_CUDA(ctx, cudaMemcpyAsync(dst, src, 96 * n, cudaMemcpyDeviceToDevice, ctx->stream));
    return ZK_OK;
}
template int msm_bases_to_internal<G1_377Params>(zkaes_ctx*, const void*, void*, size_t);
template int msm_bases_to_internal<G1_381Params>(zkaes_ctx*, const void*, void*, size_t);

template int msm_window_sums<G1_377Params>(zkaes_ctx*, const void*, const void*, size_t, int, const MsmPlan&, void*, int, size_t);
template int msm_window_sums<G1_381Params>(zkaes_ctx*, const void*, const void*, size_t, int, const MsmPlan&, void*, int, size_t);
template Affine<G1_377Params> msm_fold_windows_host<G1_377Params>(const XYZZ<G1_377Params>*, int, const MsmPlan&);
template Affine<G1_381Params> msm_fold_windows_host<G1_381Params>(const XYZZ<G1_381Params>*, int, const MsmPlan&);

}  // namespace zk
