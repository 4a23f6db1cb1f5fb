// Per-thread bodies of the MSM kernels (msm.cu), written as host/device functions of the thread index so that the
// same code runs inside the CUDA kernels and, sequentially, inside the CPU emulation that tests/ uses to check the index
// bookkeeping without a GPU (msm_emul.cpp).  Nothing here is a CPU fallback: the product only ever launches the kernels.
//
// Pipeline (per chunk of at most `chunk` points; buckets persist across chunks):
//   digits     s -> min(s, r - s) (sign folded into the point), then signed c-bit digits; key = window * 2^(c-1) + |d| - 1
//   sort       counting sort of point indices by key (histogram, exclusive scan, scatter)
//   accumulate the sorted entry array is cut into SLICES OF EQUAL LENGTH L, one thread each, regardless of bucket
//              boundaries: every thread performs exactly L mixed additions, so a warp never waits for its most loaded
//              bucket (the thread-per-bucket form ran at 26.4 / 32 active lanes, profiles/r1_ncu_full_msm_accumulate_c16.txt).
//              A bucket that lies inside one slice is updated in place; a bucket cut by slice boundaries leaves one
//              partial per slice (`tail` for the slice it starts in, `head` for the slices it continues into).
//   merge      bucket = tail + sum of heads, for the buckets that were cut
//   reduce     S_w = sum_j (j + 1) B[w][j]
#pragma once
#include "ec.cuh"

namespace zk {

struct MsmPlan {
    int c;         // window bits (signed digits in [-2^(c-1), 2^(c-1)])
    int W;         // number of windows = ceil(Fr bits / c): scalars are folded to s <= (r-1)/2 first
    uint32_t nbw;  // buckets per window = 2^(c-1)
    uint32_t nb;   // total buckets
};

// ---- digits --------------------------------------------------------------------------------------------------------
// s (canonical, < r) -> min(s, r - s); returns 1 when r - s was taken (the term becomes (r - s) * (-P)).
template <class FrP>
ZK_HD uint32_t msm_fold_scalar(uint32_t* s) {
    uint32_t m[8], t[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = FrP::MOD(i);
    sub_n<8>(t, m, s);                  // r - s  (s = 0 gives r, which is not < s)
    uint32_t lt = sub_n<8>(d, t, s);    // borrow <=> r - s < s
    if (lt) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = t[i];
    }
    return lt;
}

ZK_HD uint32_t msm_scalar_bits(const uint32_t* s, int pos, int c) {
    int limb = pos >> 5, off = pos & 31;
    if (limb >= 8) return 0;
    uint64_t v = s[limb];
    if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
    return (uint32_t)(v >> off) & ((1u << c) - 1);
}

// The signed digit of window w alone (0 = no entry): the per-window sort passes call this once per (scalar, window) so
// that the histogram counters and the scatter targets of one window stay L2-resident.
ZK_HD uint32_t msm_digit_of_window(const uint32_t* s, uint32_t flip, const MsmPlan& p, int w, uint32_t* neg_out) {
    const uint32_t half = 1u << (p.c - 1);
    uint32_t carry = 0, d = 0, neg = 0;
    for (int i = 0; i <= w; ++i) {
        d = msm_scalar_bits(s, i * p.c, p.c) + carry;
        neg = 0;
        if (d > half) {
            d = (1u << p.c) - d;
            carry = 1;
            neg = 1;
        } else {
            carry = 0;
        }
    }
    *neg_out = neg ^ flip;
    return d;
}

// All W signed digits of one scalar in ONE walk (the carry propagates once).  Encoding of out[w]: MSM_DIGIT_NONE for a zero digit,
// else (|d| - 1) | sign << 31 -- i.e. the bucket index inside the window and the sign that goes into bit 31 of the sorted entry.
// Round 1 recomputed the digit of window w by walking windows 0..w in every (scalar, window) thread of both sort passes, after a
// fresh Montgomery conversion each time: 9 % of the prover's kernel time for a 32 B/term pass.
static constexpr uint32_t MSM_DIGIT_NONE = 0xffffffffu;
ZK_HD void msm_digits_all(const uint32_t* s, uint32_t flip, const MsmPlan& p, uint32_t* out, size_t out_stride) {
    const uint32_t half = 1u << (p.c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < p.W; ++w) {
        uint32_t d = msm_scalar_bits(s, w * p.c, p.c) + carry, neg = 0;
        if (d > half) {
            d = (1u << p.c) - d;
            carry = 1;
            neg = 1;
        } else {
            carry = 0;
        }
        out[(size_t)w * out_stride] = d ? ((d - 1) | ((neg ^ flip) << 31)) : MSM_DIGIT_NONE;
    }
}

// Small-scalar MSM (msm.cu, msm_small_window_sums): the whole scalar is one signed digit, same encoding as above.
ZK_HD uint32_t msm_small_digit(int32_t v) {
    if (v > 0) return (uint32_t)(v - 1);
    if (v < 0) return (uint32_t)(-(int64_t)v - 1) | 0x80000000u;
    return MSM_DIGIT_NONE;
}

// ---- 128-bit moves of points -----------------------------------------------------------------------------------------
template <class C>
ZK_HD Affine<C> msm_load_affine(const uint32_t* bases, uint32_t idx) {
    using Fq = typename Affine<C>::Fq;
    uint32_t w[24];
#if defined(__CUDA_ARCH__)
    const uint4* p = reinterpret_cast<const uint4*>(bases + (size_t)idx * 24);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        uint4 v = __ldg(p + k);
        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
    }
#else
    for (int k = 0; k < 24; ++k) w[k] = bases[(size_t)idx * 24 + k];
#endif
    Affine<C> r;
    r.x = Fq::unpack(w);
    r.y = Fq::unpack(w + 12);
    return r;
}
template <class C>
ZK_HD XYZZ<C> msm_load_xyzz(const XYZZ<C>* p) {
#if defined(__CUDA_ARCH__)
    static_assert(sizeof(XYZZ<C>) % 16 == 0, "XYZZ must be a whole number of 128-bit words");
    constexpr int Q = sizeof(XYZZ<C>) / 16;
    union { XYZZ<C> v; uint4 q[Q]; } u;
    const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int k = 0; k < Q; ++k) u.q[k] = s[k];
    return u.v;
#else
    return *p;
#endif
}
template <class C>
ZK_HD void msm_store_xyzz(XYZZ<C>* p, const XYZZ<C>& v) {
#if defined(__CUDA_ARCH__)
    constexpr int Q = sizeof(XYZZ<C>) / 16;
    union { XYZZ<C> v; uint4 q[Q]; } u;
    u.v = v;
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int k = 0; k < Q; ++k) d[k] = u.q[k];
#else
    *p = v;
#endif
}

// ---- asynchronous staging of the NEXT entry's point in shared memory (device only) -----------------------------------------------------
// north_star sketches "point windows staged into shared memory via TMA".  After the sort a bucket's points are a gather of 96-byte
// records, so the only thing to stage is the next record of each thread: while the ten products of entry `pos` run, the point of entry
// pos + 1 travels global -> shared without passing through registers, and the sorted index of entry pos + 2 is already in a register.
//   PF = 1: six 16-byte cp.async (LDGSTS) per point, cp.async.wait_group per thread
//   PF = 2: ONE 96-byte bulk copy (cp.async.bulk, the TMA engine) per point, completion on a per-thread mbarrier (tx-count 96)
// Shared memory per block of 128 threads: 2 stages x 96 B x 128 = 24 KB (+ 2 KB of mbarriers).  Measured: profiles/r2_quick_perf_prefetch.txt.
#if defined(__CUDA_ARCH__)
struct MsmStage {
    uint32_t pts;   // shared-space address of this thread's two 96-byte slots (slot s at pts + s * 96 * blockDim.x ... see msm_stage_slot)
    uint32_t bars;  // shared-space address of this thread's two mbarriers (PF = 2)
};
__device__ __forceinline__ uint32_t msm_stage_slot(const MsmStage& st, int s) { return st.pts + (uint32_t)s * 96u; }
template <int PF>
__device__ __forceinline__ void msm_stage_init(MsmStage& st, void* smem_pts, void* smem_bars) {
    // thread-private slots: [thread][stage][96 B] so that the bulk copy's destination is contiguous
    st.pts = (uint32_t)__cvta_generic_to_shared(smem_pts) + threadIdx.x * 192u;
    st.bars = (uint32_t)__cvta_generic_to_shared(smem_bars) + threadIdx.x * 16u;
    if (PF == 2) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st.bars));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st.bars + 8u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}
template <int PF>
__device__ __forceinline__ void msm_stage_issue(const MsmStage& st, int s, const uint32_t* bases, uint32_t idx) {
    const uint32_t dst = msm_stage_slot(st, s);
    const void* src = bases + (size_t)idx * 24;
    if (PF == 1) {
#pragma unroll
        for (int k = 0; k < 6; ++k)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * k), "l"((const char*)src + 16 * k) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        const uint32_t bar = st.bars + 8u * s;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this thread's earlier generic-proxy reads of the slot before the async-proxy write
        asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], 96; }" ::"r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 96, [%2];" ::"r"(dst), "l"(src), "r"(bar) : "memory");
    }
}
// waits for the copy into slot s (`pending` = copies issued after it that may still be in flight: 0 or 1), then reads the point
template <class C, int PF>
__device__ __forceinline__ Affine<C> msm_stage_take(const MsmStage& st, int s, uint32_t parity, int pending) {
    if (PF == 1) {
        if (pending)
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        else
            asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        const uint32_t bar = st.bars + 8u * s;
        asm volatile(
            "{ .reg .pred p;\n"
            "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@!p bra W; }" ::"r"(bar), "r"(parity) : "memory");
    }
    using Fq = typename Affine<C>::Fq;
    uint32_t w[24];
    const uint32_t src = msm_stage_slot(st, s);
#pragma unroll
    for (int k = 0; k < 6; ++k)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * k]), "=r"(w[4 * k + 1]), "=r"(w[4 * k + 2]), "=r"(w[4 * k + 3]) : "r"(src + 16u * k) : "memory");
    Affine<C> r;
    r.x = Fq::unpack(w);
    r.y = Fq::unpack(w + 12);
    return r;
}
#endif

// ---- accumulate: thread t owns sorted[t L, min((t+1) L, E)) ------------------------------------------------------------
// offsets has nb + 1 entries (offsets[nb] = E = number of sorted entries).  buckets must hold valid points (zeroed =
// infinity before the first chunk).  head / tail / tail_bucket have one slot per slice; tail_bucket[t] names the bucket
// whose partial sits in tail[t] (MSM_NO_BUCKET if none), which is all the merge pass needs.
static constexpr uint32_t MSM_NO_BUCKET = 0xffffffffu;
template <class C, bool CALL = false, int PF = 0>
ZK_HD void msm_slice_accumulate(uint32_t t, uint32_t n_slices, uint32_t L, const uint32_t* offsets, uint32_t nb, const uint32_t* sorted,
                                const uint32_t* bases, XYZZ<C>* buckets, XYZZ<C>* head, XYZZ<C>* tail, uint32_t* tail_bucket,
                                void* smem_pts = nullptr, void* smem_bars = nullptr) {
    (void)smem_pts;
    (void)smem_bars;
    if (t >= n_slices) return;
    const uint32_t E = offsets[nb];
    const uint64_t lo64 = (uint64_t)t * L;
    tail_bucket[t] = MSM_NO_BUCKET;
    if (lo64 >= E) return;
    const uint32_t lo = (uint32_t)lo64;
    const uint32_t hi = (E - lo > L) ? lo + L : E;
    // bucket of entry lo: the largest b with offsets[b] <= lo (then offsets[b + 1] > lo)
    uint32_t a = 0, z = nb;
    while (z - a > 1) {
        uint32_t mid = (a + z) >> 1;
        if (offsets[mid] <= lo) a = mid; else z = mid;
    }
    uint32_t b = a;
    uint32_t seg_end = offsets[b + 1];
    bool open_head = offsets[b] < lo;  // the bucket began in an earlier slice
    XYZZ<C> acc = open_head ? XYZZ<C>::inf() : msm_load_xyzz<C>(buckets + b);
#if defined(__CUDA_ARCH__)
    // software pipeline of the staged variants: e_cur = entry pos (its point is in flight to slot (pos - lo) & 1), e_next = entry pos + 1
    MsmStage stage;
    uint32_t e_cur = 0, e_next = 0;
    if (PF) {
        msm_stage_init<PF>(stage, smem_pts, smem_bars);
        e_cur = sorted ? sorted[lo] : lo;
        msm_stage_issue<PF>(stage, 0, bases, e_cur & 0x7fffffffu);
        if (lo + 1 < hi) e_next = sorted ? sorted[lo + 1] : lo + 1;
    }
#endif
    for (uint32_t pos = lo; pos < hi; ++pos) {
#if defined(__CUDA_ARCH__)
        Affine<C> staged;
        uint32_t e_staged = 0;
        if (PF) {
            const int slot = (int)((pos - lo) & 1u);
            const bool more = pos + 1 < hi;
            if (more) msm_stage_issue<PF>(stage, slot ^ 1, bases, e_next & 0x7fffffffu);  // slot ^ 1 was consumed in the previous iteration
            const uint32_t e_next2 = pos + 2 < hi ? (sorted ? sorted[pos + 2] : pos + 2) : 0u;
            staged = msm_stage_take<C, PF>(stage, slot, ((pos - lo) >> 1) & 1u, more ? 1 : 0);
            e_staged = e_cur;
            e_cur = e_next;
            e_next = e_next2;
        }
#endif
        if (pos == seg_end) {  // bucket b is complete: flush, move to the next non-empty bucket
            if (open_head) {
                msm_store_xyzz<C>(head + t, acc);
                open_head = false;
            } else {
                msm_store_xyzz<C>(buckets + b, acc);
            }
            // next non-empty bucket: probe a few, then binary search (a sparse chunk can leave millions of empty buckets
            // between two entries).  Terminates inside the array because offsets[nb] = E > pos.
            ++b;
            seg_end = offsets[b + 1];
            for (int probe = 0; probe < 4 && seg_end == pos; ++probe) {
                ++b;
                seg_end = offsets[b + 1];
            }
            if (seg_end == pos) {
                uint32_t a2 = b, z2 = nb;  // offsets[a2] <= pos < offsets[z2]
                while (z2 - a2 > 1) {
                    uint32_t mid = (a2 + z2) >> 1;
                    if (offsets[mid] <= pos) a2 = mid; else z2 = mid;
                }
                b = a2;
                seg_end = offsets[b + 1];
            }
            acc = msm_load_xyzz<C>(buckets + b);
        }
#if defined(__CUDA_ARCH__)
        const uint32_t e = PF ? e_staged : (sorted ? sorted[pos] : pos);
        Affine<C> pt = PF ? staged : msm_load_affine<C>(bases, e & 0x7fffffffu);
#else
        const uint32_t e = sorted ? sorted[pos] : pos;  // no index array: the points themselves are in bucket order (pair round output)
        Affine<C> pt = msm_load_affine<C>(bases, e & 0x7fffffffu);
#endif
        if (e >> 31) pt.y = pt.y.neg();
        if (CALL)
            acc.madd_call(pt);
        else
            acc.madd(pt);
    }
    if (open_head)
        msm_store_xyzz<C>(head + t, acc);       // the whole slice lies inside a bucket that began earlier
    else if (seg_end == hi)
        msm_store_xyzz<C>(buckets + b, acc);    // began here (or at lo) and ends exactly at the slice end
    else {
        msm_store_xyzz<C>(tail + t, acc);       // began here, continues in the next slice (includes the old bucket value)
        tail_bucket[t] = b;
    }
}

// ---- pair rounds (batched-affine first levels of the bucket sums) -----------------------------------------------------------
// Before the XYZZ accumulation the entries of every bucket are added in PAIRS as affine points, R times over: with the
// inverse of x1 - x0 in hand an affine addition is 3 field products (lambda, lambda^2, y) instead of the 10 of a mixed XYZZ
// addition, and Montgomery's trick shares one inversion over a whole launch (3 more products per pair).  The sort places
// every bucket at an offset that is a multiple of 2^R (buckets are padded with MSM_NONE), so in every round pair j is
// simply entries (2j, 2j+1), its sum lands at slot j of a dense array, and after R rounds the bucket offsets are the sorted
// offsets >> R.  Round 1 gathers SRS points through the sorted indices (sign in bit 31); later rounds read the previous
// round's sums, which are contiguous.  Per round:
//   A  msm_pair_products: thread t walks pairs [tG, tG+G): den_j (x1-x0, or 2y for a doubling, or 1 when nothing is to be
//      inverted), prefix[j] = product of the earlier den in the group, tprod[t] = product of the whole group.  In round 1
//      only the x coordinates are gathered (48 of the 96 bytes) unless the pair is degenerate.
//   B  msm_pair_invert  : thread u inverts G2 consecutive group products with one Fermat inversion (every lane busy)
//   C  msm_pair_add     : thread t walks its group backwards: 1/den_j = (running inverse) * prefix[j], then the addition
static constexpr uint32_t MSM_NONE = 0xffffffffu;
enum MsmPairKind { MSM_PAIR_COPY0 = 0, MSM_PAIR_COPY1 = 1, MSM_PAIR_ADD = 2, MSM_PAIR_DBL = 3, MSM_PAIR_INF = 4 };

template <class C>
ZK_HD Affine<C> msm_load_entry(const uint32_t* bases, uint32_t e) {
    Affine<C> pt = msm_load_affine<C>(bases, e & 0x7fffffffu);
    if (e >> 31) pt.y = pt.y.neg();
    return pt;
}
template <class C>
ZK_HD typename Affine<C>::Fq msm_load_x(const uint32_t* bases, uint32_t idx) {
    using Fq = typename Affine<C>::Fq;
    uint32_t w[12];
#if defined(__CUDA_ARCH__)
    const uint4* p = reinterpret_cast<const uint4*>(bases + (size_t)idx * 24);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint4 v = __ldg(p + k);
        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
    }
#else
    for (int k = 0; k < 12; ++k) w[k] = bases[(size_t)idx * 24 + k];
#endif
    return Fq::unpack(w);
}
// what is to be inverted for the pair (p0, p1); p1 is absent for the padding slot of an odd bucket
template <class C>
ZK_HD int msm_pair_den(const Affine<C>& p0, bool has1, const Affine<C>& p1, typename Affine<C>::Fq* den) {
    using Fq = typename Affine<C>::Fq;
    *den = Fq::one();
    if (!has1 || p1.is_inf()) return MSM_PAIR_COPY0;
    if (p0.is_inf()) return MSM_PAIR_COPY1;
    Fq dx = p1.x - p0.x;
    if (!dx.is_zero()) {
        *den = dx;
        return MSM_PAIR_ADD;
    }
    if (p0.y == p1.y && !p0.y.is_zero()) {
        *den = p0.y.dbl();
        return MSM_PAIR_DBL;
    }
    return MSM_PAIR_INF;  // P + (-P), or a point of order two doubled
}
// the two operands of pair j: through the sorted indices (round 1) or straight from the previous round's sums (idx == nullptr)
template <class C>
ZK_HD void msm_pair_operands(const uint32_t* idx, const uint32_t* pts, uint32_t j, Affine<C>* p0, Affine<C>* p1, bool* has1) {
    if (idx) {
        const uint32_t e0 = idx[2 * (size_t)j], e1 = idx[2 * (size_t)j + 1];
        *p0 = e0 != MSM_NONE ? msm_load_entry<C>(pts, e0) : Affine<C>::inf();  // (NONE, NONE): padding up to the 2^R alignment
        *has1 = e1 != MSM_NONE;
        *p1 = *has1 ? msm_load_entry<C>(pts, e1) : Affine<C>::inf();
    } else {
        *p0 = msm_load_affine<C>(pts, 2 * j);
        *p1 = msm_load_affine<C>(pts, 2 * j + 1);
        *has1 = true;
    }
}
// den_j for pass A.  Round 1 reads x coordinates only: a non-zero difference of two non-zero x is an ordinary addition (the
// point at infinity is stored as (0, 0)); anything else is classified on the full points, exactly as pass C will.
template <class C>
ZK_HD typename Affine<C>::Fq msm_pair_den_lazy(const uint32_t* idx, const uint32_t* pts, uint32_t j) {
    using Fq = typename Affine<C>::Fq;
    if (idx) {
        const uint32_t e0 = idx[2 * (size_t)j], e1 = idx[2 * (size_t)j + 1];
        if (e1 == MSM_NONE) return Fq::one();
        const Fq x0 = msm_load_x<C>(pts, e0 & 0x7fffffffu), x1 = msm_load_x<C>(pts, e1 & 0x7fffffffu);
        const Fq dx = x1 - x0;
        if (!dx.is_zero() && !x0.is_zero() && !x1.is_zero()) return dx;
    }
    Affine<C> p0, p1;
    bool has1;
    msm_pair_operands<C>(idx, pts, j, &p0, &p1, &has1);
    Fq den;
    msm_pair_den<C>(p0, has1, p1, &den);
    return den;
}
template <class C>
ZK_HD void msm_pair_products(uint32_t t, uint32_t G, uint32_t n_pairs, const uint32_t* idx, const uint32_t* pts,
                             typename Affine<C>::Fq* prefix, typename Affine<C>::Fq* tprod) {
    using Fq = typename Affine<C>::Fq;
    const uint64_t lo = (uint64_t)t * G;
    if (lo >= n_pairs) return;
    const uint32_t hi = n_pairs - lo > G ? (uint32_t)lo + G : n_pairs;
    Fq run = Fq::one();
    for (uint32_t j = (uint32_t)lo; j < hi; ++j) {
        const Fq den = msm_pair_den_lazy<C>(idx, pts, j);
        prefix[j] = run;
        run = run * den;
    }
    tprod[t] = run;
}
// in place: vals[i] -> 1 / vals[i] for this thread's G2 values; scratch holds the running prefixes
template <class Fq>
ZK_HD void msm_pair_invert(uint32_t u, uint32_t G2, uint32_t count, Fq* vals, Fq* scratch) {
    const uint64_t lo = (uint64_t)u * G2;
    if (lo >= count) return;
    const uint32_t hi = count - lo > G2 ? (uint32_t)lo + G2 : count;
    Fq run = Fq::one();
    for (uint32_t i = (uint32_t)lo; i < hi; ++i) {
        scratch[i] = run;
        run = run * vals[i];
    }
    Fq inv = run.inverse();
    for (uint32_t i = hi; i-- > (uint32_t)lo;) {
        const Fq v = vals[i];
        vals[i] = inv * scratch[i];
        inv = inv * v;
    }
}
template <class C>
ZK_HD void msm_pair_add(uint32_t t, uint32_t G, uint32_t n_pairs, const uint32_t* idx, const uint32_t* pts,
                        const typename Affine<C>::Fq* prefix, const typename Affine<C>::Fq* tinv, uint32_t* out /* n_pairs x 24 words */) {
    using Fq = typename Affine<C>::Fq;
    const uint64_t lo = (uint64_t)t * G;
    if (lo >= n_pairs) return;
    const uint32_t hi = n_pairs - lo > G ? (uint32_t)lo + G : n_pairs;
    Fq inv = tinv[t];  // 1 / (product of the group's den)
    for (uint32_t j = hi; j-- > (uint32_t)lo;) {
        Affine<C> p0, p1;
        bool has1;
        msm_pair_operands<C>(idx, pts, j, &p0, &p1, &has1);
        Fq den;
        const int kind = msm_pair_den<C>(p0, has1, p1, &den);
        const Fq dinv = inv * prefix[j];  // 1 / den_j
        inv = inv * den;
        Affine<C> r;
        if (kind == MSM_PAIR_ADD || kind == MSM_PAIR_DBL) {
            Fq lam;
            if (kind == MSM_PAIR_ADD) {
                lam = (p1.y - p0.y) * dinv;
            } else {
                Fq x2 = p0.x.sqr();
                lam = (x2.dbl() + x2) * dinv;
            }
            r.x = lam.sqr() - p0.x - p1.x;
            r.y = lam * (p0.x - r.x) - p0.y;
        } else if (kind == MSM_PAIR_COPY0) {
            r = p0;
        } else if (kind == MSM_PAIR_COPY1) {
            r = p1;
        } else {
            r = Affine<C>::inf();
        }
        uint32_t w[24];
        r.x.pack(w);
        r.y.pack(w + 12);
#if defined(__CUDA_ARCH__)
        uint4* o = reinterpret_cast<uint4*>(out + (size_t)j * 24);
#pragma unroll
        for (int k = 0; k < 6; ++k) o[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
#else
        for (int k = 0; k < 24; ++k) out[(size_t)j * 24 + k] = w[k];
#endif
    }
}

// ---- merge: the bucket that began in slice t and ran past its end = tail[t] + the heads of the slices it continues into.
// Usually that is one head, handled by one thread per slice (msm_merge_slice).  A heavy bucket (the partly filled top
// window: 2^10 buckets share all the entries; or many equal scalars) continues through hundreds of slices: one thread adding
// them one after the other took 35 ms per MSM, one warp per slice for ALL slices 34 ms (latency-bound at 8 warps per SM;
// profiles/r1_launches_msm_2p26_pair_round.txt, r1_launches_msm_2p26_warp_merge.txt).  So: chains longer than
// MSM_MERGE_SERIAL_MAX are left to a second kernel with one warp per slice in which lane l sums the heads t+1+l, t+1+l+32, ...
// and the 32 lane sums are folded (shared-memory tree in the kernel, a plain loop in the CPU emulation).
static constexpr uint32_t MSM_MERGE_SERIAL_MAX = 32;
template <class C>
ZK_HD void msm_merge_slice(uint32_t t, uint32_t n_slices, uint32_t L, const uint32_t* offsets, XYZZ<C>* buckets, const XYZZ<C>* head,
                           const XYZZ<C>* tail, const uint32_t* tail_bucket) {
    if (t >= n_slices) return;
    const uint32_t b = tail_bucket[t];
    if (b == MSM_NO_BUCKET) return;
    const uint32_t t1 = (offsets[b + 1] - 1) / L;  // last slice the bucket reaches (> t)
    if (t1 - t > MSM_MERGE_SERIAL_MAX) return;     // msm_merge_lane's job
    XYZZ<C> acc = msm_load_xyzz<C>(tail + t);
    for (uint32_t u = t + 1; u <= t1; ++u) acc.add(msm_load_xyzz<C>(head + u));
    msm_store_xyzz<C>(buckets + b, acc);
}
template <class C>
ZK_HD bool msm_merge_lane(uint32_t t, uint32_t lane, uint32_t n_slices, uint32_t L, const uint32_t* offsets, const XYZZ<C>* head,
                          const XYZZ<C>* tail, const uint32_t* tail_bucket, XYZZ<C>* lane_sum, uint32_t* bucket) {
    if (t >= n_slices) return false;
    const uint32_t b = tail_bucket[t];
    if (b == MSM_NO_BUCKET) return false;
    const uint32_t t1 = (offsets[b + 1] - 1) / L;
    if (t1 - t <= MSM_MERGE_SERIAL_MAX) return false;
    *bucket = b;
    XYZZ<C> acc = lane == 0 ? msm_load_xyzz<C>(tail + t) : XYZZ<C>::inf();
    for (uint64_t u = (uint64_t)t + 1 + lane; u <= t1; u += 32) acc.add(msm_load_xyzz<C>(head + u));
    *lane_sum = acc;
    return true;
}

// ---- reduce: segment `seg_id` of window w covers buckets [seg_id * seg, ...) ------------------------------------------------
template <class C>
ZK_HD_COLD XYZZ<C> msm_mul_small(const XYZZ<C>& p, uint32_t k) {
    XYZZ<C> r = XYZZ<C>::inf();
    if (!k) return r;
    int top = 31;
    while (!((k >> top) & 1)) --top;
    for (int b = top; b >= 0; --b) {
        r = r.dbl();
        if ((k >> b) & 1) r.add(p);
    }
    return r;
}
// sum_{j in segment} (j + 1) B_j
template <class C>
ZK_HD XYZZ<C> msm_reduce_segment(const XYZZ<C>* B, uint32_t nbw, uint32_t seg, uint32_t seg_id) {
    const uint64_t lo64 = (uint64_t)seg_id * seg;
    XYZZ<C> contrib = XYZZ<C>::inf();
    if (lo64 >= nbw) return contrib;
    const uint32_t lo = (uint32_t)lo64;
    const uint32_t hi = nbw - lo > seg ? lo + seg : nbw;
    XYZZ<C> running = XYZZ<C>::inf(), acc = XYZZ<C>::inf();
    for (uint32_t j = hi; j-- > lo;) {
        running.add(msm_load_xyzz<C>(B + j));
        acc.add(running);
    }
    // sum (j + 1) B_j = acc + lo * running      (acc = sum (j - lo + 1) B_j)
    contrib = msm_mul_small<C>(running, lo);
    contrib.add(acc);
    return contrib;
}

// ---- plan ---------------------------------------------------------------------------------------------------------------
// Window size: minimise  W * (1.15 n_local + 5 * 2^(c-1))  -- per window n_local mixed additions in the accumulation plus
// the sort passes (~0.15 of a mixed addition per entry), and two full additions per bucket in the reduction (measured
// 4-5 mixed-addition times per bucket, profiles/r1_launches_msm_2p26_slices.txt, r2_launches_bench_4k.txt) -- over c <= c_max.
// c_max bounds the bucket array (2^(c-1) W XYZZ points: 4.8 GB at c = 22, 8.9 GB at c = 23).  253 = 11 x 23: from ~2^27 terms on
// the plan is 11 windows of 23 bits instead of 12 of 22 -- the 4 KiB proof went 14.94 -> 14.60 s with the cap raised from 22 to 23
// (same proof bytes; gpurun_out/j3_bench_c2{2,3}.json).  n_local = points per rank (the plan is a function of the global
// (n, nranks) so that every rank derives the same windows).
inline MsmPlan msm_make_plan(size_t n, int fr_bits, int forced_c, int nranks = 1, int c_max = 23) {
    MsmPlan p;
    int c = forced_c;
    if (c <= 0) {
        const double n_local = (double)((n + (size_t)nranks - 1) / (size_t)nranks);
        double best = 0;
        for (int cc = 3; cc <= c_max; ++cc) {
            const int W = (fr_bits + cc - 1) / cc;
            const double cost = W * (1.15 * n_local + 5.0 * (double)((size_t)1 << (cc - 1)));
            if (c <= 0 || cost < best) {
                best = cost;
                c = cc;
            }
        }
    }
    p.c = c;
    p.W = (fr_bits + c - 1) / c;
    p.nbw = 1u << (c - 1);
    p.nb = p.nbw * (uint32_t)p.W;
    return p;
}

}  // namespace zk
