// CPU emulation of the MSM kernel pipeline (test infrastructure, built on demand by tests/test_msm_emul.py).
// It runs the SAME per-thread bodies the CUDA kernels run (aes_zero_knowledge_proof_circuit_b200/csrc/msm_core.cuh),
// one "thread" after another, so that the slice / head / tail / merge / reduce index bookkeeping is checked against the
// oracle in the CPU-only test tier.  It is not part of the product and is never used as a fallback.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../aes_zero_knowledge_proof_circuit_b200/csrc/msm_core.cuh"

using namespace zk;

template <class C>
static int emul(const uint32_t* bases, const uint32_t* scalars, size_t n, int forced_c, uint32_t L, size_t chunk, uint32_t seg, uint32_t* out96) {
    using FrP = typename C::FrP;
    MsmPlan p = msm_make_plan(n ? n : 1, FrP::BITS, forced_c);
    std::vector<XYZZ<C>> buckets(p.nb, XYZZ<C>::inf());
    std::vector<uint32_t> counts(p.nb + 1), offsets(p.nb + 1);
    for (size_t base = 0; base < n; base += chunk) {
        size_t m = n - base < chunk ? n - base : chunk;
        std::fill(counts.begin(), counts.end(), 0u);
        for (size_t i = 0; i < m; ++i) {
            uint32_t s[8];
            memcpy(s, scalars + 8 * (base + i), 32);
            uint32_t flip = msm_fold_scalar<FrP>(s);
            std::vector<uint32_t> dig(p.W);
            msm_digits_all(s, flip, p, dig.data(), 1);  // the kernels' one-walk digits must equal the per-window definition
            for (int w = 0; w < p.W; ++w) {
                uint32_t neg, d = msm_digit_of_window(s, flip, p, w, &neg);
                if ((d ? ((d - 1) | (neg << 31)) : MSM_DIGIT_NONE) != dig[w]) return -1;
                if (d) counts[(uint32_t)w * p.nbw + d - 1]++;
            }
        }
        uint32_t run = 0;
        for (size_t k = 0; k <= p.nb; ++k) {
            offsets[k] = run;
            run += counts[k];
        }
        std::vector<uint32_t> sorted(run ? run : 1);
        std::fill(counts.begin(), counts.end(), 0u);
        for (size_t i = 0; i < m; ++i) {
            uint32_t s[8];
            memcpy(s, scalars + 8 * (base + i), 32);
            uint32_t flip = msm_fold_scalar<FrP>(s);
            for (int w = 0; w < p.W; ++w) {
                uint32_t neg, d = msm_digit_of_window(s, flip, p, w, &neg);
                if (!d) continue;
                uint32_t key = (uint32_t)w * p.nbw + d - 1;
                sorted[offsets[key] + counts[key]++] = (uint32_t)(base + i) | (neg << 31);
            }
        }
        size_t slices = (m * (size_t)p.W + L - 1) / L;  // the same upper bound the host code launches
        std::vector<XYZZ<C>> head(slices + 1), tail(slices + 1);
        std::vector<uint32_t> tail_bucket(slices + 1, 0x12345678u);
        // poison the partial arrays: a slot that is read without having been written shows up as a wrong result
        memset((void*)head.data(), 0x5a, sizeof(XYZZ<C>) * head.size());
        memset((void*)tail.data(), 0x5a, sizeof(XYZZ<C>) * tail.size());
        for (size_t t = 0; t < slices; ++t)
            msm_slice_accumulate<C>((uint32_t)t, (uint32_t)slices, L, offsets.data(), p.nb, sorted.data(), bases, buckets.data(), head.data(), tail.data(),
                                    tail_bucket.data());
        for (size_t t = 0; t < slices + 1; ++t)  // one thread / warp past the end, as a partly filled last block launches
            msm_merge_slice<C>((uint32_t)t, (uint32_t)slices, L, offsets.data(), buckets.data(), head.data(), tail.data(), tail_bucket.data());
        for (size_t t = 0; t < slices + 1; ++t) {
            XYZZ<C> sum = XYZZ<C>::inf(), lane_sum;
            uint32_t b = 0;
            bool any = false;
            for (uint32_t lane = 0; lane < 32; ++lane)
                if (msm_merge_lane<C>((uint32_t)t, lane, (uint32_t)slices, L, offsets.data(), head.data(), tail.data(), tail_bucket.data(), &lane_sum, &b)) {
                    sum.add(lane_sum);
                    any = true;
                }
            if (any) buckets[b] = sum;
        }
    }
    XYZZ<C> total = XYZZ<C>::inf();
    for (int w = p.W - 1; w >= 0; --w) {
        for (int k = 0; k < p.c; ++k) total = total.dbl();
        uint32_t nseg = (p.nbw + seg - 1) / seg + 1;  // one segment past the end: must contribute nothing
        for (uint32_t sid = 0; sid < nseg; ++sid) total.add(msm_reduce_segment<C>(buckets.data() + (size_t)w * p.nbw, p.nbw, seg, sid));
    }
    Affine<C> r = total.to_affine();
    memcpy(out96, r.x.v, 48);
    memcpy(out96 + 12, r.y.v, 48);
    return p.c;
}

// The small-scalar pipeline (msm.cu, msm_small_window_sums): one signed digit per term, S pseudo-windows of m = ceil(n / S) consecutive terms with
// 2^(c-1) buckets each, the entry of a term is its own index, slice accumulation / merges over the whole entry array, one bucket per reduction
// segment, and the PLAIN sum of the S window sums.
template <class C>
static int emul_small(const uint32_t* bases, const int32_t* vals, size_t n, size_t start, size_t stride, int c, int S, uint32_t L, uint32_t* out96) {
    MsmPlan p;
    p.c = c;
    p.W = S;
    p.nbw = 1u << (c - 1);
    p.nb = p.nbw * (uint32_t)S;
    const size_t m = (n + (size_t)S - 1) / (size_t)S, padded = m * (size_t)S;
    std::vector<uint32_t> digits(padded ? padded : 1), counts(p.nb + 1, 0u), offsets(p.nb + 1);
    for (size_t j = 0; j < padded; ++j) digits[j] = j < n ? msm_small_digit(vals[start + j * stride]) : MSM_DIGIT_NONE;
    for (size_t j = 0; j < padded; ++j)
        if (digits[j] != MSM_DIGIT_NONE) {
            if ((digits[j] & 0x7fffffffu) >= p.nbw) return -2;  // value wider than the digit
            counts[(uint32_t)(j / m) * p.nbw + (digits[j] & 0x7fffffffu)]++;
        }
    uint32_t run = 0;
    for (size_t k = 0; k <= p.nb; ++k) {
        offsets[k] = run;
        run += counts[k];
    }
    std::vector<uint32_t> sorted(run ? run : 1);
    std::fill(counts.begin(), counts.end(), 0u);
    for (size_t j = 0; j < padded; ++j)
        if (digits[j] != MSM_DIGIT_NONE) {
            const uint32_t key = (uint32_t)(j / m) * p.nbw + (digits[j] & 0x7fffffffu);
            sorted[offsets[key] + counts[key]++] = (uint32_t)j | (digits[j] & 0x80000000u);
        }
    std::vector<XYZZ<C>> buckets(p.nb, XYZZ<C>::inf());
    const size_t slices = (padded + L - 1) / L;
    std::vector<XYZZ<C>> head(slices + 1), tail(slices + 1);
    std::vector<uint32_t> tail_bucket(slices + 1, 0x12345678u);
    memset((void*)head.data(), 0x5a, sizeof(XYZZ<C>) * head.size());
    memset((void*)tail.data(), 0x5a, sizeof(XYZZ<C>) * tail.size());
    for (size_t t = 0; t < slices; ++t)
        msm_slice_accumulate<C>((uint32_t)t, (uint32_t)slices, L, offsets.data(), p.nb, sorted.data(), bases, buckets.data(), head.data(), tail.data(),
                                tail_bucket.data());
    for (size_t t = 0; t < slices + 1; ++t)
        msm_merge_slice<C>((uint32_t)t, (uint32_t)slices, L, offsets.data(), buckets.data(), head.data(), tail.data(), tail_bucket.data());
    for (size_t t = 0; t < slices + 1; ++t) {
        XYZZ<C> sum = XYZZ<C>::inf(), lane_sum;
        uint32_t b = 0;
        bool any = false;
        for (uint32_t lane = 0; lane < 32; ++lane)
            if (msm_merge_lane<C>((uint32_t)t, lane, (uint32_t)slices, L, offsets.data(), head.data(), tail.data(), tail_bucket.data(), &lane_sum, &b)) {
                sum.add(lane_sum);
                any = true;
            }
        if (any) buckets[b] = sum;
    }
    XYZZ<C> total = XYZZ<C>::inf();
    for (int w = 0; w < S; ++w)
        for (uint32_t sid = 0; sid < p.nbw; ++sid) total.add(msm_reduce_segment<C>(buckets.data() + (size_t)w * p.nbw, p.nbw, 1, sid));
    Affine<C> r = total.to_affine();
    memcpy(out96, r.x.v, 48);
    memcpy(out96 + 12, r.y.v, 48);
    return 1;
}
extern "C" int msm_emul_small(int curve, const uint32_t* bases, const int32_t* vals, size_t n, size_t start, size_t stride, int c, int S, uint32_t L,
                              uint32_t* out96) {
    if (curve == 377) return emul_small<G1_377Params>(bases, vals, n, start, stride, c, S, L, out96);
    if (curve == 381) return emul_small<G1_381Params>(bases, vals, n, start, stride, c, S, L, out96);
    return -1;
}

// The pair-round pipeline (per window: 2^R-aligned sort, R rounds of pair products / inversion / affine additions, then the
// slice accumulation over the pair sums), as msm.cu's msm_accumulate_paired sequences it.
template <class C>
static int emul_paired(const uint32_t* bases, const uint32_t* scalars, size_t n, int forced_c, uint32_t L, size_t chunk, uint32_t G, uint32_t G2,
                       int R, uint32_t* out96) {
    using FrP = typename C::FrP;
    using Fq = typename Affine<C>::Fq;
    MsmPlan p = msm_make_plan(n ? n : 1, FrP::BITS, forced_c);
    std::vector<XYZZ<C>> buckets(p.nb, XYZZ<C>::inf());
    const uint32_t align = 1u << R;
    for (size_t base = 0; base < n; base += chunk) {
        size_t m = n - base < chunk ? n - base : chunk;
        for (int w = 0; w < p.W; ++w) {
            std::vector<uint32_t> counts(p.nbw + 1, 0), off2(p.nbw + 1), poff(p.nbw + 1);
            auto digit = [&](size_t i, uint32_t* neg) {
                uint32_t s[8];
                memcpy(s, scalars + 8 * (base + i), 32);
                uint32_t flip = msm_fold_scalar<FrP>(s);
                return msm_digit_of_window(s, flip, p, w, neg);
            };
            for (size_t i = 0; i < m; ++i) {
                uint32_t neg, d = digit(i, &neg);
                if (d) counts[d - 1]++;
            }
            uint32_t run = 0;
            for (size_t k = 0; k <= p.nbw; ++k) {
                off2[k] = run;
                run += (counts[k] + align - 1) & ~(align - 1);
                poff[k] = off2[k] >> R;
            }
            const uint32_t n_entries = off2[p.nbw];
            std::vector<uint32_t> sorted2((size_t)n_entries + 2, MSM_NONE);
            std::fill(counts.begin(), counts.end(), 0u);
            for (size_t i = 0; i < m; ++i) {
                uint32_t neg, d = digit(i, &neg);
                if (d) sorted2[off2[d - 1] + counts[d - 1]++] = (uint32_t)(base + i) | (neg << 31);
            }
            std::vector<uint32_t> sums[2];
            const uint32_t* pts = bases;
            const uint32_t* idx = sorted2.data();
            uint32_t n_pairs = n_entries;
            for (int r = 1; r <= R; ++r) {
                n_pairs >>= 1;
                const uint32_t T = (n_pairs + G - 1) / G;
                std::vector<Fq> prefix(n_pairs + 1), tprod(T + 1), scratch(T + 1);
                memset((void*)prefix.data(), 0x5a, sizeof(Fq) * prefix.size());
                std::vector<uint32_t>& out = sums[r & 1];
                out.assign((size_t)24 * (n_pairs + 1), 0x5a5a5a5au);
                for (uint32_t t = 0; t < T + 1; ++t) msm_pair_products<C>(t, G, n_pairs, idx, pts, prefix.data(), tprod.data());
                for (uint32_t u = 0; u < (T + G2 - 1) / G2 + 1; ++u) msm_pair_invert<Fq>(u, G2, T, tprod.data(), scratch.data());
                for (uint32_t t = 0; t < T + 1; ++t) msm_pair_add<C>(t, G, n_pairs, idx, pts, prefix.data(), tprod.data(), out.data());
                pts = out.data();
                idx = nullptr;
            }
            size_t slices = ((size_t)n_pairs + L - 1) / L + 1;
            std::vector<XYZZ<C>> head(slices + 1), tail(slices + 1);
            std::vector<uint32_t> tail_bucket(slices + 1, 0x12345678u);
            memset((void*)head.data(), 0x5a, sizeof(XYZZ<C>) * head.size());
            memset((void*)tail.data(), 0x5a, sizeof(XYZZ<C>) * tail.size());
            XYZZ<C>* B = buckets.data() + (size_t)w * p.nbw;
            for (size_t t = 0; t < slices; ++t)
                msm_slice_accumulate<C>((uint32_t)t, (uint32_t)slices, L, poff.data(), p.nbw, nullptr, pts, B, head.data(), tail.data(), tail_bucket.data());
            for (size_t t = 0; t < slices + 1; ++t)
                msm_merge_slice<C>((uint32_t)t, (uint32_t)slices, L, poff.data(), B, head.data(), tail.data(), tail_bucket.data());
            for (size_t t = 0; t < slices + 1; ++t) {
                XYZZ<C> sum = XYZZ<C>::inf(), lane_sum;
                uint32_t b = 0;
                bool any = false;
                for (uint32_t lane = 0; lane < 32; ++lane)
                    if (msm_merge_lane<C>((uint32_t)t, lane, (uint32_t)slices, L, poff.data(), head.data(), tail.data(), tail_bucket.data(), &lane_sum, &b)) {
                        sum.add(lane_sum);
                        any = true;
                    }
                if (any) B[b] = sum;
            }
        }
    }
    XYZZ<C> total = XYZZ<C>::inf();
    for (int w = p.W - 1; w >= 0; --w) {
        for (int k = 0; k < p.c; ++k) total = total.dbl();
        for (uint32_t sid = 0; sid < (p.nbw + 63) / 64; ++sid) total.add(msm_reduce_segment<C>(buckets.data() + (size_t)w * p.nbw, p.nbw, 64, sid));
    }
    Affine<C> r = total.to_affine();
    memcpy(out96, r.x.v, 48);
    memcpy(out96 + 12, r.y.v, 48);
    return p.c;
}
extern "C" int msm_emul_paired(int curve, const uint32_t* bases, const uint32_t* scalars, size_t n, int forced_c, uint32_t L, size_t chunk, uint32_t G,
                               uint32_t G2, int R, uint32_t* out96) {
    if (curve == 377) return emul_paired<G1_377Params>(bases, scalars, n, forced_c, L, chunk, G, G2, R, out96);
    if (curve == 381) return emul_paired<G1_381Params>(bases, scalars, n, forced_c, L, chunk, G, G2, R, out96);
    return -1;
}

extern "C" int msm_emul(int curve, const uint32_t* bases, const uint32_t* scalars, size_t n, int forced_c, uint32_t L, size_t chunk, uint32_t seg,
                        uint32_t* out96) {
    if (curve == 377) return emul<G1_377Params>(bases, scalars, n, forced_c, L, chunk, seg, out96);
    if (curve == 381) return emul<G1_381Params>(bases, scalars, n, forced_c, L, chunk, seg, out96);
    return -1;
}
extern "C" void msm_plan(size_t n, int fr_bits, int forced_c, int nranks, int c_max, int* out) {
    MsmPlan p = msm_make_plan(n, fr_bits, forced_c, nranks, c_max);
    out[0] = p.c; out[1] = p.W; out[2] = (int)p.nbw;
}
