// Host-side randomness of the Marlin prover: Blake2s, ChaCha20, the Fiat-Shamir rng and ark-ff's field sampling.
//
// Follows the pinned third-party crates the reference reaches through simpleworks::marlin::generate_proof
// (src/lib.rs:111):
//   blake2 0.9.2            Blake2s-256, unkeyed
//   rand_chacha 0.3.1       ChaCha20Rng: key = seed, 64-bit block counter (words 12-13), stream 0; BlockRng next_u64
//   ark-marlin 0.3.0 rng.rs FiatShamirRng<Blake2s>: seed = H(bytes); absorb: seed' = H(bytes || seed)
//   ark-ff 0.3.0            UniformRand for Fp256: four u64, clear the top REPR_SHAVE_BITS, accept if < modulus; the
//                           accepted words are used AS the Montgomery representation
// Only u64-granular draws happen on this path, so the generator is a stream of u64 (two consecutive ChaCha words).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace zk {

// ---- Blake2s-256 (RFC 7693), unkeyed ---------------------------------------------------------------------------------
struct Blake2s {
    uint32_t h[8];
    uint8_t buf[64];
    size_t buflen = 0;
    uint64_t t = 0;
    static constexpr uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    Blake2s() {
        for (int i = 0; i < 8; ++i) h[i] = IV[i];
        h[0] ^= 0x01010000u ^ 32u;
    }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void compress(const uint8_t* block, bool last) {
        static const uint8_t S[10][16] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
                                          {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
                                          {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
                                          {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
                                          {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
        uint32_t m[16], v[16];
        for (int i = 0; i < 16; ++i) memcpy(&m[i], block + 4 * i, 4);
        for (int i = 0; i < 8; ++i) {
            v[i] = h[i];
            v[i + 8] = IV[i];
        }
        v[12] ^= (uint32_t)t;
        v[13] ^= (uint32_t)(t >> 32);
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 12);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 8);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 7);
        };
        for (int r = 0; r < 10; ++r) {
            const uint8_t* s = S[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const uint8_t* p, size_t n) {
        while (n) {
            if (buflen == 64) {  // keep the last block for finalisation
                t += 64;
                compress(buf, false);
                buflen = 0;
            }
            size_t k = 64 - buflen < n ? 64 - buflen : n;
            memcpy(buf + buflen, p, k);
            buflen += k;
            p += k;
            n -= k;
        }
    }
    void finish(uint8_t out[32]) {
        t += buflen;
        memset(buf + buflen, 0, 64 - buflen);
        compress(buf, true);
        memcpy(out, h, 32);
    }
    static void digest(const std::vector<uint8_t>& data, uint8_t out[32]) {
        Blake2s b;
        b.update(data.data(), data.size());
        b.finish(out);
    }
};

// ---- ChaCha20 keystream as a stream of u64 ----------------------------------------------------------------------------
inline void chacha20_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
    uint32_t x[16];
    memcpy(x, s, sizeof(s));
    auto rotl = [](uint32_t v, int n) { return (v << n) | (v >> (32 - n)); };
    auto QR = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    };
    for (int i = 0; i < 10; ++i) {
        QR(0, 4, 8, 12); QR(1, 5, 9, 13); QR(2, 6, 10, 14); QR(3, 7, 11, 15);
        QR(0, 5, 10, 15); QR(1, 6, 11, 12); QR(2, 7, 8, 13); QR(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}

struct ChaCha20Rng {
    uint32_t key[8];
    uint64_t pos = 0;  // in u64 units; block = pos / 8
    uint32_t cur[16];
    uint64_t cur_block = ~0ull;
    explicit ChaCha20Rng(const uint8_t seed[32]) { memcpy(key, seed, 32); }
    uint64_t next_u64() {
        uint64_t blk = pos >> 3;
        if (blk != cur_block) {
            chacha20_block(key, blk, cur);
            cur_block = blk;
        }
        int i = (int)(pos & 7);
        ++pos;
        return (uint64_t)cur[2 * i] | ((uint64_t)cur[2 * i + 1] << 32);
    }
};

// ark-ff UniformRand for a 4-limb field: raw limbs (Montgomery representation) in out[4]
template <class FrP>
inline void fr_rand_raw(ChaCha20Rng& rng, uint64_t out[4]) {
    constexpr int shave = 256 - FrP::BITS;
    uint64_t mod[4];
    for (int i = 0; i < 4; ++i) mod[i] = (uint64_t)FrP::MOD(2 * i) | ((uint64_t)FrP::MOD(2 * i + 1) << 32);
    for (;;) {
        for (int i = 0; i < 4; ++i) out[i] = rng.next_u64();
        out[3] &= ~0ull >> shave;
        bool lt = false;
        for (int i = 3; i >= 0; --i) {
            if (out[i] != mod[i]) {
                lt = out[i] < mod[i];
                break;
            }
        }
        if (lt) return;
    }
}

struct FiatShamirRng {
    uint8_t seed[32];
    ChaCha20Rng rng;
    explicit FiatShamirRng(const std::vector<uint8_t>& bytes) : rng(init(bytes, seed)) {}
    static const uint8_t* init(const std::vector<uint8_t>& bytes, uint8_t* seed_out) {
        Blake2s::digest(bytes, seed_out);
        return seed_out;
    }
    void absorb(const std::vector<uint8_t>& bytes) {
        std::vector<uint8_t> b(bytes);
        b.insert(b.end(), seed, seed + 32);
        Blake2s::digest(b, seed);
        rng = ChaCha20Rng(seed);
    }
};

}  // namespace zk
