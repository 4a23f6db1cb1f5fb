// Host test driver for DevArena (csrc/common.cuh), built on demand by tests/test_arena.py: random allocate / free traffic,
// checked for overlap, bounds, alignment, accounting and full coalescing.  Test infrastructure only.
#include <cstdint>
#include <cstdlib>
#include <map>
#include <vector>

#include "../aes_zero_knowledge_proof_circuit_b200/csrc/common.cuh"

extern "C" int arena_selftest(uint64_t seed, int rounds) {
    const size_t SIZE = (size_t)1 << 30;
    char* base = reinterpret_cast<char*>((uintptr_t)1 << 40);  // never dereferenced
    DevArena a;
    a.reset(base, SIZE);
    std::map<size_t, size_t> livemap;  // offset -> rounded length
    std::vector<std::pair<void*, size_t>> held;
    uint64_t x = seed * 6364136223846793005ull + 1442695040888963407ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    size_t live = 0;
    for (int r = 0; r < rounds; ++r) {
        const bool do_alloc = held.empty() || (rnd() % 3 != 0);
        if (do_alloc) {
            size_t n = 1 + rnd() % ((size_t)48 << 20);
            void* p = a.alloc(n);
            if (!p) {  // must only fail when no free segment is large enough
                for (auto& s : a.free_segs)
                    if (s.second >= DevArena::round_up(n)) return -1;
                continue;
            }
            size_t off = (size_t)((char*)p - base), len = DevArena::round_up(n);
            if (!a.owns(p) || off % 512 || off + len > SIZE) return -2;
            auto nx = livemap.lower_bound(off);
            if (nx != livemap.end() && nx->first < off + len) return -3;                                  // overlaps the next block
            if (nx != livemap.begin() && std::prev(nx)->first + std::prev(nx)->second > off) return -4;   // overlaps the previous block
            livemap[off] = len;
            held.emplace_back(p, n);
            live += len;
        } else {
            size_t i = rnd() % held.size();
            a.free(held[i].first, held[i].second);
            size_t off = (size_t)((char*)held[i].first - base);
            live -= livemap[off];
            livemap.erase(off);
            held[i] = held.back();
            held.pop_back();
        }
        if (a.live != live) return -5;
        size_t free_total = 0, prev_end = (size_t)-1;
        for (auto& s : a.free_segs) {
            if (s.first == prev_end) return -6;  // adjacent free segments must have been merged
            prev_end = s.first + s.second;
            free_total += s.second;
        }
        if (free_total + live != SIZE) return -7;
    }
    for (auto& h : held) a.free(h.first, h.second);
    if (a.live != 0 || a.free_segs.size() != 1 || a.free_segs.begin()->first != 0 || a.free_segs.begin()->second != SIZE) return -8;
    return 0;
}
