// Host-side check of bonsai_b200/csrc/bns_host_util.h (the loader's value dictionary) against sort + unique.
#include <cstdio>
#include <random>
#include "../../bonsai_b200/csrc/bns_host_util.h"

int main() {
    std::mt19937_64 rng(11);
    int failures = 0;
    struct Case { const char *name; uint64_t nb; int fill_pct; uint32_t n_small, n_big; };
    const Case cases[] = {{"empty", 0, 0, 0, 0}, {"tiny4", 4, 75, 3, 0}, {"tiny15", 15, 100, 2, 2}, {"one_word", 16, 50, 4, 1},
                          {"taxids", 1u << 20, 63, 6, 0}, {"many_small", 1u << 20, 77, 200000, 0}, {"big_values", 1u << 18, 60, 10, 5000},
                          {"all_big", 1u << 16, 99, 0, 70000}, {"odd_size", (1u << 16) + 7, 50, 100, 100}};
    for(const Case &c : cases) {
        std::vector<uint32_t> small(c.n_small), big(c.n_big);
        for(auto &v : small) v = (uint32_t)(rng() % (1u << 24));
        for(auto &v : big) v = (uint32_t)((1u << 24) + rng() % (0xffffffffull - (1u << 24) + 1));
        if(c.n_big) big[0] = 0xffffffffu;
        if(c.n_small) small[0] = 0;
        if(c.n_small > 1) small[1] = (1u << 24) - 1;
        std::vector<uint32_t> flags(c.nb < 16 ? 1 : (c.nb + 15) >> 4, 0xaaaaaaaau), vals(c.nb, 0xdeadbeefu), expect, dense;
        uint64_t n_occ = 0;
        for(uint64_t i = 0; i < c.nb; ++i) {
            const unsigned r = (unsigned)(rng() % 100);
            if((int)r < c.fill_pct && c.n_small + c.n_big) {
                flags[i >> 4] &= ~(3u << ((i & 15) << 1));
                const uint64_t pick = rng() % (c.n_small + c.n_big);
                vals[i] = pick < c.n_small ? small[pick] : big[pick - c.n_small];
                expect.push_back(vals[i]); dense.push_back(vals[i]);
                ++n_occ;
            } else if(r % 7 == 0) flags[i >> 4] = (flags[i >> 4] & ~(3u << ((i & 15) << 1))) | (1u << ((i & 15) << 1));   // deleted
        }
        std::sort(expect.begin(), expect.end());
        expect.erase(std::unique(expect.begin(), expect.end()), expect.end());
        uint64_t got_occ = ~0ull;
        const auto a = bns::distinct_values_khash(vals.data(), flags.data(), c.nb, &got_occ);
        const auto b = bns::distinct_values(dense.data(), dense.size());
        const bool ok = a == expect && b == expect && got_occ == n_occ;
        printf("%s buckets=%llu occupied=%llu distinct=%zu %s\n", c.name, (unsigned long long)c.nb, (unsigned long long)n_occ, expect.size(), ok ? "ok" : "MISMATCH");
        failures += !ok;
    }
    return failures ? 1 : 0;
}
