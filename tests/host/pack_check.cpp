// pack_check.cpp -- the host packer (bonsai_b200/csrc/bns_pack.cpp) against a byte-at-a-time model of its format, for the
// instruction set BNS_B200_PACK_ISA selects; `pack_check bench [threads]` times it instead.
//   g++ -O2 -std=c++17 -pthread -I bonsai_b200/csrc tests/host/pack_check.cpp bonsai_b200/csrc/bns_pack.cpp -o pack_check
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include "bns_pack.h"

static void model(const std::string &s, std::vector<uint16_t> &units, std::vector<uint32_t> &susp, std::vector<uint64_t> &exc, uint64_t unit0) {
    const size_t nu = (s.size() + 7) / 8;
    units.assign(nu, 0); susp.assign((s.size() + 255) / 256, 0); exc.clear();
    for(size_t u = 0; u < nu; ++u) {
        unsigned v = 0, bad = 0;
        for(unsigned i = 0; i < 8; ++i) {
            const size_t p = 8 * u + i;
            const unsigned char ch = p < s.size() ? (unsigned char)s[p] : 'A';
            unsigned code = 0; bool ok = true;
            switch(ch) { case 'A': case 'a': code = 0; break; case 'C': case 'c': code = 1; break; case 'G': case 'g': code = 2; break;
                         case 'T': case 't': code = 3; break; default: ok = false; code = ((ch >> 1) & 3u) ^ ((ch >> 2) & 1u); }
            v = (v << 2) | code; bad = (bad << 1) | (ok ? 0u : 1u);
        }
        units[u] = (uint16_t)v;
        if(bad) { susp[u >> 5] |= 1u << (u & 31); exc.push_back(((unit0 + u) << 8) | bad); }
    }
}

int main(int argc, char **argv) {
    if(argc > 1 && !strcmp(argv[1], "bench")) {
        const unsigned nt = argc > 2 ? atoi(argv[2]) : 8;
        const size_t n = (size_t)1500 << 20;
        std::string s(n, 'A');
        std::mt19937_64 rng(1);
        for(size_t i = 0; i < n; i += 8) { uint64_t r = rng(); for(int j = 0; j < 8 && i + j < n; ++j) s[i + j] = "ACGT"[(r >> (2 * j)) & 3]; }
        std::vector<uint16_t> units(n / 8 + 64); std::vector<uint32_t> susp(n / 256 + 8);
        bns::PackPool pool(nt);
        const unsigned tasks = nt * 4;
        const size_t per = ((n / tasks) + 255) / 256 * 256;
        std::vector<std::vector<uint64_t>> exc(tasks);
        for(int rep = 0; rep < 4; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            pool.start(tasks, [&](unsigned t) {
                const size_t b = std::min(n, per * t), e = std::min(n, per * (t + 1));
                bns::pack_range(s.data() + b, e - b, units.data() + b / 8, susp.data() + b / 256, exc[t], b / 8);
            });
            pool.wait();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("%s, %u threads: %.2f GB/s (%.0f Mreads/s of 150 bp)\n", bns::pack_isa(), nt, n / dt / 1e9, n / 150.0 / dt / 1e6);
        }
        return 0;
    }
    std::mt19937_64 rng(7);
    int fails = 0, cases = 0;
    for(int it = 0; it < 3000; ++it) {
        const size_t n = it < 600 ? it : rng() % 5000;
        std::string s(n, 'A');
        const int mode = it % 4;
        for(size_t i = 0; i < n; ++i) {
            const uint64_t r = rng();
            if(mode == 0) s[i] = "ACGT"[r & 3];
            else if(mode == 1) s[i] = (r % 97 == 0) ? 'N' : "ACGTacgt"[r & 7];
            else if(mode == 2) s[i] = (char)(r & 0xff);
            else s[i] = (r % 11 == 0) ? "NnRYKM.-*\n"[(r >> 8) % 10] : "ACGT"[r & 3];
        }
        const uint64_t unit0 = rng() % 1000000;
        std::vector<uint16_t> mu, u((n + 7) / 8 + 1, 0xdead); std::vector<uint32_t> ms, sp((n + 255) / 256 + 1, 0xdeadbeef); std::vector<uint64_t> me, e;
        model(s, mu, ms, me, unit0);
        bns::pack_range(s.data(), n, u.data(), sp.data(), e, unit0);
        ++cases;
        bool ok = e == me && u.back() == 0xdead && sp.back() == 0xdeadbeef;
        for(size_t i = 0; ok && i < mu.size(); ++i) ok = mu[i] == u[i];
        for(size_t i = 0; ok && i < ms.size(); ++i) ok = ms[i] == sp[i];
        if(!ok) { ++fails; if(fails < 5) printf("FAIL case %d n=%zu mode=%d\n", it, n, mode); }
    }
    // pieces cut at multiples of 256 bases and run on the pool give the same bytes as one call
    {
        const size_t n = 1000003;
        std::string s(n, 'A');
        for(size_t i = 0; i < n; ++i) { const uint64_t r = rng(); s[i] = (r % 1009 == 0) ? 'N' : "ACGT"[r & 3]; }
        std::vector<uint16_t> mu; std::vector<uint32_t> ms; std::vector<uint64_t> me;
        model(s, mu, ms, me, 0);
        std::vector<uint16_t> u((n + 7) / 8); std::vector<uint32_t> sp((n + 255) / 256);
        bns::PackPool pool(5);
        const unsigned tasks = 13;
        const size_t per = ((n / tasks) + 255) / 256 * 256;
        std::vector<std::vector<uint64_t>> exc(tasks);
        for(int rep = 0; rep < 3; ++rep) {
            for(auto &v : exc) v.clear();
            pool.start(tasks, [&](unsigned t) {
                const size_t b = std::min(n, per * t), e = std::min(n, per * (t + 1));
                if(e > b) bns::pack_range(s.data() + b, e - b, u.data() + b / 8, sp.data() + b / 256, exc[t], b / 8);
            });
            pool.wait();
            std::vector<uint64_t> all;
            for(auto &v : exc) all.insert(all.end(), v.begin(), v.end());
            ++cases;
            if(all != me || u != mu || sp != ms) { ++fails; printf("FAIL pooled\n"); }
        }
    }
    printf("%s: %d cases, %d failures\n", bns::pack_isa(), cases, fails);
    return fails != 0;
}
