// Host-side check of the LAYOUT_MINIMIZER bijection (bonsai_b200/csrc/bns_device.cuh): key <-> (home bucket, remainder).
// Built with nvcc as plain host code by tests/test_layout_cpu.py; prints one line per (k, b) and a locality figure.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../bonsai_b200/csrc/bns_device.cuh"
using namespace bns;

static u64 revcomp(u64 x, u32 k) { u64 r = 0; for(u32 i = 0; i < k; ++i) { r = (r << 2) | (3 - (x & 3)); x >>= 2; } return r; }

int main() {
    std::mt19937_64 rng(1);
    int failures = 0;
    for(u32 k : {23u, 25u, 27u, 31u})
        for(u32 b : {12u, 16u, 20u, 24u, 27u, 29u, 31u}) {
            if(b - LOC_GB > LOC_MB) continue;
            const int fmt = (int)loc_fmt_bits(k, b);
            if(fmt > 28 || fmt < 8) continue;
            const u64 mask = (1ull << (2 * k)) - 1;
            const u32 rembits = loc_rembits(k, b);
            size_t bad = 0;
            const size_t n = 100000;
            for(size_t i = 0; i < n; ++i) {
                u64 x = rng() & mask;
                if(i == 0) x = 0; else if(i == 1) x = mask; else if(i == 2) x = 0x5555555555555555ull & mask; else if(i == 3) x = 1;
                else if(i == 4) x = 0x1b1b1b1b1b1b1b1bull & mask;           // ACGT repeats: tied minimizers
                const TableHash t = loc_encode(x, k, b);
                const u64 y = loc_decode(t.home, t.tag, k, b);
                if(y != x || (t.home >> b) != 0 || (t.tag & ((1ull << (64 - rembits)) - 1)) != 0) ++bad;
                // probe sequence and its inverse
                for(u32 d : {0u, 1u, 3u, 4u, 9u, 61u})
                    if(probe_home(LAYOUT_MINIMIZER, probe_bucket(LAYOUT_MINIMIZER, t.home, d, b), d, b) != t.home) ++bad;
            }
            printf("k=%u b=%u fmt=%d bad=%zu\n", k, b, fmt, bad);
            failures += bad != 0;
        }
    // locality: distinct 128-byte lines holding the home buckets of the 120 canonical 31-mers of a random 150 bp read
    const u32 k = 31, b = 28;
    double tot = 0;
    const int R = 500;
    for(int r = 0; r < R; ++r) {
        u64 f = 0;
        const u64 mask = (1ull << 62) - 1;
        u64 lines[120];
        int n = 0;
        for(int i = 0; i < 150; ++i) {
            f = ((f << 2) | (rng() & 3)) & mask;
            if(i >= 30) lines[n++] = loc_encode(std::min(f, revcomp(f, 31)), k, b).home >> 2;      // line = 4 buckets
        }
        std::sort(lines, lines + n);
        tot += std::unique(lines, lines + n) - lines;
    }
    printf("lines_per_read=%.2f\n", tot / R);
    return failures ? 1 : 0;
}
