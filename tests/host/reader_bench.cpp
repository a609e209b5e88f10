// Measurement harness, host only (links against tests/host/abi_stub.cpp): how fast the parallel FASTQ / FASTA ingest of
// include/bonsai_b200/bonsai.hpp (detail::SimpleFile + detail::fill_pinned) turns a file in the page cache into pinned batches,
// by thread count, next to two floors measured in the same run: a parallel memchr pass over the mapping and a parallel memcpy.
//   reader_bench <file> <chunk bases> <threads> [<threads> ...]
#include <chrono>
#include "../../include/bonsai_b200/bonsai.hpp"
using namespace bns;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    if(argc < 4) { std::fprintf(stderr, "usage: %s file chunk threads...\n", argv[0]); return 1; }
    const int chunk = std::atoi(argv[2]);
    for(int a = 3; a < argc; ++a) {
        const unsigned nt = (unsigned)std::atoi(argv[a]);
        {
            detail::MappedFile m(argv[1]);
            double t0 = now();
            std::vector<std::thread> pool; std::vector<size_t> cnt(nt);
            for(unsigned t = 0; t < nt; ++t)
                pool.emplace_back([&, t] { size_t lo = m.n * t / nt, hi = m.n * (t + 1) / nt, c = 0; const char *p = m.p + lo, *e = m.p + hi;
                                           while((p = (const char *)std::memchr(p, '\n', (size_t)(e - p)))) { ++c; ++p; } cnt[t] = c; });
            for(auto &th : pool) th.join();
            const double t1 = now();
            std::vector<char> dst(m.n);
            pool.clear();
            const double t2 = now();
            for(unsigned t = 0; t < nt; ++t) pool.emplace_back([&, t] { size_t lo = m.n * t / nt, hi = m.n * (t + 1) / nt; std::memcpy(dst.data() + lo, m.p + lo, hi - lo); });
            for(auto &th : pool) th.join();
            const double t3 = now();
            std::printf("threads %2u: memchr pass %.1f GB/s, memcpy %.1f GB/s", nt, m.n / (t1 - t0) / 1e9, m.n / (t3 - t2) / 1e9);
        }
        double best = 1e9; size_t n = 0, nb = 0;
        for(int rep = 0; rep < 3; ++rep) {
            const double t0 = now();
            detail::SimpleFile f(argv[1], nt);
            detail::PinnedBatch b; b.reserve((size_t)chunk); b.keep_qual = false;
            n = nb = 0;
            while(detail::fill_pinned(chunk, b, f)) { n += b.n; ++nb; }
            best = std::min(best, now() - t0);
        }
        std::printf("; ingest %zu reads in %zu batches: %.3f s = %.1f Mreads/s\n", n, nb, best, n / best / 1e6);
    }
    return 0;
}
