// The file-level overloads of the C++ Encoder mirror (include/bonsai_b200/bonsai.hpp: for_each over a path / std::string / gzFile /
// list of paths, for_each_canon / for_each_uncanon) batch the records of a file into few device calls. Checked against the DUMMY ABI
// (tests/host/abi_stub.cpp: an arbitrary word per position that depends on the canonicalize flag): whatever the batch size, fn
// must see exactly what record-by-record calls produce, in file order, from plain, gzip and xz files alike.
#include <cstdio>
#include <random>
#include "../../include/bonsai_b200/bonsai.hpp"
using namespace bns;

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    std::mt19937 rng(3);
    std::string fa;
    for(int i = 0; i < 3000; ++i) {
        const int n = i % 97 == 0 ? 5 : 20 + (int)(rng() % 400);                  // some records shorter than k
        std::string s(n, 'A');
        for(auto &c : s) c = "ACGT"[rng() % 4];
        fa += ">r" + std::to_string(i) + "\n" + s + "\n";
    }
    const std::string plain = dir + "/enc.fa", gz = dir + "/enc.fa.gz", xz = dir + "/enc.fa.xz";
    { FILE *f = fopen(plain.c_str(), "wb"); fwrite(fa.data(), 1, fa.size(), f); fclose(f); }
    { gzFile g = gzopen(gz.c_str(), "wb"); gzwrite(g, fa.data(), (unsigned)fa.size()); gzclose(g); }
    const bool have_xz = std::system(("xz -k -f -1 '" + plain + "' 2>/dev/null").c_str()) == 0;
    int failures = 0;
    for(int canon = 0; canon < 2; ++canon) {
        Encoder<score::Lex> enc(Spacer(31, 31), canon != 0);
        // record by record
        std::vector<u64> ref[2];
        for(int want = 0; want < 2; ++want) {
            Encoder<score::Lex> e1(Spacer(31, 31), want != 0);
            detail::KSeq ks(plain.c_str());
            while(ks.read() >= 0) e1.for_each_record([&](u64 x) { ref[want].push_back(x); }, ks.seq.data(), ks.seq.size());
        }
        for(const char *batch : {"1", "5000", "100000", ""}) {
            if(*batch) setenv("BNS_B200_ENCODE_BATCH", batch, 1); else unsetenv("BNS_B200_ENCODE_BATCH");
            auto check = [&](const char *what, const std::vector<u64> &got, int want) {
                const bool ok = got == ref[want];
                printf("canon=%d batch=%s %s: %zu/%zu %s\n", canon, *batch ? batch : "default", what, got.size(), ref[want].size(), ok ? "ok" : "MISMATCH");
                failures += !ok;
            };
            std::vector<u64> got;
            auto fn = [&](u64 x) { got.push_back(x); };
            enc.for_each(fn, plain.c_str()); check("path", got, canon); got.clear();
            enc.for_each(fn, gz); check("string(gz)", got, canon); got.clear();
            if(have_xz) { enc.for_each(fn, xz.c_str()); check("path(xz)", got, canon); got.clear(); }
            { gzFile fp = gzopen(gz.c_str(), "rb"); enc.for_each(fn, fp); gzclose(fp); check("gzFile", got, canon); got.clear(); }
            enc.for_each_canon(fn, plain.c_str()); check("for_each_canon", got, 1); got.clear();
            enc.for_each_uncanon(fn, gz.c_str()); check("for_each_uncanon", got, 0); got.clear();
            enc.for_each(fn, std::vector<std::string>{plain, gz});
            { std::vector<u64> twice(ref[canon]); twice.insert(twice.end(), ref[canon].begin(), ref[canon].end()); const bool ok = got == twice;
              printf("canon=%d batch=%s list: %s\n", canon, *batch ? batch : "default", ok ? "ok" : "MISMATCH"); failures += !ok; }
        }
    }
    return failures ? 1 : 0;
}
