// The C++ mirror of the reference's Encoder surface (include/bonsai_b200/bonsai.hpp) against expected streams handed in by the
// test (tests/test_gpu_parity.py::test_cpp_encoder_surface computes them with the CPU oracle): for_each(fn, str, l), the record
// overloads, and the call-by-call surface assign / has_next_kmer / next_kmer / next_minimizer / next_canonicalized_minimizer.
// argv[1]: case file, one case per line:  kind k w score canon gaps(comma separated, k-1 of them, or -) seq(or -) n v0 v1 ... (hex)
//   kind 0 for_each(fn, str, l)   1 for_each_record   2 next_kmer   3 next_minimizer   4 next_canonicalized_minimizer
#include <cstdio>
#include <fstream>
#include <sstream>
#include "../../include/bonsai_b200/bonsai.hpp"

using namespace bns;

template <typename Score>
static std::vector<u64> run_case(int kind, const Spacer &sp, bool canon, const std::string &seq) {
    Encoder<Score> enc(sp, canon);
    std::vector<u64> got;
    auto fn = [&](u64 x) { got.push_back(x); };
    if(kind == 0) enc.for_each(fn, seq.data(), seq.size());
    else if(kind == 1) enc.for_each_record(fn, seq.data(), seq.size());
    else {
        enc.assign(seq.data(), seq.size());
        while(enc.has_next_kmer()) got.push_back(kind == 2 ? enc.next_kmer() : kind == 3 ? enc.next_minimizer() : enc.next_canonicalized_minimizer());
        // a second string through the same object: assign() starts over
        if(!seq.empty()) {
            std::vector<u64> again;
            enc.assign(seq.data(), seq.size());
            while(enc.has_next_kmer()) again.push_back(kind == 2 ? enc.next_kmer() : kind == 3 ? enc.next_minimizer() : enc.next_canonicalized_minimizer());
            if(again != got) got.push_back(0xbadbadbadull);
        }
    }
    return got;
}

int main(int argc, char **argv) {
    if(argc < 2) return 2;
    std::ifstream in(argv[1]);
    std::string line;
    int failures = 0, n_cases = 0;
    while(std::getline(in, line)) {
        if(line.empty()) continue;
        std::istringstream ss(line);
        int kind, k, w, score, canon; std::string gaps_s, seq; size_t n;
        ss >> kind >> k >> w >> score >> canon >> gaps_s >> seq >> n;
        if(seq == "-") seq.clear();
        std::vector<u64> expect(n);
        for(auto &v : expect) { std::string h; ss >> h; v = std::strtoull(h.c_str(), nullptr, 16); }
        spvec_t gaps;
        if(gaps_s != "-") { std::istringstream gs(gaps_s); std::string t; while(std::getline(gs, t, ',')) gaps.push_back((u16)std::atoi(t.c_str())); }
        try {
            const Spacer sp((unsigned)k, (u32)w, gaps);
            const std::vector<u64> got = score ? run_case<score::Entropy>(kind, sp, canon != 0, seq) : run_case<score::Lex>(kind, sp, canon != 0, seq);
            const bool ok = got == expect;
            if(!ok) std::printf("case %d kind=%d k=%d w=%d score=%d canon=%d len=%zu: %zu values, expected %zu MISMATCH\n", n_cases, kind, k, w, score, canon,
                                seq.size(), got.size(), expect.size());
            failures += !ok;
        } catch(const std::exception &e) {
            std::printf("case %d: exception %s MISMATCH\n", n_cases, e.what());
            ++failures;
        }
        ++n_cases;
    }
    std::printf("%d cases, %d failures\n", n_cases, failures);
    return failures ? 1 : 0;
}
