// Host-side check of the parallel plain-file ingest (include/bonsai_b200/bonsai.hpp, detail::SimpleFile) against the
// kseq state machine of the same header: same records, and a clean hand-over to kseq where a file leaves the simple form.
// Built with g++ -lz by tests/test_cli_cpu.py; argv[1] = directory to write the test files into.
#include <cstdio>
#include <random>
#include <string>
#include <vector>
#include "../../include/bonsai_b200/bonsai.hpp"

using namespace bns;
struct Rec { std::string name, seq, qual; bool operator==(const Rec &o) const { return name == o.name && seq == o.seq && qual == o.qual; } };

static std::vector<Rec> by_kseq(const std::string &path, size_t from = 0) {
    detail::KSeq ks(path.c_str());
    if(from) gzseek(ks.fp, (z_off_t)from, SEEK_SET);
    std::vector<Rec> out;
    while(ks.read() >= 0) { trim_readno(ks.name); out.push_back(Rec{ks.name, ks.seq, ks.qual}); }
    return out;
}
// what process_dataset's reader does: the index while the file is simple, kseq from the first window that is not
static std::vector<Rec> by_index(const std::string &path, unsigned nt, bool *used_index, bool *fell_back) {
    std::vector<Rec> out;
    detail::SimpleFile f(path.c_str(), nt);
    *used_index = f.ok; *fell_back = false;
    if(f.ok) {
        while(f.refill())
            for(const auto &r : f.recs)
                out.push_back(Rec{std::string(f.map.p + r.name_off, r.name_len), std::string(f.map.p + r.seq_off, r.seq_len),
                                  r.qual_off == ~0ull ? std::string() : std::string(f.map.p + r.qual_off, r.seq_len)});
        if(f.ok || f.cursor >= f.map.n) return out;
        *fell_back = true;
        auto rest = by_kseq(path, f.cursor);
        out.insert(out.end(), rest.begin(), rest.end());
        return out;
    }
    return by_kseq(path);
}
static void write(const std::string &path, const std::string &txt) { FILE *f = fopen(path.c_str(), "wb"); fwrite(txt.data(), 1, txt.size(), f); fclose(f); }

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    std::mt19937 rng(5);
    auto seq = [&](int n) { std::string s(n, 'A'); for(auto &c : s) c = "ACGTN"[rng() % 100 < 2 ? 4 : rng() % 4]; return s; };
    auto qual = [&](int n, int i) { std::string q(n, 'I'); for(auto &c : q) c = (char)(33 + rng() % 60); if(i % 7 == 0) q[0] = '@'; if(i % 11 == 0) q[0] = '+'; return q; };
    int failures = 0;
    struct Case { const char *name; std::string txt; bool expect_index, expect_fallback; };
    std::vector<Case> cases;
    {   // plain 4-line FASTQ, names with comments / tabs / read numbers, quality lines starting with @ and +, no final newline
        std::string t;
        for(int i = 0; i < 20000; ++i) {
            const int n = 30 + rng() % 200;
            t += "@read" + std::to_string(i) + (i % 3 == 0 ? "/1" : i % 3 == 1 ? "/2 some comment" : "\tx") + "\n" + seq(n) + "\n+" + (i % 5 ? "" : "read") + "\n" + qual(n, i) + "\n";
        }
        t.pop_back();
        cases.push_back({"fastq_simple", t, true, false});
    }
    {   // 2-line FASTA
        std::string t;
        for(int i = 0; i < 20000; ++i) t += ">r" + std::to_string(i) + " c\n" + seq(20 + rng() % 300) + "\n";
        cases.push_back({"fasta_simple", t, true, false});
    }
    {   // FASTA that becomes multi-line half way: the index serves the first windows, kseq the rest
        std::string t;
        for(int i = 0; i < 20000; ++i) {
            const std::string s = seq(100 + rng() % 100);
            t += ">r" + std::to_string(i) + "\n" + (i > 12000 ? s.substr(0, 60) + "\n" + s.substr(60) : s) + "\n";
        }
        cases.push_back({"fasta_then_multiline", t, true, true});
    }
    {   // CRLF FASTQ: never simple
        std::string t;
        for(int i = 0; i < 3000; ++i) { const int n = 50 + rng() % 50; t += "@q" + std::to_string(i) + "\r\n" + seq(n) + "\r\n+\r\n" + qual(n, i) + "\r\n"; }
        cases.push_back({"fastq_crlf", t, true, true});
    }
    {   // empty lines between records
        std::string t;
        for(int i = 0; i < 3000; ++i) { const int n = 50 + rng() % 50; t += "@q" + std::to_string(i) + "\n" + seq(n) + "\n+\n" + qual(n, i) + "\n" + (i % 100 == 99 ? "\n" : ""); }
        cases.push_back({"fastq_empty_lines", t, true, true});
    }
    setenv("BNS_B200_FASTQ_WINDOW", "300000", 1);                      // many windows, several threads per window
    for(auto &c : cases) {
        const std::string path = dir + "/" + c.name + ".txt";
        write(path, c.txt);
        const auto ref = by_kseq(path);
        for(unsigned nt : {1u, 4u, 7u}) {
            bool used = false, fell = false;
            const auto got = by_index(path, nt, &used, &fell);
            const bool same = got.size() == ref.size() && std::equal(got.begin(), got.end(), ref.begin());
            const bool ok = same && used == c.expect_index && fell == c.expect_fallback;
            printf("%s nt=%u records=%zu/%zu index=%d fallback=%d %s\n", c.name, nt, got.size(), ref.size(), (int)used, (int)fell, ok ? "ok" : "MISMATCH");
            failures += !ok;
        }
    }
    return failures ? 1 : 0;
}
