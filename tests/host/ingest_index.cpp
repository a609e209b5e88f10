// Host-side check of the parallel plain-file ingest (include/bonsai_b200/bonsai.hpp, detail::SimpleFile) against the
// kseq state machine of the same header: same records, and a clean hand-over to kseq where a file leaves the simple form.
// The same for gzip input (detail::GzWindows: decompressor thread, windows cut at record ends, tails carried over), and
// detail::fill_pinned over single and mate files (plain and gzip) against the kseq records, batch by batch.
// Built with g++ -lz by tests/test_cli_cpu.py; argv[1] = directory to write the test files into.
#include <cstdio>
#include <random>
#include <string>
#include <vector>
#include "../../include/bonsai_b200/bonsai.hpp"

// the pinned allocator of the C ABI, stood in for by malloc: this test never touches the device library
extern "C" int bns_b200_host_alloc(void **p, size_t n) { *p = std::malloc(n); return *p ? 0 : -1; }
extern "C" int bns_b200_host_free(void *p) { std::free(p); return 0; }

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
                out.push_back(Rec{std::string(f.text() + r.name_off, r.name_len), std::string(f.text() + r.seq_off, r.seq_len),
                                  r.qual_off == ~0ull ? std::string() : std::string(f.text() + r.qual_off, r.seq_len)});
        if(f.ok) return out;
        *fell_back = true;
        auto rest = by_kseq(path, f.resume_offset());
        out.insert(out.end(), rest.begin(), rest.end());
        return out;
    }
    return by_kseq(path);
}
static void write(const std::string &path, const std::string &txt) { FILE *f = fopen(path.c_str(), "wb"); fwrite(txt.data(), 1, txt.size(), f); fclose(f); }
static void write_gz(const std::string &path, const std::string &txt, size_t members = 1) {
    // `members` > 1: concatenated gzip members (what bgzip / `cat a.gz b.gz` produce); gzread reads through them
    FILE *out = fopen(path.c_str(), "wb"); fclose(out);
    for(size_t m = 0; m < members; ++m) {
        gzFile g = gzopen(path.c_str(), "ab1");
        const size_t lo = txt.size() * m / members, hi = txt.size() * (m + 1) / members;
        for(size_t x = lo; x < hi;) { const int r = gzwrite(g, txt.data() + x, (unsigned)std::min<size_t>(hi - x, 1u << 20)); x += (size_t)r; }
        gzclose(g);
    }
}
// BGZF as bgzip writes it: independent gzip members of <= 64 KiB with a "BC" extra subfield holding the block size, closed by an empty
// block. `plain_member_at`: block index at which an ordinary gzip member (no extra field) is spliced in instead -- still a valid
// gzip file for gzread, but not BGZF from there on.
static void write_bgzf(const std::string &path, const std::string &txt, size_t plain_member_at = ~(size_t)0) {
    FILE *out = fopen(path.c_str(), "wb");
    const size_t BS = 0xff00;
    std::vector<unsigned char> comp(BS + 1024);
    size_t blk = 0;
    for(size_t x = 0;; x += BS, ++blk) {
        const size_t len = x < txt.size() ? std::min(BS, txt.size() - x) : 0;
        if(blk == plain_member_at && len) {
            fclose(out);
            gzFile g = gzopen(path.c_str(), "ab6"); gzwrite(g, txt.data() + x, (unsigned)len); gzclose(g);
            out = fopen(path.c_str(), "ab");
            continue;
        }
        z_stream zs; memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = (Bytef *)(len ? txt.data() + x : ""); zs.avail_in = (uInt)len;
        zs.next_out = comp.data(); zs.avail_out = (uInt)comp.size();
        deflate(&zs, Z_FINISH);
        const size_t clen = comp.size() - zs.avail_out;
        deflateEnd(&zs);
        const unsigned bsize = (unsigned)(clen + 25);                    // total block size - 1
        const unsigned char hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (unsigned char)(bsize & 0xff), (unsigned char)(bsize >> 8)};
        const unsigned long crc = crc32(crc32(0L, Z_NULL, 0), (const Bytef *)txt.data() + (len ? x : 0), (uInt)len);
        const unsigned char tail[8] = {(unsigned char)crc, (unsigned char)(crc >> 8), (unsigned char)(crc >> 16), (unsigned char)(crc >> 24),
                                       (unsigned char)len, (unsigned char)(len >> 8), (unsigned char)(len >> 16), (unsigned char)(len >> 24)};
        fwrite(hdr, 1, 18, out); fwrite(comp.data(), 1, clen, out); fwrite(tail, 1, 8, out);
        if(!len) break;                                                   // that was the end-of-file block
    }
    fclose(out);
}
// process_dataset's reader, batch by batch: fill_pinned while the index serves, read_pinned (kseq) after the hand-over
static std::vector<Rec> by_batches(const std::string &p1, const std::string *p2, unsigned nt, int chunk, size_t *n_batches, size_t *n_handovers) {
    std::vector<Rec> out;
    detail::KSeq ks1(p1.c_str());
    std::unique_ptr<detail::KSeq> ks2(p2 ? new detail::KSeq(p2->c_str()) : nullptr);
    std::unique_ptr<detail::SimpleFile> f1(new detail::SimpleFile(p1.c_str(), nt)), f2(p2 ? new detail::SimpleFile(p2->c_str(), nt) : nullptr);
    bool use_index = f1->ok && (!f2 || f2->ok);
    detail::PinnedBatch ring[3];                                       // batches stay alive for two more rounds, like the ring in flight
    std::vector<std::vector<Rec>> pending(3);
    *n_batches = 0; *n_handovers = 0;
    for(int i = 0;; i = (i + 1) % 3) {
        detail::PinnedBatch &b = ring[i];
        bool got = false;
        if(use_index) {
            got = detail::fill_pinned(chunk, b, *f1, f2.get());
            if(!got && f1->drained() && (!f2 || f2->drained())) break;
            if(!got) {
                ++*n_handovers;
                gzseek(ks1.fp, (z_off_t)f1->resume_offset(), SEEK_SET);
                if(f2) gzseek(ks2->fp, (z_off_t)f2->resume_offset(), SEEK_SET);
                f1->stop(); if(f2) f2->stop();
                use_index = false;
            }
        }
        if(!got && !use_index) got = detail::read_pinned(chunk, b, &ks1, ks2.get());
        if(!got) break;
        ++*n_batches;
        for(size_t r = 0; r < b.n; ++r) {
            const std::string seq(b.bases + b.offs[r], b.offs[r + 1] - b.offs[r]);
            if(b.map) {
                const detail::RecRef &ref = b.refs[r];
                const char *mp = b.map_of(r);
                out.push_back(Rec{std::string(mp + ref.name_off, ref.name_len), seq, ref.qual_off == ~0ull ? std::string() : std::string(mp + ref.qual_off, ref.seq_len)});
            } else out.push_back(Rec{b.names[r], seq, b.quals[r]});
        }
    }
    return out;
}

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    std::mt19937 rng(5);
    auto seq = [&](int n) { std::string s(n, 'A'); for(auto &c : s) c = "ACGTN"[rng() % 100 < 2 ? 4 : rng() % 4]; return s; };
    auto qual = [&](int n, int i) { std::string q(n, 'I'); for(auto &c : q) c = (char)(33 + rng() % 60); if(i % 7 == 0) q[0] = '@'; if(i % 11 == 0) q[0] = '+'; return q; };
    int failures = 0;
    {   // the emitters' decimal formatting against printf
        std::vector<long> vals = {0, 1, 9, 10, 11, 99, 100, 101, 150, 999, 1000, 65535, 65536, 999999, 1000000, 2147483647L, 2147483648L, 4294967295L,
                                  -1, -9, -10, -150, -2147483648L, 9223372036854775807L, -9223372036854775807L - 1};
        for(int i = 0; i < 2000; ++i) vals.push_back((long)(rng() >> (rng() % 64)) * ((rng() & 1) ? 1 : -1));
        bool ok = true;
        for(long v : vals) {
            char b[32];
            std::string a, c;
            detail::put_i(a, v);
            snprintf(b, sizeof b, "%ld", v);
            ok = ok && a == b;
            if(v >= 0 && v <= 4294967295L) { detail::put_u(c, (u32)v); snprintf(b, sizeof b, "%u", (u32)v); ok = ok && c == b; }
        }
        printf("decimal formatting of %zu values %s\n", vals.size(), ok ? "ok" : "MISMATCH");
        failures += !ok;
    }
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
    // .xz / .bz2 through `xz|bzip2 -dc` (encoder.h:511-524), where those tools exist: the same records; never the index
    for(const char *tool : {"xz", "bzip2"}) {
        const std::string ext = tool[0] == 'x' ? ".xz" : ".bz2", plain = dir + "/fastq_simple.txt", comp = dir + "/popen it's.fq" + ext;
        const std::string cmd = std::string(tool) + " -c '" + plain + "' > '" + dir + "/popen it'\\''s.fq" + ext + "' 2>/dev/null";
        if(std::system(cmd.c_str()) != 0) { printf("%s not available: skipped ok\n", tool); continue; }
        const auto ref = by_kseq(plain);
        bool used = true, fell = true;
        const auto got = by_index(comp, 4, &used, &fell);
        const bool ok = got.size() == ref.size() && std::equal(got.begin(), got.end(), ref.begin()) && !used && !fell;
        printf("popen %s records=%zu/%zu %s\n", tool, got.size(), ref.size(), ok ? "ok" : "MISMATCH");
        failures += !ok;
    }
    // gzip input: the same texts, compressed (single- and multi-member), small inflate windows
    setenv("BNS_B200_GZ_WINDOW", "300000", 1);
    for(auto &c : cases) {
        for(size_t members : {(size_t)1, (size_t)5}) {
            const std::string path = dir + "/" + c.name + "_m" + std::to_string(members) + ".gz";
            write_gz(path, c.txt, members);
            const auto ref = by_kseq(path);
            for(unsigned nt : {1u, 4u}) {
                bool used = false, fell = false;
                const auto got = by_index(path, nt, &used, &fell);
                const bool same = got.size() == ref.size() && std::equal(got.begin(), got.end(), ref.begin());
                const bool ok = same && used == c.expect_index && fell == c.expect_fallback;
                printf("gz %s members=%zu nt=%u records=%zu/%zu index=%d fallback=%d %s\n", c.name, members, nt, got.size(), ref.size(), (int)used, (int)fell, ok ? "ok" : "MISMATCH");
                failures += !ok;
            }
        }
    }
    // BGZF: blocks inflate in parallel out of the mapping; a file that stops being BGZF half way is finished by kseq / gzread
    for(auto &c : cases) {
        for(int spliced = 0; spliced < 2; ++spliced) {
            const std::string path = dir + "/" + c.name + (spliced ? "_spliced" : "") + ".bgz";
            write_bgzf(path, c.txt, spliced ? 20 : ~(size_t)0);
            { detail::SimpleFile probe(path.c_str(), 2); if(!probe.gz || !probe.gz->bgzf) { printf("bgzf %s not detected MISMATCH\n", c.name); ++failures; } }
            const auto ref = by_kseq(path);
            for(unsigned nt : {1u, 4u}) {
                bool used = false, fell = false;
                const auto got = by_index(path, nt, &used, &fell);
                const bool same = got.size() == ref.size() && std::equal(got.begin(), got.end(), ref.begin());
                const bool ok = same && used == c.expect_index && fell == (c.expect_fallback || spliced);
                printf("bgzf %s spliced=%d nt=%u records=%zu/%zu index=%d fallback=%d %s\n", c.name, spliced, nt, got.size(), ref.size(), (int)used, (int)fell, ok ? "ok" : "MISMATCH");
                failures += !ok;
            }
        }
    }
    // batches (fill_pinned / read_pinned), single and mate files; mates of different line lengths so that their windows do not line up;
    // the second mate file of the last pairing is one record short (bseq_read stops there)
    {
        std::string m1, m2, m2short, m2multi;
        for(int i = 0; i < 15000; ++i) {
            const int n1 = 30 + rng() % 200, n2 = 80 + rng() % 300;
            m1 += "@p" + std::to_string(i) + "/1\n" + seq(n1) + "\n+\n" + qual(n1, i) + "\n";
            const std::string r2 = "@p" + std::to_string(i) + "/2\n" + seq(n2) + "\n+\n" + qual(n2, i) + "\n";
            m2 += r2;
            if(i + 1 < 15000) m2short += r2;
            m2multi += i == 9000 ? "@p9000/2\nACGT\nACGT\n+\nIIIIIIII\n" : r2;
        }
        struct Pairing { const char *name; std::string a, b; int gz_a, gz_b; size_t handovers; };   // 0 plain, 1 gzip, 2 BGZF   // handovers: kseq takes over (not at a clean end)
        std::vector<Pairing> pairings = {{"single_plain", m1, "", false, false, 0}, {"single_gz", m1, "", true, false, 0},
                                         {"pair_plain", m1, m2, false, false, 0}, {"pair_gz", m1, m2, true, true, 0}, {"pair_mixed", m1, m2, false, true, 0},
                                         {"pair_gz_short", m1, m2short, true, true, 1}, {"pair_short_gz", m2short, m1, true, true, 1},
                                         {"pair_gz_multiline", m1, m2multi, true, true, 1},
                                         {"single_bgzf", m1, "", 2, 0, 0}, {"pair_bgzf", m1, m2, 2, 2, 0}, {"pair_bgzf_gz_short", m1, m2short, 2, 1, 1}};
        for(auto &pr : pairings) {
            const std::string pa = dir + "/" + pr.name + "_1" + (pr.gz_a ? ".gz" : ".fq"), pb = dir + "/" + pr.name + "_2" + (pr.gz_b ? ".gz" : ".fq");
            auto put = [](int kind, const std::string &path, const std::string &txt) { kind == 2 ? write_bgzf(path, txt) : kind == 1 ? write_gz(path, txt) : write(path, txt); };
            put(pr.gz_a, pa, pr.a);
            const bool paired = !pr.b.empty();
            if(paired) put(pr.gz_b, pb, pr.b);
            std::vector<Rec> ref;
            {
                const auto ra = by_kseq(pa);
                if(!paired) ref = ra;
                else { const auto rb = by_kseq(pb); for(size_t i = 0; i < std::min(ra.size(), rb.size()); ++i) { ref.push_back(ra[i]); ref.push_back(rb[i]); } }
            }
            for(int chunk : {1 << 16, 1 << 22}) {
                size_t nb = 0, nh = 0;
                const auto got = by_batches(pa, paired ? &pb : nullptr, 4, chunk, &nb, &nh);
                const bool ok = got.size() == ref.size() && std::equal(got.begin(), got.end(), ref.begin()) && nh == pr.handovers;
                printf("batches %s chunk=%d records=%zu/%zu batches=%zu handovers=%zu %s\n", pr.name, chunk, got.size(), ref.size(), nb, nh, ok ? "ok" : "MISMATCH");
                failures += !ok;
            }
        }
    }
    return failures ? 1 : 0;
}
