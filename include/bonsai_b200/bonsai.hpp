// bonsai.hpp -- the reference's C++ surface for the classify path, re-created over the B200 C ABI.
//
// Same names, argument meaning and error behaviour as dnbaker/bonsai @ 6741de9c (paths below are in that tree):
//   bns::Spacer, parse_spacing            include/bonsai/spacer.h:29-71
//   bns::Encoder<Score>::for_each         include/bonsai/encoder.h:416 (string overload) / :448-530 (record overloads)
//   bns::Database                         include/bonsai/database.h:17-102 (file layout as the reference INTENDS it)
//   bns::build_parent_map                 include/bonsai/util.h:766-785
//   bns::ClassifierGeneric, classify_seqs, process_dataset, bseq1_t, the Kraken/FASTQ emitters
//                                         include/bonsai/classifier.h:23-337, include/bonsai/kseq_declare.h:40-175
// All k-mer / lookup / resolve work happens in libbonsai_b200.so on the GPU; this header is host plumbing and
// text formatting only. There is no CPU fallback: constructing an Encoder or ClassifierGeneric without a CUDA
// device throws std::runtime_error (the reference's RUNTIME_ERROR convention, util.h:540-551).
#pragma once
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cinttypes>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <map>
#include <type_traits>
#include <string>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

#include "../bonsai_b200.h"

namespace bns {

using u8 = std::uint8_t;
using u16 = std::uint16_t;
using u32 = std::uint32_t;
using u64 = std::uint64_t;
using tax_t = u32;
using spvec_t = std::vector<u16>;

#define BNS_RUNTIME_ERROR(msg) throw std::runtime_error(std::string("[") + __FILE__ + ":" + std::to_string(__LINE__) + "] " + (msg))

// ---- Spacer ----------------------------------------------------------------------------------------------------
inline u32 comb_size(const spvec_t &spaces) {                       // spacer.h:14-19
    u32 ret(spaces.size() + 1);
    for(const auto i : spaces) ret += i;
    return ret;
}
// "1x3,0x5": comma-separated `gap` or `gap x repeat` items (spacer.h:29-47). An empty string means unspaced.
inline spvec_t parse_spacing(const char *text, unsigned k) {
    if(text == nullptr || text[0] == '\0') return spvec_t(k - 1, 0);
    spvec_t gaps;
    const char *p = text;
    while(*p) {
        char *endp = nullptr;
        const int gap = (int)std::strtoul(p, &endp, 10);
        int times = 1;
        if(*endp == 'x') {
            times = (int)std::strtoul(endp + 1, &endp, 10);
            if(times < 1) times = 1;                       // "gx0" still contributes one entry, as in the reference
        }
        gaps.insert(gaps.end(), (size_t)times, (u16)gap);
        const char *comma = std::strchr(endp, ',');
        if(comma == nullptr) break;
        p = comma + 1;
    }
    return gaps;
}
struct Spacer {
    spvec_t s_;        // offsets (gap + 1), spacer.h:65
    u32 k_, c_, w_;
    Spacer(unsigned k, u32 w, spvec_t spaces = spvec_t{})
        : s_(spaces.size() ? spaces : spvec_t(k - 1, 0)), k_(k), c_(comb_size(s_)), w_(std::max((int)c_, (int)w)) {
        for(auto &i : s_) ++i;
        if(s_.size() + 1 != k) BNS_RUNTIME_ERROR("Error: input vector must have size 1 less than k.");
    }
    Spacer(unsigned k, u32 w, const char *space_string) : Spacer(k, w, parse_spacing(space_string, k)) {}
    explicit Spacer(unsigned k) : Spacer(k, k) {}
    u32 k() const { return k_; }
    u32 w() const { return w_; }
    u32 c() const { return c_; }
    bool unspaced() const { return std::all_of(s_.begin(), s_.end(), [](u16 x) { return x == 1; }); }
    bool unwindowed() const { return k_ == w_; }
    spvec_t sub1() const { spvec_t r(s_); for(auto &x : r) --x; return r; }            // spacer.h:173
    // the comb spelled out: bases of `kmer` at their offsets, '-' in the gaps (Spacer::to_string, spacer.h:127-139)
    std::string to_string(u64 kmer) const {
        std::string out(c_, '-');
        size_t at = 0;
        for(u32 j = 0; j < k_; ++j) {
            out[at] = "ACGT"[(kmer >> (2 * (k_ - 1 - j))) & 3u];
            if(j + 1 < k_) at += s_[j];
        }
        return out;
    }
};

namespace score {
struct Lex { static constexpr u32 id = BNS_SCORE_LEX; };
struct Entropy { static constexpr u32 id = BNS_SCORE_ENTROPY; };
}  // namespace score

namespace detail {
struct Handle {                        // RAII over bns_b200_t
    bns_b200_t *h = nullptr;
    Handle() = default;
    Handle(const Handle &) = delete;
    Handle &operator=(const Handle &) = delete;
    ~Handle() { if(h) bns_b200_close(h); }
};
inline std::shared_ptr<Handle> open_handle(const Spacer &sp, u32 score, bool canon, u32 api, int device = -1) {
    bns_b200_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.k = sp.k_; cfg.w = sp.w_;
    const spvec_t gaps = sp.sub1();
    std::copy(gaps.begin(), gaps.end(), cfg.gaps);
    cfg.score = score; cfg.canonicalize = canon; cfg.api = api; cfg.device = device;
    // the reference's entropy score casts a negative double to u64 (UB): do what a `-march=native` build of the reference
    // does on THIS host -- vcvttsd2usi saturates where AVX-512 exists, the cvttsd2si sequence wraps elsewhere
    cfg.entropy_cast = __builtin_cpu_supports("avx512f") ? BNS_CAST_SATURATE : BNS_CAST_WRAP;
    // the callers of this header run their own threads next to the device calls (process_dataset's readers and formatters, the
    // reference's kt_for workers): the library does not add packing threads of its own unless BNS_B200_HOST_PACK asks for them
    cfg.host_pack_threads = 0xffffffffu;
    auto ret = std::make_shared<Handle>();
    const int rc = bns_b200_open(&cfg, &ret->h);
    if(rc) BNS_RUNTIME_ERROR(std::string("bns_b200_open: ") + bns_b200_last_error(nullptr));
    return ret;
}
inline void check(bns_b200_t *h, int rc, const char *what) {
    if(rc) BNS_RUNTIME_ERROR(std::string(what) + ": " + bns_b200_last_error(h));
}

// minimal kseq: FASTA/FASTQ records over gzread (klib/kseq.h:178 semantics: name up to the first space, comment the
// rest of the header line, sequence lines concatenated, optional '+' quality)
struct KSeq {
    gzFile fp = nullptr;
    std::vector<unsigned char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    int last_char = 0;
    std::string name, comment, seq, qual;
    std::FILE *pfp = nullptr;                              // .xz / .bz2 / .zst: read through `xz|bzip2|zstd -dc`, encoder.h:511-524
    bool owns_fp = true;
    explicit KSeq(const char *path) : buf(1 << 18) {
        const std::string p(path);
        auto ends = [&](const char *suf) { const size_t n = std::strlen(suf); return p.size() >= n && p.compare(p.size() - n, n, suf) == 0; };
        const char *tool = ends(".xz") ? "xz" : ends(".bz2") ? "bzip2" : ends(".zst") ? "zstd" : nullptr;
        if(tool) {
            std::string quoted = "'";
            for(char ch : p) { if(ch == '\'') quoted += "'\\''"; else quoted.push_back(ch); }
            quoted.push_back('\'');
            const std::string cmd = std::string(tool) + " -dc -- " + quoted;
            pfp = ::popen(cmd.c_str(), "r");
            if(!pfp) BNS_RUNTIME_ERROR(std::string("Failed to open popen call: ") + cmd);
            fp = gzdopen(::dup(::fileno(pfp)), "rb");
        } else fp = gzopen(path, "rb");
        if(!fp) { if(pfp) ::pclose(pfp); BNS_RUNTIME_ERROR(std::string("Could not open file at ") + path); }
        gzbuffer(fp, 1 << 18);
    }
    explicit KSeq(gzFile borrowed) : fp(borrowed), buf(1 << 18), owns_fp(false) {}     // for_each(fn, gzFile): the caller closes it
    KSeq(const KSeq &) = delete;
    ~KSeq() { if(fp && owns_fp) gzclose(fp); if(pfp) ::pclose(pfp); }
    int getc_() {
        if(pos >= end) {
            if(eof) return -1;
            const int n = gzread(fp, buf.data(), (unsigned)buf.size());
            if(n <= 0) { eof = true; return -1; }
            pos = 0; end = (size_t)n;
        }
        return buf[pos++];
    }
    // ks_getuntil: append bytes up to (not including) a delimiter (SEP_SPACE: isspace, SEP_LINE: newline, with a
    // trailing CR stripped). Returns the number of bytes appended, or -1 at end of file with nothing read.
    enum { SEP_SPACE = 0, SEP_LINE = 1 };
    long getuntil(int delim, std::string &out, bool append, int *dret) {
        if(!append) out.clear();
        const size_t before = out.size();
        bool got_any = false;
        int c = -1;
        for(;;) {
            if(pos >= end) {
                if(eof) break;
                const int n = gzread(fp, buf.data(), (unsigned)buf.size());
                if(n <= 0) { eof = true; break; }
                pos = 0; end = (size_t)n;
            }
            got_any = true;
            const unsigned char *p = buf.data() + pos, *stop = buf.data() + end, *q;
            if(delim == SEP_LINE) q = static_cast<const unsigned char *>(std::memchr(p, '\n', (size_t)(stop - p)));
            else { q = p; while(q < stop && !std::isspace(*q)) ++q; if(q == stop) q = nullptr; }
            if(q) {                                            // delimiter inside the buffer: take the run in one append
                out.append(reinterpret_cast<const char *>(p), (size_t)(q - p));
                c = *q;
                pos = (size_t)(q - buf.data()) + 1;
                break;
            }
            out.append(reinterpret_cast<const char *>(p), (size_t)(stop - p));
            pos = end;
            c = -1;
        }
        if(!got_any) return -1;
        if(delim == SEP_LINE && out.size() > before && out.back() == '\r') out.pop_back();
        if(dret) *dret = c;
        return (long)(out.size() - before);
    }
    // kseq_read (klib/kseq.h:178-217): >= 0 sequence length; -1 end of file; -2 truncated quality string
    long read() {
        int c;
        if(last_char == 0) {                                   // jump to the next header line
            while((c = getc_()) >= 0 && c != '>' && c != '@') {}
            if(c < 0) return -1;
            last_char = c;
        }
        comment.clear(); seq.clear(); qual.clear();
        if(getuntil(SEP_SPACE, name, false, &c) < 0) return -1;
        if(c != '\n') getuntil(SEP_LINE, comment, false, nullptr);
        while((c = getc_()) >= 0 && c != '>' && c != '+' && c != '@') {
            if(c == '\n') continue;                            // skip empty lines
            seq.push_back((char)c);
            getuntil(SEP_LINE, seq, true, nullptr);            // the rest of the line
        }
        if(c == '>' || c == '@') last_char = c;                // the first header char has been read
        if(c != '+') return (long)seq.size();                  // FASTA
        while((c = getc_()) >= 0 && c != '\n') {}              // skip the rest of the '+' line
        if(c < 0) return -2;
        while(qual.size() < seq.size() && getuntil(SEP_LINE, qual, true, nullptr) >= 0) {}
        last_char = 0;
        if(seq.size() != qual.size()) return -2;
        return (long)seq.size();
    }
};
}  // namespace detail

// ---- Encoder ---------------------------------------------------------------------------------------------------
// Encoder<Score>: borrows the string, calls fn(kmer) synchronously and in stream order (encoder.h:416). Not
// thread-safe, one copy per worker (classifier.h:258); a copy opens device contexts of its own.
template <typename ScoreType = score::Lex>
class Encoder {
    std::shared_ptr<detail::Handle> str_, path_;       // lazily opened contexts for the two overload families
    bool canonicalize_;
    std::vector<u64> kmers_;
    detail::Handle *get(bool path_api) {
        auto &p = path_api ? path_ : str_;
        if(!p) p = detail::open_handle(sp_, ScoreType::id, canonicalize_, path_api ? BNS_API_PATH : BNS_API_STRING);
        return p.get();
    }
    std::shared_ptr<detail::Handle> forced_[2];        // record-overload contexts with canonicalisation forced off / on
    // every record of a file through the record overloads on context `h`: records are gathered into batches of ~32 MB of bases
    // (one device call each instead of one per record) and fn sees the k-mers record by record, in file order
    template <typename F> void for_each_records_of(const F &fn, detail::KSeq &ks, bns_b200_t *h) {
        const char *e = std::getenv("BNS_B200_ENCODE_BATCH");                          // bases per batch (tests use a small one)
        const size_t BATCH_BASES = e && std::atoll(e) > 0 ? (size_t)std::atoll(e) : ((size_t)32 << 20);
        std::string bases;
        std::vector<u64> offs(1, 0), ooffs(1, 0);
        std::vector<u32> counts;
        auto flush = [&]() {
            const u64 n = offs.size() - 1;
            if(!n) return;
            kmers_.resize(ooffs.back() + 1);
            counts.assign(n, 0);
            detail::check(h, bns_b200_encode_batch(h, bases.data(), offs.data(), n, kmers_.data(), ooffs.data(), counts.data()), "bns_b200_encode_batch");
            for(u64 i = 0; i < n; ++i) for(u32 j = 0; j < counts[i]; ++j) fn(kmers_[ooffs[i] + j]);
            bases.clear(); offs.assign(1, 0); ooffs.assign(1, 0);
        };
        while(ks.read() >= 0) {
            bases += ks.seq;
            offs.push_back(bases.size());
            ooffs.push_back(ooffs.back() + bns_b200_encode_bound(h, ks.seq.size()));
            if(bases.size() >= BATCH_BASES) flush();
        }
        flush();
    }
    template <typename F> void for_each_as(const F &fn, const char *path, bool canon) {
        detail::Handle *hd;
        if(canon == canonicalize_) hd = get(true);
        else {
            auto &p = forced_[canon];
            if(!p) p = detail::open_handle(sp_, ScoreType::id, canon, BNS_API_PATH);
            hd = p.get();
        }
        detail::KSeq ks(path);
        for_each_records_of(fn, ks, hd->h);
    }
    // call-by-call state (assign / next_*)
    const char *s_ = nullptr;
    u64 l_ = 0, pos_ = 0;
    int iter_kind_ = -1;                               // 0 next_kmer, 1 next_minimizer, 2 next_canonicalized_minimizer
    std::vector<u64> iter_;                            // iter_[p]: what call number p returns
    std::shared_ptr<detail::Handle> iter_h_[3];
    u64 iter_value(int kind) {
        if(!has_next_kmer()) return ENCODE_OVERFLOW;   // the reference only asserts here
        if(iter_kind_ != kind) {
            if(pos_ != 0) BNS_RUNTIME_ERROR("Encoder: one kind of next_*() call per assigned string");
            // next_kmer is the window-of-one case of next_minimizer: the same comb with w = c
            const Spacer sp = kind == 0 ? Spacer(sp_.k_, 0, sp_.sub1()) : sp_;
            if(!iter_h_[kind]) iter_h_[kind] = detail::open_handle(sp, ScoreType::id, kind == 2, BNS_API_ITER);
            bns_b200_t *h = iter_h_[kind]->h;
            const u64 npos = l_ - sp_.c_ + 1, lead = std::min<u64>(npos, sp.w_ - sp.c_);   // calls before the first full window
            iter_.assign(npos + 1, ENCODE_OVERFLOW);
            const u64 offs[2] = {0, l_}, ooffs[2] = {0, npos - lead + 1};
            u32 count = 0;
            detail::check(h, bns_b200_encode_batch(h, s_, offs, 1, iter_.data() + lead, ooffs, &count), "bns_b200_encode_batch");
            if(count != npos - lead) BNS_RUNTIME_ERROR("Encoder: unexpected stream length from the device");
            iter_kind_ = kind;
        }
        return iter_[pos_++];
    }
    template <typename F>
    void run(const F &fn, const char *str, u64 l, bool path_api) {
        bns_b200_t *h = get(path_api)->h;
        const u64 offs[2] = {0, l};
        const u64 bound = bns_b200_encode_bound(h, l);
        const u64 ooffs[2] = {0, bound};
        kmers_.resize(bound + 1);
        u32 count = 0;
        detail::check(h, bns_b200_encode_batch(h, str, offs, 1, kmers_.data(), ooffs, &count), "bns_b200_encode_batch");
        for(u32 i = 0; i < count; ++i) fn(kmers_[i]);
    }
public:
    Spacer sp_;
    static constexpr u64 ENCODE_OVERFLOW = ~u64(0);
    Encoder(const Spacer &sp, bool canonicalize = true) : canonicalize_(canonicalize && sp.unspaced()), sp_(sp) {}   // encoder.h:148-150
    // A copy is another worker's encoder (classifier.h:258): it opens its own device contexts at its first use instead of
    // sharing the original's (a context runs one call at a time).
    Encoder(const Encoder &o) : canonicalize_(o.canonicalize_), sp_(o.sp_) {}
    Encoder &operator=(const Encoder &o) {
        if(this != &o) {
            str_.reset(); path_.reset(); forced_[0].reset(); forced_[1].reset();
            for(auto &h : iter_h_) h.reset();
            canonicalize_ = o.canonicalize_; sp_ = o.sp_; s_ = nullptr; l_ = pos_ = 0; iter_kind_ = -1;
        }
        return *this;
    }
    explicit Encoder(unsigned k, bool canonicalize = true) : Encoder(Spacer(k), canonicalize) {}
    bool canonicalize() const { return canonicalize_; }
    u32 k() const { return sp_.k_; }
    // Encoder::for_each(fn, str, l), encoder.h:416-442
    template <typename F> void for_each(const F &fn, const char *str, u64 l) { run(fn, str, l, false); }
    // for_each_canon / for_each_uncanon on one record, encoder.h:448-464
    template <typename F> void for_each_record(const F &fn, const char *str, u64 l) { run(fn, str, l, true); }
    // for_each(fn, path), encoder.h:511-530: every record of a FASTA/FASTQ(.gz/.xz/.bz2/.zst) file through the record overloads
    template <typename F> void for_each(const F &fn, const char *path) {
        detail::KSeq ks(path);
        for_each_records_of(fn, ks, get(true)->h);
    }
    template <typename F> void for_each(const F &fn, const std::string &path) { for_each(fn, path.c_str()); }      // :507-510
    template <typename F> void for_each(const F &fn, gzFile fp) {                                                  // :497-506
        detail::KSeq ks(fp);
        for_each_records_of(fn, ks, get(true)->h);
    }
    // every file of a list, encoder.h:531-545
    template <typename F> void for_each(const F &fn, const std::vector<std::string> &paths) { for(const auto &p : paths) for_each(fn, p.c_str()); }
    // for_each_canon / for_each_uncanon over a file, encoder.h:448-496: the canonical / uncanonical bodies whatever canonicalize_
    // says. (A spaced seed stays uncanonical here, as the constructor decided; the reference would canonicalise it in
    // for_each_canon, which nothing on the classify path does.)
    template <typename F> void for_each_canon(const F &fn, const char *path) { for_each_as(fn, path, true); }
    template <typename F> void for_each_uncanon(const F &fn, const char *path) { for_each_as(fn, path, false); }
    // ---- the call-by-call surface, encoder.h:201-206,594-628 ---------------------------------------------------------
    // assign() borrows the string. The first next_*() call after it computes the value of that call for EVERY position of
    // the string on the device (one BNS_API_ITER encode: nothing filtered, ENCODE_OVERFLOW where the reference returns it --
    // an invalid base under the comb, a window that is not full yet, an invalid k-mer winning its window) and the calls
    // hand them out one by one. One kind of call per assigned string (the reference shares pos_ and qmap_ between them).
    void assign(const char *s, u64 l) { s_ = s; l_ = l; pos_ = 0; iter_kind_ = -1; }
    int has_next_kmer() const { return (pos_ + sp_.c_ - 1) < l_; }                       // encoder.h:594-597
    u64 next_kmer() { return iter_value(0); }                                            // kmer(pos_++), :601-604
    u64 next_minimizer() { return iter_value(1); }                                       // :616-621
    u64 next_canonicalized_minimizer() { return iter_value(2); }                         // :622-628
    // batched form of the string overload: fn(sequence index, kmer)
    template <typename F> void for_each_batch(const F &fn, const char *bases, const u64 *offsets, u64 n) {
        bns_b200_t *h = get(false)->h;
        std::vector<u64> oo(n + 1, 0);
        for(u64 i = 0; i < n; ++i) oo[i + 1] = oo[i] + bns_b200_encode_bound(h, offsets[i + 1] - offsets[i]);
        kmers_.resize(oo[n] + 1);
        std::vector<u32> counts(n);
        detail::check(h, bns_b200_encode_batch(h, bases, offsets, n, kmers_.data(), oo.data(), counts.data()), "bns_b200_encode_batch");
        for(u64 i = 0; i < n; ++i) for(u32 j = 0; j < counts[i]; ++j) fn(i, kmers_[oo[i] + j]);
    }
};

// ---- taxonomy --------------------------------------------------------------------------------------------------
struct TaxMap { std::vector<tax_t> child, parent; size_t size() const { return child.size(); } };
inline TaxMap *build_parent_map(const char *fn) {                   // util.h:766-785
    std::FILE *fp = std::fopen(fn, "r");
    if(!fp) BNS_RUNTIME_ERROR(std::string("Failed to create taxmap from ") + fn);
    auto ret = std::make_unique<TaxMap>();
    char *line = nullptr; size_t cap = 0; ssize_t len;
    while((len = getline(&line, &cap, fp)) >= 0) {
        if(len && line[len - 1] == '\n') line[len - 1] = 0;
        switch(line[0]) { case '\n': case '\0': case '#': continue; }
        const char *p = std::strchr(line, '|');
        ret->child.push_back((tax_t)std::atoi(line));
        ret->parent.push_back(p ? (tax_t)std::atoi(p + 2) : tax_t(-1));
    }
    std::free(line);
    std::fclose(fp);
    ret->child.push_back(1); ret->parent.push_back(0);             // "Root of the tree"
    if(ret->size() < 2) BNS_RUNTIME_ERROR(std::string("Failed to create taxmap from ") + fn);
    return ret.release();
}

// ---- Database --------------------------------------------------------------------------------------------------
// File layout the reference intends (database.h:81-102 + khash_write_impl util.h:281-296):
//   u32 k, u32 w, (k-1) x u8 gaps, u64 n_buckets, u64 n_occupied, u64 size, u64 upper_bound,
//   u32 flags[max(1, n_buckets/16)], u64 keys[n_buckets], u32 vals[n_buckets]          (little endian, no padding)
// i.e. the raw khash_t(c) arrays (Wang64 hash, triangular probing, 2 flag bits per bucket, khash64.h:169-263).
struct DatabaseImpl {
    u32 k_ = 0, w_ = 0;
    spvec_t s_;                                  // gaps (before the Spacer's +1)
    u64 n_buckets = 0, n_occupied = 0, size = 0, upper_bound = 0;
    std::vector<u32> flags;
    std::vector<u64> keys;
    std::vector<u32> vals;

    static u64 wang64(u64 key) {                 // khash64.h:202-211
        key = (~key) + (key << 21); key = key ^ (key >> 24);
        key = (key + (key << 3)) + (key << 8); key = key ^ (key >> 14);
        key = (key + (key << 2)) + (key << 4); key = key ^ (key >> 28);
        return key + (key << 31);
    }
    bool exists(u64 i) const { return ((flags[i >> 4] >> ((i & 0xfU) << 1)) & 3u) == 0; }
    DatabaseImpl() = default;
    explicit DatabaseImpl(const char *path) {    // database.h:33-56 (with the fread / popen defects of App. B-1 not reproduced)
        gzFile fp = gzopen(path, "rb");
        if(!fp) BNS_RUNTIME_ERROR(std::string("Could not open database at ") + path);
        auto rd = [&](void *p, size_t n) {
            size_t got = 0;
            while(got < n) {
                const unsigned want = (unsigned)std::min<size_t>(n - got, 1u << 30);
                const int r = gzread(fp, (char *)p + got, want);
                if(r <= 0) { gzclose(fp); BNS_RUNTIME_ERROR("Could not read from database file"); }
                got += (size_t)r;
            }
        };
        rd(&k_, 4); rd(&w_, 4);
        if(k_ < 1 || k_ > 32) { gzclose(fp); BNS_RUNTIME_ERROR("database: bad k"); }
        std::vector<u8> g(k_ - 1);
        if(k_ > 1) rd(g.data(), k_ - 1);
        s_.assign(g.begin(), g.end());
        rd(&n_buckets, 8); rd(&n_occupied, 8); rd(&size, 8); rd(&upper_bound, 8);
        flags.resize(n_buckets < 16 ? 1 : n_buckets >> 4);
        keys.resize(n_buckets); vals.resize(n_buckets);
        rd(flags.data(), flags.size() * 4);
        if(n_buckets) { rd(keys.data(), n_buckets * 8); rd(vals.data(), n_buckets * 4); }
        gzclose(fp);
    }
    // build the khash arrays from distinct (key, value) pairs: kh_resize to hold n at load <= 0.77, then kh_put each
    void assign(u32 k, u32 w, const spvec_t &gaps, const u64 *ks, const u32 *vs, u64 n) {
        k_ = k; w_ = w; s_ = gaps.size() ? gaps : spvec_t(k - 1, 0);
        n_buckets = 4;
        while((u64)(n_buckets * 0.77 + 0.5) <= n) n_buckets <<= 1;
        upper_bound = (u64)(n_buckets * 0.77 + 0.5);
        flags.assign(n_buckets < 16 ? 1 : n_buckets >> 4, 0xaaaaaaaau);
        keys.assign(n_buckets, 0); vals.assign(n_buckets, 0);
        const u64 mask = n_buckets - 1;
        size = 0;
        for(u64 j = 0; j < n; ++j) {
            u64 i = wang64(ks[j]) & mask, step = 0;
            while(exists(i) && keys[i] != ks[j]) i = (i + (++step)) & mask;
            if(!exists(i)) { flags[i >> 4] &= ~(3u << ((i & 0xfU) << 1)); keys[i] = ks[j]; ++size; }
            vals[i] = vs[j];
        }
        n_occupied = size;
    }
    void write(const char *fn, bool write_gz = false) const {        // database.h:81-102
        auto fail = [&]() { BNS_RUNTIME_ERROR(std::string("Error writing database to ") + fn); };
        std::vector<u8> g(s_.begin(), s_.end());
        g.resize(k_ ? k_ - 1 : 0, 0);
        if(write_gz) {
            gzFile fp = gzopen(fn, "wb");
            if(!fp) fail();
            auto wr = [&](const void *p, size_t n) {
                size_t put = 0;
                while(put < n) {
                    const int r = gzwrite(fp, (const char *)p + put, (unsigned)std::min<size_t>(n - put, 1u << 30));
                    if(r <= 0) { gzclose(fp); fail(); }
                    put += (size_t)r;
                }
            };
            wr(&k_, 4); wr(&w_, 4); if(!g.empty()) wr(g.data(), g.size());
            wr(&n_buckets, 8); wr(&n_occupied, 8); wr(&size, 8); wr(&upper_bound, 8);
            wr(flags.data(), flags.size() * 4); wr(keys.data(), keys.size() * 8); wr(vals.data(), vals.size() * 4);
            gzclose(fp);
            return;
        }
        std::FILE *fp = std::fopen(fn, "wb");
        if(!fp) fail();
        auto wr = [&](const void *p, size_t n) { if(n && std::fwrite(p, 1, n, fp) != n) { std::fclose(fp); fail(); } };
        wr(&k_, 4); wr(&w_, 4); wr(g.data(), g.size());
        wr(&n_buckets, 8); wr(&n_occupied, 8); wr(&size, 8); wr(&upper_bound, 8);
        wr(flags.data(), flags.size() * 4); wr(keys.data(), keys.size() * 8); wr(vals.data(), vals.size() * 4);
        std::fclose(fp);
    }
};
// The raw khash_t(c) as the reference declares it (struct kh_c_s, khash64.h:213-219, khint_t = u64): what Database<T>::db_
// points to and what ClassifierGeneric's constructor takes (classifier.h:155). Here it is a VIEW of the vectors above.
struct kh_c_t { u64 n_buckets, size, n_occupied, upper_bound; u32 *flags; u64 *keys; u32 *vals; };
// Database<khash_t(c)> (database.h:17): the reference's spelling. T must be kh_c_t; db_ views the arrays of this object.
template <typename T = kh_c_t>
struct Database : DatabaseImpl {
    static_assert(std::is_same<T, kh_c_t>::value, "Database<T>: only khash_t(c) databases exist on the classify path");
    kh_c_t view_{};
    T *db_ = nullptr;
    void sync() {                                // after the arrays changed (assign)
        view_ = kh_c_t{n_buckets, size, n_occupied, upper_bound, flags.data(), keys.data(), vals.data()};
        db_ = &view_;
    }
    Database() { sync(); }
    explicit Database(const char *path) : DatabaseImpl(path) { sync(); }
    Database(const Database &o) : DatabaseImpl(o) { sync(); }
    Database &operator=(const Database &o) { DatabaseImpl::operator=(o); sync(); return *this; }
    void assign(u32 k, u32 w, const spvec_t &gaps, const u64 *ks, const u32 *vs, u64 n) { DatabaseImpl::assign(k, w, gaps, ks, vs, n); sync(); }
};

// ---- reads -----------------------------------------------------------------------------------------------------
struct bseq1_t {                                  // kseq_declare.h:40-44 (strings own their storage here)
    int l_seq = 0, id = 0;
    std::string name, comment, seq, qual, sam;
};
inline void trim_readno(std::string &s) {         // kseq_declare.h:106-110
    if(s.size() > 2 && s[s.size() - 2] == '/' && std::isdigit((unsigned char)s.back())) s.resize(s.size() - 2);
}
// bseq_read, kseq_declare.h:112-145: records until the batch holds >= chunk_size BASES (and an even count)
inline bool bseq_read(int chunk_size, std::vector<bseq1_t> &seqs, detail::KSeq *ks, detail::KSeq *ks2) {
    seqs.clear();
    long size = 0;
    auto take = [&](detail::KSeq *k) {
        seqs.emplace_back();
        bseq1_t &s = seqs.back();
        trim_readno(k->name);
        s.name = k->name; s.comment = k->comment; s.seq = k->seq; s.qual = k->qual;
        s.l_seq = (int)s.seq.size(); s.id = (int)seqs.size() - 1;
        size += s.l_seq;
    };
    while(ks->read() >= 0) {
        if(ks2 && ks2->read() < 0) { std::fprintf(stderr, "[W::%s] the 2nd file has fewer sequences.\n", __func__); break; }
        take(ks);
        if(ks2) take(ks2);
        if(size >= chunk_size && (seqs.size() & 1) == 0) break;
    }
    if(size == 0 && ks2 && ks2->read() >= 0) std::fprintf(stderr, "[W::%s] the 1st file has fewer sequences.\n", __func__);
    return !seqs.empty();
}

// ---- classifier ------------------------------------------------------------------------------------------------
enum output_format : int { KRAKEN = 1, FASTQ = 2, EMIT_ALL = 4 };   // classifier.h:23-27

template <typename ScoreType>
struct ClassifierGeneric {
    const Spacer sp_;
    Encoder<ScoreType> enc_;
    u32 nt_ : 16;
    u32 output_flag_ : 16;
    std::shared_ptr<detail::Handle> h_;
    std::vector<std::shared_ptr<detail::Handle>> replicas_;      // contexts on GPUs 1 .. n-1 holding a copy of the database (set_gpus)
    int n_gpus_ = 1;
    bool tax_loaded_ = false;
    bool exit_follows_ = false;                                  // the caller ends the process right after process_dataset (the CLI):
                                                                 // pinned buffers are left to the operating system instead of being
                                                                 // unpinned one by one (0.03 - 0.24 s for the three ring slots)
    void set_exit_follows(bool v) { exit_follows_ = v; }
    std::atomic<u64> t_device_ns_{0}, t_format_ns_{0}, t_join_ns_{0};   // BNS_B200_VERBOSE: where classify_views spent its time (all workers)
    std::atomic<u32> runs_per_record_hint_{4};                   // run-buffer entries per record the batches so far needed (classify_views)
    void set_emit_all(bool s) { if(s) output_flag_ |= EMIT_ALL; else output_flag_ &= ~EMIT_ALL; }
    void set_emit_kraken(bool s) { if(s) output_flag_ |= KRAKEN; else output_flag_ &= ~KRAKEN; }
    void set_emit_fastq(bool s) { if(s) output_flag_ |= FASTQ; else output_flag_ &= ~FASTQ; }
    int get_emit_all() const { return output_flag_ & EMIT_ALL; }
    int get_emit_kraken() const { return output_flag_ & KRAKEN; }
    int get_emit_fastq() const { return output_flag_ & FASTQ; }
    // classifier.h:155-166; `map` = the database whose raw khash arrays go to the device
    ClassifierGeneric(const DatabaseImpl &map, const spvec_t &spaces, u8 k, u16 wsz, int num_threads = 16, bool emit_all = true,
                      bool emit_fastq = true, bool emit_kraken = false, bool canonicalize = true)
        : sp_(k, wsz, spaces), enc_(sp_, canonicalize), nt_(num_threads > 0 ? (u16)num_threads : (u16)std::max(1u, std::thread::hardware_concurrency())), output_flag_(0) {
        set_emit_all(emit_all); set_emit_fastq(emit_fastq); set_emit_kraken(emit_kraken);
        h_ = detail::open_handle(sp_, ScoreType::id, enc_.canonicalize(), BNS_API_STRING, 0);
        detail::check(h_->h, bns_b200_load_table(h_->h, map.keys.data(), map.vals.data(), map.flags.data(), map.n_buckets),
                      "bns_b200_load_table");
    }
    // the reference's own argument list: the raw khash_t(c) (classifier.h:155; kh_c_t below)
    ClassifierGeneric(const kh_c_t *map, const spvec_t &spaces, u8 k, u16 wsz, int num_threads = 16, bool emit_all = true,
                      bool emit_fastq = true, bool emit_kraken = false, bool canonicalize = true)
        : sp_(k, wsz, spaces), enc_(sp_, canonicalize), nt_(num_threads > 0 ? (u16)num_threads : (u16)std::max(1u, std::thread::hardware_concurrency())), output_flag_(0) {
        set_emit_all(emit_all); set_emit_fastq(emit_fastq); set_emit_kraken(emit_kraken);
        h_ = detail::open_handle(sp_, ScoreType::id, enc_.canonicalize(), BNS_API_STRING, 0);
        detail::check(h_->h, bns_b200_load_table(h_->h, map->keys, map->vals, map->flags, map->n_buckets), "bns_b200_load_table");
    }
    // Reads shard over n GPUs of this process: GPU 0 holds the table; the others receive a copy with one NCCL
    // broadcast per segment (bns_b200_replicate) once the taxonomy is there, and process_dataset deals its chunks round-robin.
    void set_gpus(int n) {
        if(n < 1) BNS_RUNTIME_ERROR("set_gpus: need at least one GPU");
        if(!replicas_.empty() && n != n_gpus_) BNS_RUNTIME_ERROR("set_gpus: replicas already exist");
        n_gpus_ = n;
        if(tax_loaded_) replicate();
    }
    int n_gpus() const { return n_gpus_; }
    bns_b200_t *gpu(int g) const { return g == 0 ? h_->h : replicas_[(size_t)g - 1]->h; }
    void replicate() {
        if(n_gpus_ <= 1 || !replicas_.empty()) return;
        std::vector<bns_b200_t *> all(1, h_->h);
        for(int g = 1; g < n_gpus_; ++g) {
            replicas_.push_back(detail::open_handle(sp_, ScoreType::id, enc_.canonicalize(), BNS_API_STRING, g));
            all.push_back(replicas_.back()->h);
        }
        detail::check(h_->h, bns_b200_replicate(all.data(), n_gpus_, 0), "bns_b200_replicate");
    }
    void load_taxonomy(const TaxMap *t) {
        detail::check(h_->h, bns_b200_load_taxonomy(h_->h, t->child.data(), t->parent.data(), t->size()), "bns_b200_load_taxonomy");
        tax_loaded_ = true;
        replicate();
    }
    u64 n_classified() const {                                    // classifier.h:170
        u64 n = 0;
        for(int g = 0; g < (replicas_.empty() ? 1 : n_gpus_); ++g) { bns_b200_stats s; bns_b200_stats_get(gpu(g), &s); n += s.n_classified; }
        return n;
    }
    u64 n_unclassified() const {                                  // :171
        u64 n = 0;
        for(int g = 0; g < (replicas_.empty() ? 1 : n_gpus_); ++g) { bns_b200_stats s; bns_b200_stats_get(gpu(g), &s); n += s.n_unclassified; }
        return n;
    }
};
using Classifier = ClassifierGeneric<score::Lex>;

namespace detail {
struct ReadView {                                                     // what the emitters need of a bseq1_t
    const char *name, *seq, *qual; int l_seq;
    int l_name = -1;                                                  // < 0: name is NUL-terminated
    void put_name(std::string &s) const { if(l_name < 0) s += name; else s.append(name, (size_t)l_name); }
};
// decimal text without snprintf (a Kraken line holds a dozen numbers; this is most of the formatter's time)
inline void put_u64(std::string &s, u64 x) {
    char b[24];
    int at = 24;
    do { b[--at] = (char)('0' + x % 10); x /= 10; } while(x);
    s.append(b + at, (size_t)(24 - at));
}
inline void put_u(std::string &s, u32 x) { put_u64(s, x); }
inline void put_i(std::string &s, long x) {
    if(x < 0) { s.push_back('-'); put_u64(s, (u64)0 - (u64)x); } else put_u64(s, (u64)x);
}
inline void append_taxa_run(tax_t last, u32 run, std::string &s) {                // classifier.h:30-43
    if(last == 0) s.push_back('U'); else if(last == (tax_t)-1) s.push_back('A'); else put_u(s, last);
    s.push_back(':'); put_u(s, run); s.push_back('\t');
}
inline void append_taxa_runs(tax_t taxon, const tax_t *taxa, u32 n, std::string &s) {   // :46-61
    if(taxon) {
        tax_t last = taxa[0]; u32 run = 1;
        for(u32 i = 1; i != n; ++i) {
            if(taxa[i] == last) ++run;
            else { append_taxa_run(last, run, s); last = taxa[i]; run = 1; }
        }
        append_taxa_run(last, run, s);
        s.back() = '\n';
    } else s.append("0:0\n", 4);
}
// the same line from run-length encoded hits, (taxid << 32 | run length) words in k-mer order (bns_b200_classify_batch_runs)
inline void append_taxa_runs_rle(tax_t taxon, const u64 *runs, u32 n_runs, std::string &s) {
    if(taxon) {
        for(u32 i = 0; i != n_runs; ++i) append_taxa_run((tax_t)(runs[i] >> 32), (u32)runs[i], s);
        s.back() = '\n';
    } else s.append("0:0\n", 4);
}
inline void append_counts(u32 count, char ch, std::string &s) {                   // :63-70
    if(count) { s.push_back(ch); s.push_back(':'); put_u(s, count); s.push_back('\t'); }
}
inline void append_kraken_classification(const tax_t *taxa, u32 ntaxa, tax_t taxon, u32 ambig, u32 missing,
                                         const ReadView &bs, std::string &s, const u64 *runs = nullptr, u32 n_runs = 0) {     // :112-129
    s.push_back(taxon ? 'C' : 'U'); s.push_back('\t');
    bs.put_name(s); s.push_back('\t');
    put_u(s, taxon); s.push_back('\t');
    put_i(s, bs.l_seq); s.push_back('\t');
    append_counts(missing, 'M', s); append_counts(ambig, 'A', s);
    if(runs) append_taxa_runs_rle(taxon, runs, n_runs, s); else append_taxa_runs(taxon, taxa, ntaxa, s);
}
inline void append_fastq_classification(const tax_t *taxa, u32 ntaxa, tax_t taxon, u32 ambig, u32 missing,
                                        const ReadView *bs, std::string &s, int verbose, int is_paired,
                                        const u64 *runs = nullptr, u32 n_runs = 0) {   // :72-108
    bs->put_name(s); s.push_back(' ');
    const size_t cms = s.size();
    s.push_back(taxon == 0 ? 'U' : 'C'); s.push_back('\t');
    put_u(s, taxon); s.push_back('\t');
    put_i(s, bs->l_seq); s.push_back('\t');
    append_counts(missing, 'M', s); append_counts(ambig, 'A', s);
    if(verbose) { if(runs) append_taxa_runs_rle(taxon, runs, n_runs, s); else append_taxa_runs(taxon, taxa, ntaxa, s); } else s.back() = '\n';
    const std::string cm = s.substr(cms);      // the reference keeps raw pointers here and breaks on realloc
    s.append(bs->seq, bs->l_seq); s.append("\n+\n", 3);
    s.append(bs->qual ? bs->qual : bs->seq, bs->l_seq); s.push_back('\n');
    if(is_paired) {
        const ReadView *b2 = bs + 1;
        b2->put_name(s); s.push_back(' ');
        s += cm; s.push_back('\n');
        s.append(b2->seq, b2->l_seq); s.append("\n+\n", 3);
        s.append(b2->qual ? b2->qual : b2->seq, b2->l_seq); s.push_back('\n');
    }
}

// One GPU call for a batch laid out as concatenated bases + offsets, then classify_seq's epilogue per record
// (classifier.h:232-246). view_at(i): what the emitters need of read i (mates interleaved), made where it is used -- on the
// formatting threads -- instead of as an array per batch.
// A grow-only pinned buffer per calling thread for what a batch call copies back (taxa, counts, run lists: ~35 bytes per read).
// Into pageable memory those copies are staged by the driver at a few GB/s and were most of a batch call's time.
struct PinnedScratch {
    void *p = nullptr; size_t cap = 0;
    bool keep = false;                                               // the process is about to end: leave the buffer to the operating system
    ~PinnedScratch() { if(p && !keep) bns_b200_host_free(p); }
    void *get(size_t bytes) {                                        // nullptr if pinned memory cannot be had: the caller uses the heap
        if(bytes > cap) {
            if(p) bns_b200_host_free(p);
            p = nullptr; cap = 0;
            const size_t want = bytes + bytes / 4 + 4096;
            if(bns_b200_host_alloc(&p, want) != 0) { p = nullptr; return nullptr; }
            cap = want;
        }
        return p;
    }
};
// parts_out: the formatted slices are handed over as they are (process_dataset writes them one after the other) instead of being
// appended to cks
template <typename ScoreType, typename ViewAt>
void classify_views(ClassifierGeneric<ScoreType> &c, const char *bases, const u64 *offs, const ViewAt &view_at, unsigned n_reads,
                    int is_paired, std::string &cks, bns_b200_t *h = nullptr, unsigned max_threads = 0, std::mutex *device_mu = nullptr,
                    std::vector<std::string> *parts_out = nullptr) {
    const unsigned inc = is_paired ? 2 : 1, nrec = n_reads / inc;
    if(!nrec) return;
    const auto t_in = std::chrono::steady_clock::now();
    auto ns_since = [](std::chrono::steady_clock::time_point t) { return (u64)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t).count(); };
    // device_mu: a context runs one call at a time; two host workers that share a GPU take turns for the call and format their
    // batches side by side
    std::unique_lock<std::mutex> device_lock;
    if(device_mu) device_lock = std::unique_lock<std::mutex>(*device_mu);
    if(!h) h = c.h_->h;                                               // the context (GPU) this batch runs on
    // The ordered hit list is only printed by the Kraken run lists, and the k-mer count of mate 1 only differs from
    // hits + missing for pairs: without them the library runs its lean kernel and copies 12 bytes per record back.
    const bool need_taxa = (c.output_flag_ & KRAKEN) != 0;
    // result arrays the library fills (no need to zero them first), in this thread's pinned scratch where that can be had
    thread_local PinnedScratch res_scratch, runs_scratch;
    res_scratch.keep = runs_scratch.keep = c.exit_follows_;
    const size_t res_bytes = (size_t)nrec * 5 * sizeof(u32) + (need_taxa ? (size_t)nrec * sizeof(u64) : 0) + 16;
    std::unique_ptr<char[]> res_heap;
    char *res_mem = (char *)res_scratch.get(res_bytes);
    if(!res_mem) { res_heap.reset(new char[res_bytes]); res_mem = res_heap.get(); }
    u64 *const run_pos_p = (u64 *)res_mem;                            // (first: 8-byte aligned)
    u32 *const res_p = (u32 *)(res_mem + (need_taxa ? (size_t)nrec * sizeof(u64) : 0));
    struct Arr { u32 *p; u32 *data() const { return p; } u32 &operator[](size_t i) const { return p[i]; } };
    struct Arr64 { u64 *p; u64 *data() const { return p; } u64 &operator[](size_t i) const { return p[i]; } };
    const Arr taxon{res_p}, nhit{res_p + nrec}, nmiss{res_p + 2 * (size_t)nrec}, mate1{res_p + 3 * (size_t)nrec}, nruns{res_p + 4 * (size_t)nrec};
    const Arr64 run_pos{need_taxa ? run_pos_p : nullptr};
    std::unique_ptr<u64[]> runs_heap;
    u64 *runs = nullptr;
    if(need_taxa) {
        // run lists: encoded on the device, 8 bytes per run back instead of 4 per k-mer window slot. The number of runs is
        // not known beforehand: the buffer is sized from what earlier batches of this classifier needed (entries per record,
        // carried in the classifier), and a call that does not fit reports how much it needs and has counted nothing.
        u64 windows = 0;
        for(unsigned r = 0; r < nrec; ++r) windows += (offs[(r + 1) * inc] - offs[r * inc]) + 2;
        const u64 bound = windows + ((u64)1 << 21);                       // always enough (bonsai_b200.h)
        u64 cap = std::min<u64>(bound, (u64)nrec * c.runs_per_record_hint_.load() + ((u64)1 << 20)), total = 0;
        for(;;) {
            runs = (u64 *)runs_scratch.get((cap ? cap : 1) * sizeof(u64));
            if(!runs) { runs_heap.reset(new u64[cap ? cap : 1]); runs = runs_heap.get(); }
            const int rc = bns_b200_classify_batch_runs(h, bases, offs, nrec * inc, is_paired, taxon.data(), nhit.data(), nmiss.data(),
                                                        is_paired ? mate1.data() : nullptr, runs, cap, run_pos.data(), nruns.data(), &total);
            if(rc == BNS_E_CAPACITY && cap < bound) { cap = std::min<u64>(bound, std::max<u64>(cap * 2, total + total / 2)); continue; }
            check(h, rc, "bns_b200_classify_batch_runs");
            break;
        }
        const u32 per = (u32)std::min<u64>(1u << 20, total / nrec + 2);
        u32 seen = c.runs_per_record_hint_.load();
        while(per > seen && !c.runs_per_record_hint_.compare_exchange_weak(seen, per)) {}
    } else
        check(h, bns_b200_classify_batch_ex(h, bases, offs, nrec * inc, is_paired, taxon.data(), nhit.data(), nmiss.data(),
                                            nullptr, nullptr, is_paired ? mate1.data() : nullptr), "bns_b200_classify_batch");
    const u32 comb = c.sp_.c_;
    // classify_seq's epilogue (text) per record. The reference formats on its worker threads (-p, kt_for_helper,
    // classifier.h:254-266); here -p threads format contiguous slices of the batch and the slices are joined in order.
    if(device_lock.owns_lock()) device_lock.unlock();
    c.t_device_ns_ += ns_since(t_in);                                 // includes waiting for the other worker's call
    if(!(c.output_flag_ & (FASTQ | KRAKEN))) return;                  // nothing is printed (-K without -f): the counters are all there is
    const auto t_fmt = std::chrono::steady_clock::now();
    auto format_range = [&](unsigned r_lo, unsigned r_hi, std::string &out) {
        for(unsigned r = r_lo; r < r_hi; ++r) {
            const ReadView pair[2] = {view_at((size_t)r * inc), is_paired ? view_at((size_t)r * inc + 1) : ReadView{}};
            const ReadView *b = pair;
            // classifier.h:232: unsigned ambig_count(l_seq - c + 1 - taxa.size() - missing_count), evaluated after mate 1
            u32 ambig = (u32)((u64)(u32)((u32)b->l_seq - comb + 1) - (u64)(is_paired ? mate1[r] : nhit[r] + nmiss[r]));
            if(is_paired) ambig += (u32)((u64)(u32)((u32)(b + 1)->l_seq - (comb - 1)) - (u64)nhit[r] - nmiss[r]);   // :235
            if(c.get_emit_all() || taxon[r]) {
                const u64 *rn = need_taxa ? runs + run_pos[r] : nullptr;
                const u32 nrn = need_taxa ? nruns[r] : 0;
                if(c.output_flag_ & FASTQ)
                    append_fastq_classification(nullptr, nhit[r], taxon[r], ambig, nmiss[r], b, out, c.get_emit_kraken(), is_paired, rn, nrn);
                else if(c.output_flag_ & KRAKEN)
                    append_kraken_classification(nullptr, nhit[r], taxon[r], ambig, nmiss[r], *b, out, rn, nrn);
            }
        }
    };
    const unsigned nthreads = std::max(1u, std::min<unsigned>(max_threads ? max_threads : c.nt_, nrec / 256 + 1));
    if(nthreads == 1) {
        format_range(0, nrec, cks);
        c.t_format_ns_ += ns_since(t_fmt);
        if(parts_out) { parts_out->push_back(std::move(cks)); cks.clear(); }
        return;
    }
    std::vector<std::string> parts(nthreads);
    std::vector<std::thread> pool;
    // a line is a few dozen bytes (plus the sequence and quality in FASTQ style): one allocation per slice instead of doublings
    const size_t per_rec = (c.output_flag_ & FASTQ) ? 96 + 2 * (size_t)((offs[nrec * inc] - offs[0]) / nrec + 2) * inc : ((c.output_flag_ & KRAKEN) ? 96 : 0);
    for(unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t] {
            const unsigned lo = (unsigned)((u64)nrec * t / nthreads), hi = (unsigned)((u64)nrec * (t + 1) / nthreads);
            parts[t].reserve((size_t)(hi - lo) * per_rec);
            format_range(lo, hi, parts[t]);
        });
    for(auto &th : pool) th.join();
    c.t_format_ns_ += ns_since(t_fmt);
    if(parts_out) { for(auto &part : parts) parts_out->push_back(std::move(part)); return; }
    const auto t_join = std::chrono::steady_clock::now();
    size_t total = cks.size();
    for(auto &part : parts) total += part.size();
    cks.reserve(total);
    for(auto &part : parts) cks += part;
    c.t_join_ns_ += ns_since(t_join);
}

// one record of a mapped plain file (parallel ingest, below): offsets into the mapping
struct RecRef { u64 name_off, seq_off, qual_off; u32 name_len, seq_len; };
// An array of trivially constructible elements whose resize() does not value-initialise them: the record indices are tens of
// millions of 40-byte entries that the -p threads fill in parallel right after -- a std::vector would first zero them (and
// first-touch their pages) on one thread, which was a quarter of the ingest's time.
template <class T>
struct RawVec {
    std::unique_ptr<T[]> p;
    size_t n = 0, cap = 0;
    void resize(size_t m) { if(m > cap) { p.reset(new T[m]); cap = m; } n = m; }     // contents are NOT kept and NOT initialised
    void clear() { n = 0; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T *data() { return p.get(); }
    const T *data() const { return p.get(); }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
    const T *begin() const { return p.get(); }
    const T *end() const { return p.get() + n; }
};
// A batch parsed straight into PINNED host memory (bns_b200_host_alloc): the library DMAs from it without staging.
struct PinnedBatch {
    char *bases = nullptr; size_t cap_bases = 0, n_bases = 0;
    u64 *offs = nullptr; size_t cap_offs = 0, n = 0;
    std::vector<std::string> names, quals;
    std::vector<char> has_qual;
    bool keep_qual = true;             // qualities are only printed by the FASTQ-style output
    RawVec<RecRef> refs;        // parallel-ingest batches: names / qualities stay in the file mapping `map`
    const char *map = nullptr, *map2 = nullptr;                        // map2: the mates' file (records at odd indices)
    std::shared_ptr<char> keep, keep2;                                 // gzip input: the inflated windows `map` / `map2` point into
    const char *map_of(size_t i) const { return (map2 && (i & 1)) ? map2 : map; }
    PinnedBatch() = default;
    PinnedBatch(const PinnedBatch &) = delete;
    ~PinnedBatch() { bns_b200_host_free(bases); bns_b200_host_free(offs); }
    template <class T> static void grow(T *&p, size_t &cap, size_t used, size_t need) {
        if(need <= cap) return;
        size_t ncap = std::max<size_t>(need, cap ? cap * 2 : 1 << 16);
        void *np = nullptr;
        if(bns_b200_host_alloc(&np, ncap * sizeof(T))) BNS_RUNTIME_ERROR("pinned host allocation failed");
        if(used && p) std::memcpy(np, p, used * sizeof(T));
        bns_b200_host_free(p);
        p = (T *)np; cap = ncap;
    }
    void clear() { n_bases = 0; n = 0; names.clear(); quals.clear(); has_qual.clear(); map = map2 = nullptr; keep.reset(); keep2.reset(); }
    void reserve(size_t bases_hint) {                   // pinned allocations are slow: size the ring once per dataset
        // a batch ends with the record that takes it past bases_hint (bseq_read's rule): 1/16 of headroom covers reads of up to
        // 4 M bases at the default chunk; longer records and batches of reads shorter than 64 bases grow the buffers (slowly: pinned)
        grow(bases, cap_bases, n_bases, bases_hint + bases_hint / 16 + (1 << 16));
        grow(offs, cap_offs, n ? n + 1 : 0, bases_hint / 64 + 1024);
    }
    void push(KSeq *k) {
        trim_readno(k->name);
        grow(bases, cap_bases, n_bases, n_bases + k->seq.size() + 16);
        grow(offs, cap_offs, n + 1, n + 2);
        if(n == 0) offs[0] = 0;
        std::memcpy(bases + n_bases, k->seq.data(), k->seq.size());
        n_bases += k->seq.size();
        offs[++n] = n_bases;
        names.push_back(k->name);
        has_qual.push_back(keep_qual && !k->qual.empty());
        if(keep_qual) quals.push_back(k->qual); else quals.emplace_back();
    }
};
// The -p threads of the ingest as a fixed set: run(n, fn) executes fn(0 .. n-1) on them and on the calling thread and returns when
// all are done. (Spawning threads per window and per batch -- ~700 spawns for 8 M reads -- was a tenth of the reader's time.)
class WorkPool {
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, cvd_;
    const std::function<void(unsigned)> *fn_ = nullptr;
    unsigned n_tasks_ = 0, active_ = 0;
    std::atomic<unsigned> next_{0};
    u64 gen_ = 0;
    bool stop_ = false;
    void loop() {
        u64 seen = 0;
        for(;;) {
            const std::function<void(unsigned)> *fn;
            unsigned n;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if(stop_) return;
                seen = gen_; fn = fn_; n = n_tasks_;
            }
            for(unsigned t; (t = next_.fetch_add(1)) < n;) (*fn)(t);
            std::lock_guard<std::mutex> lk(m_);
            if(--active_ == 0) cvd_.notify_all();
        }
    }
  public:
    explicit WorkPool(unsigned n) { for(unsigned i = 1; i < n; ++i) th_.emplace_back([this] { loop(); }); }
    WorkPool(const WorkPool &) = delete;
    ~WorkPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_.notify_all();
        for(auto &t : th_) t.join();
    }
    unsigned size() const { return (unsigned)th_.size() + 1; }
    void run(unsigned n_tasks, const std::function<void(unsigned)> &fn) {
        if(n_tasks <= 1 || th_.empty()) { for(unsigned t = 0; t < n_tasks; ++t) fn(t); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn; n_tasks_ = n_tasks; next_.store(0); active_ = (unsigned)th_.size(); ++gen_;
        }
        cv_.notify_all();
        for(unsigned t; (t = next_.fetch_add(1)) < n_tasks;) fn(t);
        std::unique_lock<std::mutex> lk(m_);
        cvd_.wait(lk, [&] { return active_ == 0; });
    }
};
// ---- parallel ingest of plain (uncompressed) files in the simple form ---------------------------------------------
// kseq is a byte-at-a-time state machine: ~2 M reads/s on one thread, two orders of magnitude below the GPU. A plain file
// whose records are exactly 4 lines (FASTQ: @header / sequence / + / quality of the same length) or 2 lines (FASTA:
// >header / sequence) is mapped and indexed by -p threads instead, window by window; names and qualities stay in the
// mapping, sequences are copied into the pinned batch in parallel. kseq semantics are kept for such records (name = header
// up to the first white space, trim_readno). Anything else (gzip, multi-line records, empty lines, CR) takes the kseq path.
struct MappedFile {
    const char *p = nullptr; size_t n = 0; int fd = -1;
    explicit MappedFile(const char *path) {
        fd = ::open(path, O_RDONLY);
        if(fd < 0) return;
        struct stat st;
        if(::fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 4) { ::close(fd); fd = -1; return; }
        void *m = ::mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if(m == MAP_FAILED) { ::close(fd); fd = -1; return; }
        p = (const char *)m; n = (size_t)st.st_size;
        ::madvise(m, n, MADV_SEQUENTIAL);
    }
    MappedFile(const MappedFile &) = delete;
    ~MappedFile() { if(p) ::munmap((void *)p, n); if(fd >= 0) ::close(fd); }
    bool simple_candidate() const { return p && (p[0] == '@' || p[0] == '>') && !((unsigned char)p[0] == 0x1f && (unsigned char)p[1] == 0x8b); }
};
// end of the line starting at x (offset of its newline, or n)
inline size_t line_end(const char *p, size_t x, size_t n) {
    const void *q = x < n ? std::memchr(p + x, '\n', n - x) : nullptr;
    return q ? (size_t)((const char *)q - p) : n;
}
// first record start at or after x (x = 0 or any offset): a line starting with the header character whose record parses
inline size_t next_record_start(const char *p, size_t x, size_t n, bool fastq) {
    size_t ls = x;
    if(x) { ls = line_end(p, x - 1, n) + 1; }                              // the line start at or after x
    for(int tries = 0; ls < n && tries < 8; ++tries) {
        if(p[ls] == (fastq ? '@' : '>')) {
            if(!fastq) return ls;
            const size_t e1 = line_end(p, ls, n), e2 = line_end(p, e1 + 1, n);
            if(e2 + 1 < n && p[e2 + 1] == '+') {
                const size_t e3 = line_end(p, e2 + 1, n), e4 = line_end(p, e3 + 1, n);
                if(e4 - (e3 + 1) == e2 - (e1 + 1)) return ls;
            }
        }
        ls = line_end(p, ls, n) + 1;
    }
    return n;                                                              // none found nearby: the caller gives up the fast path
}
// records of [lo, hi) (both record starts or n). Returns false on anything outside the simple form.
inline bool index_range(const char *p, size_t lo, size_t hi, size_t n, bool fastq, std::vector<RecRef> &out) {
    size_t x = lo;
    while(x < hi) {
        if(p[x] != (fastq ? '@' : '>')) return false;
        const size_t e1 = line_end(p, x, n);
        if(e1 >= n) return false;
        size_t ne = x + 1;
        while(ne < e1 && !std::isspace((unsigned char)p[ne])) ++ne;
        RecRef r;
        r.name_off = x + 1; r.name_len = (u32)(ne - (x + 1));
        if(r.name_len > 2 && p[ne - 2] == '/' && std::isdigit((unsigned char)p[ne - 1])) r.name_len -= 2;   // trim_readno
        const size_t e2 = line_end(p, e1 + 1, n);
        if(e2 - (e1 + 1) > 0x7fffffffull) return false;
        r.seq_off = e1 + 1; r.seq_len = (u32)(e2 - (e1 + 1));
        if(r.seq_len == 0 || std::memchr(p + x, '\r', e2 - x)) return false;
        if(std::memchr(p + r.seq_off, '>', r.seq_len) || std::memchr(p + r.seq_off, '@', r.seq_len) || std::memchr(p + r.seq_off, '+', r.seq_len))
            return false;                                                  // kseq would end the sequence there
        r.qual_off = ~0ull;
        size_t next = e2 + 1;
        if(fastq) {
            if(e2 + 1 >= n || p[e2 + 1] != '+') return false;
            const size_t e3 = line_end(p, e2 + 1, n);
            if(e3 >= n) return false;
            const size_t e4 = line_end(p, e3 + 1, n);
            if(e4 - (e3 + 1) != r.seq_len) return false;
            r.qual_off = e3 + 1;
            next = e4 + 1;
        }
        out.push_back(r);
        x = next;
    }
    return x == hi || (x == n + 1 && hi == n);                            // a last line without a newline ends at n
}
#if defined(__x86_64__) && defined(__GNUC__)
// index_range with the line ends found 32 bytes at a time (one compare + movemask per block, every block of the text touched
// once) and the sequence line checked for the bytes kseq would stop at in one pass of four compares: ~3x the records per second
// of the memchr-per-line version on one thread. Same acceptance rules, same RecRefs.
struct NewlineScan {
    const char *p; size_t n;
    size_t blk = ~(size_t)0; unsigned mask = 0;          // newline bits of p[blk, blk + 32) not consumed yet
    __attribute__((target("avx2"))) size_t next(size_t x) {     // offset of the first '\n' at or after x, or n (x never goes backwards)
        if(x >= n) return n;
        if(blk == ~(size_t)0 || x >= blk + 32) { blk = x; mask = 0; load(); }
        else mask &= ~0u << (x - blk);
        while(mask == 0) {
            blk += 32;
            if(blk >= n) return n;
            load();
        }
        const size_t pos = blk + (size_t)__builtin_ctz(mask);
        mask &= mask - 1;
        return pos;
    }
    __attribute__((target("avx2"))) void load() {
        if(blk + 32 <= n) {
            const __m256i v = _mm256_loadu_si256((const __m256i *)(p + blk));
            mask = (unsigned)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('\n')));
        } else {
            mask = 0;
            for(size_t i = blk; i < n; ++i) if(p[i] == '\n') mask |= 1u << (i - blk);
        }
    }
};
// does s[0, len) hold a '>', '@', '+' or '\r' ?
__attribute__((target("avx2"))) inline bool seq_has_stop_byte(const char *s, size_t len) {
    const __m256i a = _mm256_set1_epi8('>'), b = _mm256_set1_epi8('@'), c = _mm256_set1_epi8('+'), d = _mm256_set1_epi8('\r');
    __m256i acc = _mm256_setzero_si256();
    size_t i = 0;
    for(; i + 32 <= len; i += 32) {
        const __m256i v = _mm256_loadu_si256((const __m256i *)(s + i));
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, a), _mm256_cmpeq_epi8(v, b)),
                                                   _mm256_or_si256(_mm256_cmpeq_epi8(v, c), _mm256_cmpeq_epi8(v, d))));
    }
    bool hit = _mm256_movemask_epi8(acc) != 0;
    for(; i < len; ++i) hit |= s[i] == '>' || s[i] == '@' || s[i] == '+' || s[i] == '\r';
    return hit;
}
__attribute__((target("avx2"))) inline bool index_range_avx2(const char *p, size_t lo, size_t hi, size_t n, bool fastq, std::vector<RecRef> &out) {
    NewlineScan nl{p, n};
    size_t x = lo;
    while(x < hi) {
        if(p[x] != (fastq ? '@' : '>')) return false;
        const size_t e1 = nl.next(x);
        if(e1 >= n) return false;
        size_t ne = x + 1;
        while(ne < e1 && !std::isspace((unsigned char)p[ne])) ++ne;
        RecRef r;
        r.name_off = x + 1; r.name_len = (u32)(ne - (x + 1));
        if(r.name_len > 2 && p[ne - 2] == '/' && std::isdigit((unsigned char)p[ne - 1])) r.name_len -= 2;   // trim_readno
        const size_t e2 = nl.next(e1 + 1);
        if(e2 - (e1 + 1) > 0x7fffffffull) return false;
        r.seq_off = e1 + 1; r.seq_len = (u32)(e2 - (e1 + 1));
        if(r.seq_len == 0 || std::memchr(p + x, '\r', e1 - x)) return false;
        if(seq_has_stop_byte(p + r.seq_off, r.seq_len)) return false;     // kseq would end the sequence there (or keep a CR)
        r.qual_off = ~0ull;
        size_t next = e2 + 1;
        if(fastq) {
            if(e2 + 1 >= n || p[e2 + 1] != '+') return false;
            const size_t e3 = nl.next(e2 + 1);
            if(e3 >= n) return false;
            const size_t e4 = nl.next(e3 + 1);
            if(e4 - (e3 + 1) != r.seq_len) return false;
            r.qual_off = e3 + 1;
            next = e4 + 1;
        }
        out.push_back(r);
        x = next;
    }
    return x == hi || (x == n + 1 && hi == n);
}
#endif
inline bool index_range_best(const char *p, size_t lo, size_t hi, size_t n, bool fastq, std::vector<RecRef> &out) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !std::getenv("BNS_B200_INDEX_SCALAR");
    if(avx2) return index_range_avx2(p, lo, hi, n, fastq, out);
#endif
    return index_range(p, lo, hi, n, fastq, out);
}
// gzip input in the same simple form: one decompressor thread per file inflates the stream window after window (zlib is a
// serial format; this takes inflate off the parser's thread and lets two mate files inflate side by side). A window is cut at
// its last complete record, the tail is carried into the headroom in front of the next window, and the window is indexed by
// the -p threads exactly like a stretch of a mapped plain file. Window buffers are recycled through a pool and stay alive
// while a batch in flight points into them (PinnedBatch::keep).
struct GzWindows {
    static constexpr size_t HEAD = 1u << 20;                              // room for the previous window's tail
    struct Chunk { std::shared_ptr<char> buf; size_t n = 0; u64 off = 0; bool eof = false; };   // data at buf + HEAD; off: uncompressed offset
    struct Pool {
        std::mutex mu; std::vector<char *> idle;
        ~Pool() { for(char *p : idle) delete[] p; }
    };
    gzFile fp = nullptr;
    size_t wsz;
    std::shared_ptr<Pool> pool{std::make_shared<Pool>()};
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Chunk> q;
    bool done = false, quit = false, failed = false;
    std::thread th;
    // BGZF (bgzip / htslib): a gzip file of independent blocks of <= 64 KiB whose headers carry their size (extra subfield
    // "BC"), so the blocks of a window inflate in parallel on the -p threads, straight out of the file mapping. Anything
    // unexpected (another kind of member, a block that does not inflate to its size and checksum) ends the indexed stream
    // there and kseq / gzread take over.
    const unsigned char *mp = nullptr;
    size_t mn = 0;
    unsigned nthreads = 1;
    bool bgzf = false;
    static u32 le32(const unsigned char *p) { return p[0] | (u32)p[1] << 8 | (u32)p[2] << 16 | (u32)p[3] << 24; }
    static bool bgzf_block(const unsigned char *p, size_t avail, size_t *total, size_t *data_off, u32 *isize) {
        if(avail < 28 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || p[3] != 4) return false;     // FEXTRA and nothing else
        const size_t xlen = p[10] | (size_t)p[11] << 8;
        if(12 + xlen + 8 > avail) return false;
        size_t bs = 0;
        for(size_t x = 12; x + 4 <= 12 + xlen;) {
            const size_t slen = p[x + 2] | (size_t)p[x + 3] << 8;
            if(p[x] == 'B' && p[x + 1] == 'C' && slen == 2 && x + 6 <= 12 + xlen) bs = (p[x + 4] | (size_t)p[x + 5] << 8) + 1;
            x += 4 + slen;
        }
        if(bs < 12 + xlen + 8 || bs > avail) return false;
        *total = bs; *data_off = 12 + xlen; *isize = le32(p + bs - 4);
        return *isize <= 65536;
    }
    GzWindows(const char *path, size_t window, const char *mapped = nullptr, size_t mapped_n = 0, unsigned nt = 1)
        : wsz(window), mp((const unsigned char *)mapped), mn(mapped_n), nthreads(std::max(1u, nt)) {
        fp = gzopen(path, "rb");
        if(!fp) return;
        gzbuffer(fp, 1 << 20);
        size_t t, d; u32 isz;
        bgzf = mp && wsz >= (1u << 17) && bgzf_block(mp, mn, &t, &d, &isz);
        th = std::thread([this] {
            try { if(bgzf) run_bgzf(); else run(); }
            catch(...) {                                                   // out of memory for a window: end the stream here, the
                { std::lock_guard<std::mutex> lk(mu); done = true; failed = true; }   // consumer hands the rest to kseq
                cv.notify_all();
            }
        });
    }
    GzWindows(const GzWindows &) = delete;
    ~GzWindows() { stop(); if(fp) gzclose(fp); }
    void stop() {
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv.notify_all();
        if(th.joinable()) th.join();
    }
    std::shared_ptr<char> acquire() {
        char *p = nullptr;
        { std::lock_guard<std::mutex> lk(pool->mu); if(!pool->idle.empty()) { p = pool->idle.back(); pool->idle.pop_back(); } }
        if(!p) p = new char[HEAD + wsz];
        std::shared_ptr<Pool> pl = pool;
        return std::shared_ptr<char>(p, [pl](char *x) { std::lock_guard<std::mutex> lk(pl->mu); pl->idle.push_back(x); });
    }
    void run() {
        u64 off = 0;
        for(;;) {
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return q.size() < 2 || quit; }); if(quit) return; }
            Chunk c;
            c.buf = acquire(); c.off = off;
            while(c.n < wsz) {                                            // a short or failed read ends the stream, as in KSeq::getc_
                const int r = gzread(fp, c.buf.get() + HEAD + c.n, (unsigned)std::min<size_t>(wsz - c.n, 1u << 30));
                if(r <= 0) { c.eof = true; break; }
                c.n += (size_t)r;
            }
            off += c.n;
            const bool last = c.eof;
            { std::lock_guard<std::mutex> lk(mu); q.push_back(std::move(c)); if(last) done = true; }
            cv.notify_all();
            if(last) return;
        }
    }
    void run_bgzf() {
        struct Blk { size_t in, in_len, out; u32 isize, crc; };
        std::vector<Blk> blks;
        size_t pos = 0;
        u64 off = 0;
        for(;;) {
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return q.size() < 2 || quit; }); if(quit) return; }
            Chunk c;
            c.buf = acquire(); c.off = off;
            blks.clear();
            size_t out = 0;
            bool bad = false;
            while(pos < mn) {                                             // the blocks of this window
                size_t total, doff; u32 isz;
                if(!bgzf_block(mp + pos, mn - pos, &total, &doff, &isz)) { bad = true; break; }
                if(out + isz > wsz) break;
                blks.push_back(Blk{pos + doff, total - doff - 8, out, isz, le32(mp + pos + total - 8)});
                out += isz; pos += total;
            }
            const size_t nb = blks.size();
            const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(nthreads, nb / 16 + 1));
            std::vector<size_t> first_bad(T, nb);
            char *dst = c.buf.get() + HEAD;
            auto work = [&](unsigned t) {
                const size_t lo = nb * t / T, hi = nb * (t + 1) / T;
                z_stream zs;
                std::memset(&zs, 0, sizeof zs);
                if(inflateInit2(&zs, -15) != Z_OK) { first_bad[t] = lo; return; }
                for(size_t i = lo; i < hi; ++i) {
                    const Blk &b = blks[i];
                    if(b.isize == 0) continue;                            // the empty end-of-file block
                    inflateReset(&zs);
                    zs.next_in = (Bytef *)(mp + b.in); zs.avail_in = (uInt)b.in_len;
                    zs.next_out = (Bytef *)(dst + b.out); zs.avail_out = b.isize;
                    const int r = inflate(&zs, Z_FINISH);
                    if(r != Z_STREAM_END || zs.avail_out != 0 || (u32)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)(dst + b.out), b.isize) != b.crc) {
                        first_bad[t] = i; break;
                    }
                }
                inflateEnd(&zs);
            };
            if(T == 1) work(0);
            else {
                std::vector<std::thread> pool;
                for(unsigned t = 0; t < T; ++t) pool.emplace_back(work, t);
                for(auto &w : pool) w.join();
            }
            const size_t fb = *std::min_element(first_bad.begin(), first_bad.end());
            if(fb < nb) { bad = true; out = blks[fb].out; }
            c.n = out; off += out;
            c.eof = !bad && pos >= mn;
            const bool last = c.eof || bad;
            {
                std::lock_guard<std::mutex> lk(mu);
                if(c.n || c.eof) q.push_back(std::move(c));
                if(bad) failed = true;
                if(last) done = true;
            }
            cv.notify_all();
            if(last) return;
        }
    }
    bool next(Chunk &out) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !q.empty() || done || quit; });
        if(q.empty()) return false;
        out = std::move(q.front()); q.pop_front();
        lk.unlock();
        cv.notify_all();
        return true;
    }
};
// the start of the first record that is not complete inside p[0, n) (n itself when the text ends on a record end); 0 when no
// record boundary is found near the end
inline size_t last_complete_record_end(const char *p, size_t n, bool fastq) {
    for(const size_t back : {(size_t)1 << 16, GzWindows::HEAD}) {
        size_t s = next_record_start(p, n > back ? n - back : 0, n, fastq);
        if(s >= n) continue;
        for(;;) {
            const size_t e1 = line_end(p, s, n); if(e1 >= n) break;
            size_t e = line_end(p, e1 + 1, n); if(e >= n) break;
            if(fastq) {
                e = line_end(p, e + 1, n); if(e >= n) break;
                e = line_end(p, e + 1, n); if(e >= n) break;
            }
            s = e + 1;
            if(s >= n) break;
        }
        return s;
    }
    return 0;
}
struct SimpleFile {
    MappedFile map;
    bool fastq = false, ok = false;
    size_t cursor = 0;                     // next unindexed byte of the (uncompressed) stream, a record start
    RawVec<RecRef> recs;                   // the current window; offsets are relative to text()
    size_t next_rec = 0;
    unsigned nthreads;
    size_t window;
    std::unique_ptr<WorkPool> pool;        // the -p threads
    std::vector<std::vector<RecRef>> part; // per-task pieces of a window's index, reused from window to window
    // gzip input
    std::unique_ptr<GzWindows> gz;
    std::shared_ptr<char> hold;            // the current window's buffer
    const char *gz_text = nullptr;
    u64 gz_base = 0;                       // uncompressed offset of gz_text[0]
    std::string carry;
    bool gz_first = true;
    bool at_end = false;                   // the whole stream has been indexed
    SimpleFile(const char *path, unsigned nt) : map(path), nthreads(std::max(1u, nt)), pool(new WorkPool(std::max(1u, nt))) {
        const char *e = std::getenv("BNS_B200_FASTQ_WINDOW");              // bytes per indexing window (tests use a small one)
        // 256 MB of text: a couple of batches per window, the first batch after a quarter of the time of a 1 GB window, and the
        // per-task pieces (reused) stay a few MB each
        window = e && std::atoll(e) > 0 ? (size_t)std::atoll(e) : ((size_t)256 << 20);
        if(map.p && (unsigned char)map.p[0] == 0x1f && (unsigned char)map.p[1] == 0x8b) {
            const char *g = std::getenv("BNS_B200_GZ_WINDOW");             // inflated bytes per window
            gz.reset(new GzWindows(path, g && std::atoll(g) > 0 ? (size_t)std::atoll(g) : ((size_t)64 << 20), map.p, map.n, nthreads));
            ok = gz->fp != nullptr;        // whether the text is in the simple form shows at the first window
            return;
        }
        if(!map.simple_candidate()) return;
        fastq = map.p[0] == '@';
        ok = true;
    }
    const char *text() const { return gz ? gz_text : map.p; }
    void stop() { if(gz) gz->stop(); }
    // where kseq must continue (an offset into the uncompressed stream): the first record not handed out yet
    size_t resume_offset() const {
        return next_rec < recs.size() ? (size_t)(gz ? gz_base : 0) + (size_t)recs[next_rec].name_off - 1 : cursor;
    }
    // every record of the file has been handed out and the file kept the simple form to its end (looks one window ahead)
    bool drained() {
        if(next_rec < recs.size() || !ok) return false;
        if(!at_end) refill();
        return ok && at_end && recs.empty();
    }
    // records of p[lo, hi) on the -p threads; n: end of the text
    bool index_parallel(const char *p, size_t lo, size_t hi, size_t n) {
        const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(nthreads, (hi - lo) / (1u << 20) + 1));
        std::vector<size_t> cut(T + 1);
        cut[0] = lo; cut[T] = hi;
        for(unsigned t = 1; t < T; ++t) cut[t] = std::min(hi, next_record_start(p, lo + (hi - lo) / T * t, n, fastq));
        for(unsigned t = 1; t <= T; ++t) if(cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
        if(part.size() < T) part.resize(T);
        std::vector<char> good(T, 1);
        pool->run(T, [&](unsigned t) {
            part[t].clear();
            part[t].reserve((cut[t + 1] - cut[t]) / 96 + 64);              // ~ one record per 100-300 bytes of text: few regrowths
            good[t] = index_range_best(p, cut[t], cut[t + 1], n, fastq, part[t]);
        });
        size_t total = 0;
        std::vector<size_t> at(T + 1, 0);
        for(unsigned t = 0; t < T; ++t) {
            if(!good[t]) { recs.clear(); return false; }
            at[t + 1] = (total += part[t].size());
        }
        // the parts become one array without a serial pass over it
        recs.resize(total);
        pool->run(T, [&](unsigned t) { if(!part[t].empty()) std::memcpy(recs.data() + at[t], part[t].data(), part[t].size() * sizeof(RecRef)); });
        return true;
    }
    // index the next window; false when the file is exhausted or leaves the simple form (then ok is false and `cursor` is where
    // kseq must take over)
    bool refill() {
        recs.clear(); next_rec = 0;
        if(gz) return refill_gz();
        if(!ok) return false;
        if(cursor >= map.n) { at_end = true; return false; }
        const char *p = map.p;
        const size_t n = map.n, lo = cursor;
        size_t hi = n;
        if(n - lo > window) { hi = next_record_start(p, lo + window, n, fastq); }
        if(!index_parallel(p, lo, hi, n)) { ok = false; return false; }   // leave `cursor` at the window start for kseq
        cursor = hi;
        return !recs.empty();
    }
    bool refill_gz() {
        if(!ok) return false;
        GzWindows::Chunk ch;
        if(!gz->next(ch)) {                                                // the stream has ended
            if(gz->failed) { ok = false; carry.clear(); return false; }    // (not cleanly: kseq continues at `cursor`)
            at_end = true; return false;
        }
        const size_t cl = carry.size();
        char *t = ch.buf.get() + GzWindows::HEAD - cl;
        if(cl) std::memcpy(t, carry.data(), cl);
        carry.clear();
        const size_t n = cl + ch.n;
        hold = ch.buf; gz_text = t;
        gz_base = ch.off - cl; cursor = (size_t)gz_base;
        if(n == 0) { if(ch.eof) at_end = true; else ok = false; return false; }
        if(gz_first) {
            gz_first = false;
            if(t[0] != '@' && t[0] != '>') { ok = false; return false; }
            fastq = t[0] == '@';
        }
        size_t hi = n;
        if(!ch.eof) {
            hi = last_complete_record_end(t, n, fastq);
            if(hi == 0 || n - hi > GzWindows::HEAD) { ok = false; return false; }   // no boundary / a record beyond the headroom
            carry.assign(t + hi, n - hi);
        }
        if(!index_parallel(t, 0, hi, hi)) { ok = false; carry.clear(); return false; }
        cursor = (size_t)gz_base + hi;
        return !recs.empty();
    }
};
// the next batch out of the index: records until >= chunk_size bases and an even count (bseq_read's rule)
// (f2: the mates' file; records are interleaved r1, r2, r1, r2 ... like bseq_read does, mates of file 2 at the odd indices)
inline bool fill_pinned(int chunk_size, PinnedBatch &b, SimpleFile &f, SimpleFile *f2 = nullptr) {
    b.clear();
    b.refs.clear();
    auto have = [](SimpleFile &x) { return x.next_rec < x.recs.size() || x.refill(); };
    if(!have(f) || (f2 && !have(*f2))) return false;                  // either file is out of indexed records: the caller hands over to kseq
    // the batch's records: a stretch [s1, e1) of file 1's window (and [s2, e2) of the mates' file), found by one read-only pass
    // over the sequence lengths; a batch points into ONE window per file, so it also ends where a window does
    const size_t s1 = f.next_rec, s2 = f2 ? f2->next_rec : 0;
    const size_t avail = f2 ? std::min(f.recs.size() - s1, f2->recs.size() - s2) : f.recs.size() - s1;
    u64 size = 0;
    size_t take = 0;
    while(take < avail) {
        size += f.recs[s1 + take].seq_len;
        if(f2) size += f2->recs[s2 + take].seq_len;
        ++take;
        const size_t nrec = f2 ? 2 * take : take;
        if((long)size >= chunk_size && (nrec & 1) == 0) break;
    }
    f.next_rec += take;
    if(f2) f2->next_rec += take;
    const size_t n = f2 ? 2 * take : take;
    if(!n) return false;
    b.map = f.text(); b.keep = f.hold;
    b.map2 = f2 ? f2->text() : nullptr;
    if(f2) b.keep2 = f2->hold;
    PinnedBatch::grow(b.bases, b.cap_bases, 0, size + 16);
    PinnedBatch::grow(b.offs, b.cap_offs, 0, n + 2);
    b.refs.resize(n);
    // two parallel passes over the stretch: bases per slice, then (with every slice's first offset known) the records' offsets,
    // their references and the sequence bytes themselves
    const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(f.nthreads, n / 8192 + 1));
    std::vector<u64> base(T + 1, 0);
    auto rec_at = [&](size_t i) -> const RecRef & { return f2 ? ((i & 1) ? f2->recs[s2 + (i >> 1)] : f.recs[s1 + (i >> 1)]) : f.recs[s1 + i]; };
    auto run = [&](const std::function<void(unsigned)> &fn) { f.pool->run(T, fn); };
    run([&](unsigned t) { u64 sum = 0; for(size_t i = n * t / T, hi = n * (t + 1) / T; i < hi; ++i) sum += rec_at(i).seq_len; base[t + 1] = sum; });
    for(unsigned t = 0; t < T; ++t) base[t + 1] += base[t];
    run([&](unsigned t) {
        u64 off = base[t];
        for(size_t i = n * t / T, hi = n * (t + 1) / T; i < hi; ++i) {
            const RecRef &r = rec_at(i);
            b.refs[i] = r;
            b.offs[i] = off;
            std::memcpy(b.bases + off, b.map_of(i) + r.seq_off, r.seq_len);
            off += r.seq_len;
        }
    });
    b.offs[n] = size;
    b.n = n; b.n_bases = size;
    return true;
}

// bseq_read (kseq_declare.h:112-145) into a pinned batch: records until >= chunk_size bases and an even count
inline bool read_pinned(int chunk_size, PinnedBatch &b, KSeq *ks, KSeq *ks2) {
    b.clear();
    while(ks->read() >= 0) {
        if(ks2 && ks2->read() < 0) { std::fprintf(stderr, "[W::%s] the 2nd file has fewer sequences.\n", __func__); break; }
        b.push(ks);
        if(ks2) b.push(ks2);
        if((long)b.n_bases >= chunk_size && (b.n & 1) == 0) break;
    }
    if(b.n_bases == 0 && ks2 && ks2->read() >= 0) std::fprintf(stderr, "[W::%s] the 1st file has fewer sequences.\n", __func__);
    return b.n != 0;
}
}  // namespace detail

// classify_seqs, classifier.h:269-289: classify `chunk_size` reads (mates interleaved when is_paired) and append each
// record's text to cks in read order. per_set / the thread pool of the reference have no role here: the batch is one
// GPU call (a caller that loops over single reads pays one device round trip per call: batch them).
template <typename ScoreType>
void classify_seqs(ClassifierGeneric<ScoreType> &c, const TaxMap *taxmap, bseq1_t *bs, std::string &cks,
                   const unsigned chunk_size, const unsigned /*per_set*/, const int is_paired) {
    if(!c.tax_loaded_) c.load_taxonomy(taxmap);
    const unsigned inc = is_paired ? 2 : 1, n = (chunk_size / inc) * inc;
    if(!n) return;
    // the batch goes to the device from a pinned staging buffer kept per thread (no pageable copy inside the library)
    thread_local detail::PinnedBatch stage;
    size_t total = 0;
    for(unsigned i = 0; i < n; ++i) total += bs[i].seq.size();
    stage.clear();
    detail::PinnedBatch::grow(stage.bases, stage.cap_bases, 0, total + 16);
    detail::PinnedBatch::grow(stage.offs, stage.cap_offs, 0, (size_t)n + 2);
    stage.offs[0] = 0;
    for(unsigned i = 0; i < n; ++i) {
        std::memcpy(stage.bases + stage.offs[i], bs[i].seq.data(), bs[i].seq.size());
        stage.offs[i + 1] = stage.offs[i] + bs[i].seq.size();
    }
    detail::classify_views(c, stage.bases, stage.offs, [&](size_t i) {
        return detail::ReadView{bs[i].name.c_str(), stage.bases + stage.offs[i], bs[i].qual.empty() ? nullptr : bs[i].qual.c_str(), bs[i].l_seq};
    }, n, is_paired, cks);
}

}  // namespace bns
// ---- the reference's spelling of the same calls -------------------------------------------------------------------------
// classify_main (bin/bonsai.cpp:107-163) is written against khash_t(c) / khash_t(p), ks::string and ForPool. These shims let
// such a caller compile against this header unchanged: khash_t(c) is the raw table view above, khash_t(p) the parent map,
// ks::string a growable byte buffer with the members classify_seqs / process_dataset use, ForPool a thread count (the GPU
// call needs no pool). They are only defined when the real headers are not part of the translation unit.
#ifndef khash_t
#define khash_t(name) ::bns::kh_##name##_t
#define BNS_B200_KHASH_SHIM 1
#endif
#ifndef kh_destroy
#define kh_destroy(name, h) ::bns::kh_destroy_##name(h)
#endif
namespace ks {
#ifndef BNS_B200_NO_KS_SHIM
class string {                                    // kspp/ks.h: the subset on the classify path
    std::string s_;
public:
    explicit string(size_t reserve = 0) { s_.reserve(reserve); }
    const char *data() const { return s_.data(); }
    size_t size() const { return s_.size(); }
    void clear() { s_.clear(); }
    void resize(size_t n) { s_.reserve(n); }      // classify_seqs sizes the buffer before filling it (classifier.h:277)
    void terminate() {}
    int putsn_(const char *p, size_t n) { s_.append(p, n); return (int)n; }
    int write(int fd) const {
        size_t put = 0;
        while(put < s_.size()) { const ssize_t r = ::write(fd, s_.data() + put, s_.size() - put); if(r <= 0) return -1; put += (size_t)r; }
        return (int)put;
    }
    std::string &str() { return s_; }
};
#endif
}  // namespace ks
namespace bns {
using kh_p_t = TaxMap;
inline void kh_destroy_p(kh_p_t *t) { delete t; }
inline void kh_destroy_c(kh_c_t *) {}              // a Database owns its arrays
struct ForPool { int nt_; explicit ForPool(int nthreads = 1) : nt_(nthreads) {} };   // util.h ForPool: kt_forpool handle
// classify_seqs with the reference's own parameter list (classifier.h:269); the pool is accepted and not needed
template <typename ScoreType>
void classify_seqs(ClassifierGeneric<ScoreType> &c, const kh_p_t *taxmap, bseq1_t *bs, ks::string &cks,
                   const unsigned chunk_size, const unsigned per_set, const int is_paired, ForPool &) {
    classify_seqs(c, taxmap, bs, cks.str(), chunk_size, per_set, is_paired);
}

// process_dataset, classifier.h:296-337, as a pipeline: a reader thread parses FASTA/FASTQ(.gz) records into a ring of
// pinned host batches; one worker thread per GPU (ClassifierGeneric::set_gpus) takes the batches dealt to it round-robin
// -- batch s goes to GPU s mod n, SURVEY 8e -- classifies it on its own context (the library DMAs straight from the pinned
// buffer on that context's streams) and formats its text; a writer thread emits the texts in batch order.
template <typename ScoreType>
void process_dataset(ClassifierGeneric<ScoreType> &c, const TaxMap *taxmap, const char *fq1, const char *fq2, std::FILE *out,
                     unsigned chunk_size, unsigned /*per_set*/) {
    const auto t_enter = std::chrono::steady_clock::now();
    auto since_enter = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_enter).count(); };
    double t_first_batch = 0, t_reader_reserve = 0, t_reader_fill = 0, t_reader_done = 0;   // BNS_B200_VERBOSE: where the reader's time goes
    if(!c.tax_loaded_) c.load_taxonomy(taxmap);
    detail::KSeq ks1(fq1);
    std::unique_ptr<detail::KSeq> ks2(fq2 ? new detail::KSeq(fq2) : nullptr);
    // plain or gzip files in the simple 4-line / 2-line form are indexed by the -p threads (detail::SimpleFile; gzip streams
    // are inflated by one thread per file, detail::GzWindows); with mates, both files must qualify
    std::unique_ptr<detail::SimpleFile> simple(new detail::SimpleFile(fq1, c.nt_)), simple2(fq2 ? new detail::SimpleFile(fq2, c.nt_) : nullptr);
    const char *ingest = std::getenv("BNS_B200_INGEST");                  // "kseq": the single-threaded reader for everything
    if(!simple->ok || (simple2 && !simple2->ok) || (ingest && !std::strcmp(ingest, "kseq"))) { simple.reset(); simple2.reset(); }
    bool use_index = simple != nullptr;          // the mappings stay alive to the end: batches in flight point into them
    const int fn = fileno(out), is_paired = fq2 != nullptr;
    const int G = c.replicas_.empty() ? 1 : c.n_gpus();
    // two host workers per GPU: batch s goes to worker s mod 2G, i.e. to GPU s mod G; while one worker formats its batch the
    // other has the device
    const int NW = 2 * G;
    // one slot per worker and one the reader fills meanwhile: pinning memory costs ~0.6 ms per MB and holds a driver lock the
    // workers' calls wait for, so the ring is as small as the pipeline allows (five slots of 100 MB were 0.28 s on 8 M reads)
    const int NB = NW + 1;
    std::vector<detail::PinnedBatch> ring((size_t)NB);
    // pinned allocations are slow (tens of ms each): a ring slot gets its buffers when the reader first fills it, while the
    // earlier batches are already on the device
    for(auto &b : ring) b.keep_qual = c.get_emit_fastq() != 0;
    std::vector<int> state((size_t)NB, 0);       // 0 free, 1 filled
    std::vector<u64> seq_of((size_t)NB, ~0ull);
    u64 end_seq = ~0ull;                         // batches [0, end_seq) exist
    std::mutex mu;
    std::condition_variable cv;
    std::string reader_error;
    // Pinned allocations take tens of milliseconds each (five slots of 84 MB were 0.2 of the reader's 0.5 s on 8 M reads): a
    // helper thread makes them, slot after slot, while the reader indexes the first window and fills the slots that are ready.
    std::vector<int> slot_ready((size_t)NB, 0);
    std::atomic<bool> alloc_stop(false);
    std::string alloc_error;
    std::thread allocator([&]() {
        for(size_t i = 0; i < (size_t)NB && !alloc_stop.load(); ++i) {
            try { ring[i].reserve(chunk_size); } catch(const std::exception &e) { std::lock_guard<std::mutex> lk(mu); alloc_error = e.what(); }
            { std::lock_guard<std::mutex> lk(mu); slot_ready[i] = 1; }
            cv.notify_all();
        }
        { std::lock_guard<std::mutex> lk(mu); for(auto &r : slot_ready) r = 1; }   // stopped early: the slots left are not needed
        cv.notify_all();
    });
    std::thread reader([&]() {
        struct StopAlloc { std::atomic<bool> &f; ~StopAlloc() { f = true; } } stop_alloc{alloc_stop};   // no more slots are needed once the input has ended
        try {
            for(u64 sq = 0;; ++sq) {
                const size_t i = (size_t)(sq % (u64)NB);
                { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return state[i] == 0; }); }
                bool got = false;
                const double tr0 = since_enter();
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return slot_ready[i] != 0; });
                    if(!alloc_error.empty()) BNS_RUNTIME_ERROR(alloc_error);
                }
                const double tr1 = since_enter();
                t_reader_reserve += tr1 - tr0;
                if(use_index) {
                    got = detail::fill_pinned((int)chunk_size, ring[i], *simple, simple2.get());
                    t_reader_fill += since_enter() - tr1;
                    if(sq == 0) t_first_batch = since_enter();
                    if(!got && simple->drained() && (!simple2 || simple2->drained())) {
                        // the index served the input to its end (a gzip stream would inflate once more to seek there)
                        use_index = false;
                        t_reader_done = since_enter();
                        { std::lock_guard<std::mutex> lk(mu); end_seq = sq; }
                        cv.notify_all();
                        return;
                    }
                    if(!got) {
                        // out of indexed records: a file left the simple form, or one mate file ended before the other (kseq
                        // then reports it the way bseq_read does). kseq takes over at the first record the index did not hand out.
                        gzseek(ks1.fp, (z_off_t)simple->resume_offset(), SEEK_SET);
                        if(simple2) gzseek(ks2->fp, (z_off_t)simple2->resume_offset(), SEEK_SET);
                        simple->stop();
                        if(simple2) simple2->stop();
                        use_index = false;
                    }
                }
                if(!got && !use_index) got = detail::read_pinned((int)chunk_size, ring[i], &ks1, ks2.get());
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if(got) { state[i] = 1; seq_of[i] = sq; } else end_seq = sq;
                }
                cv.notify_all();
                if(!got) return;
            }
        } catch(const std::exception &e) {
            std::lock_guard<std::mutex> lk(mu);
            reader_error = e.what();
            if(end_seq == ~0ull) { u64 mx = 0; for(size_t i = 0; i < (size_t)NB; ++i) if(state[i] == 1) mx = std::max(mx, seq_of[i] + 1); end_seq = mx; }
            cv.notify_all();
        }
    });
    // text leaves through a writer thread, in batch order, so write(2) overlaps the next batches
    std::fflush(out);
    std::mutex wmu;
    std::condition_variable wcv;
    std::map<u64, std::vector<std::string>> wq;  // finished texts (the formatting threads' slices, in order) by batch number (at most NB: a worker holds its ring slot until it has queued its text)
    u64 w_end = ~0ull;                           // set once the last batch number is known
    std::string werr;
    std::thread writer([&]() {
        for(u64 next = 0;; ++next) {
            std::vector<std::string> ts;
            {
                std::unique_lock<std::mutex> lk(wmu);
                wcv.wait(lk, [&] { return wq.count(next) || next >= w_end; });
                if(!wq.count(next)) return;
                ts = std::move(wq[next]); wq.erase(next);
            }
            for(const std::string &t : ts) {
                if(!werr.empty()) break;                                   // after a failed write: drain and drop
                size_t put = 0;
                while(put < t.size()) {
                    const ssize_t r = ::write(fn, t.data() + put, t.size() - put);
                    if(r <= 0) { std::lock_guard<std::mutex> lk(wmu); werr = "write failed"; break; }
                    put += (size_t)r;
                }
            }
        }
    });
    std::atomic<bool> first(true);
    std::string failure;
    const bool verbose = std::getenv("BNS_B200_VERBOSE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    std::vector<double> t_wait((size_t)NW, 0.), t_classify((size_t)NW, 0.);
    std::vector<size_t> n_batches((size_t)NW, 0);
    std::vector<std::mutex> device_mu((size_t)G);
    const unsigned fmt_threads = std::max(1u, (unsigned)c.nt_ / (unsigned)G);
    auto work = [&](int g) {                                               // worker g of NW, on GPU g mod G
        bns_b200_t *h = c.gpu(g % G);
        std::mutex *dmu = &device_mu[(size_t)(g % G)];
        for(u64 sq = (u64)g;; sq += (u64)NW) {
            const size_t i = (size_t)(sq % (u64)NB);
            const double tw = now();
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return (state[i] == 1 && seq_of[i] == sq) || sq >= end_seq; });
                if(!(state[i] == 1 && seq_of[i] == sq)) return;
            }
            t_wait[(size_t)g] += now() - tw;
            detail::PinnedBatch &b = ring[i];
            ++n_batches[(size_t)g];
            std::string text;
            std::vector<std::string> slices;
            bool failed;
            { std::lock_guard<std::mutex> lk(mu); failed = !failure.empty(); }
            try {
                if(!failed) {
                    const double tc = now();
                    if(b.map)                                          // indexed ingest: names / qualities stay in the file mapping
                        detail::classify_views(c, b.bases, b.offs, [&b](size_t r) {
                            const detail::RecRef &ref = b.refs[r];
                            const char *mp = b.map_of(r);
                            return detail::ReadView{mp + ref.name_off, b.bases + b.offs[r], b.keep_qual && ref.qual_off != ~0ull ? mp + ref.qual_off : nullptr,
                                                    (int)ref.seq_len, (int)ref.name_len};
                        }, (unsigned)b.n, is_paired, text, h, fmt_threads, dmu, &slices);
                    else
                        detail::classify_views(c, b.bases, b.offs, [&b](size_t r) {
                            return detail::ReadView{b.names[r].c_str(), b.bases + b.offs[r], b.has_qual[r] ? b.quals[r].c_str() : nullptr,
                                                    (int)(b.offs[r + 1] - b.offs[r])};
                        }, (unsigned)b.n, is_paired, text, h, fmt_threads, dmu, &slices);
                    t_classify[(size_t)g] += now() - tc;
                    if(sq == 0) { std::fprintf(stderr, "nseq: %i\n", (int)b.n); first = false; }     // classifier.h:312
                }
            } catch(const std::exception &e) { std::lock_guard<std::mutex> lk(mu); if(failure.empty()) failure = e.what(); }
            if(!text.empty()) slices.push_back(std::move(text));
            { std::lock_guard<std::mutex> lk(wmu); wq[sq] = std::move(slices); }
            wcv.notify_all();
            { std::lock_guard<std::mutex> lk(mu); state[i] = 0; }
            cv.notify_all();
        }
    };
    std::vector<std::thread> workers;
    for(int g = 1; g < NW; ++g) workers.emplace_back(work, g);
    work(0);                                                               // the first worker of GPU 0 on the caller's thread
    for(auto &t : workers) t.join();
    reader.join();
    alloc_stop = true;
    allocator.join();
    const double t_workers_done = since_enter();
    { std::lock_guard<std::mutex> lk(wmu); w_end = end_seq; }
    wcv.notify_all();
    writer.join();
    const double t_writer_done = since_enter();
    if(!reader_error.empty()) BNS_RUNTIME_ERROR(reader_error);
    if(!failure.empty()) BNS_RUNTIME_ERROR(failure);
    if(!werr.empty()) BNS_RUNTIME_ERROR(werr);
    if(first) std::fprintf(stderr, "Could not get any sequences from file, fyi.\n");
    // what the function holds is released here, step by step, so that BNS_B200_VERBOSE can say what each step costs
    const double t_rel0 = since_enter();
    if(c.exit_follows_) {
        static std::vector<std::vector<detail::PinnedBatch>> *kept = new std::vector<std::vector<detail::PinnedBatch>>();   // never destroyed
        kept->push_back(std::move(ring));
    }
    ring.clear();
    const double t_rel1 = since_enter();
    simple.reset(); simple2.reset();
    const double t_rel2 = since_enter();
    if(verbose)
        std::fprintf(stderr, "[process_dataset] releasing the pinned ring %.3f s, the file mappings and indices %.3f s\n", t_rel1 - t_rel0, t_rel2 - t_rel1);
    if(verbose)
        std::fprintf(stderr, "[process_dataset] first batch ready at %.3f s, reader done at %.3f s (waiting for pinned buffers %.3f s, index + fill %.3f s), workers done at %.3f s, "
                     "writer done at %.3f s\n", t_first_batch, t_reader_done, t_reader_reserve, t_reader_fill, t_workers_done, t_writer_done);
    if(verbose)
        std::fprintf(stderr, "[process_dataset] all workers: device calls (incl. waiting for the GPU's other worker) %.3f s, formatting %.3f s, joining the slices %.3f s\n",
                     c.t_device_ns_.load() * 1e-9, c.t_format_ns_.load() * 1e-9, c.t_join_ns_.load() * 1e-9);
    if(verbose)
        for(int g = 0; g < NW; ++g)
            std::fprintf(stderr, "[process_dataset] gpu %d worker %d: %zu batches, waiting for the reader %.2f s, classify + format %.2f s\n",
                         g % G, g / G, n_batches[(size_t)g], t_wait[(size_t)g], t_classify[(size_t)g]);
}

}  // namespace bns
