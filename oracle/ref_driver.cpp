// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Thin extern "C" driver over the UNMODIFIED reference headers under /root/reference
// (dnbaker/bonsai @ 6741de9c). It is compiled where those sources lie by oracle/Makefile
// into oracle/_ref/libbns_ref_{v3,v4}.so (git-ignored, travels to the GPU box) and is used
//   * to pin the plain-C restatement (oracle/bns_oracle.c) against the real reference, and
//   * as the `"kind": "reference"` CPU baseline of bench.py.
// No reference source is copied here: every function body below only *calls* the reference:
//   Encoder::for_each            include/bonsai/encoder.h:416
//   Encoder::for_each_canon/uncanon (path overloads)  encoder.h:448-530
//   kh_get / kh_put              include/bonsai/khash64.h:250,327
//   linear::counter::add         linear/linear.h:229
//   resolve_tree / lca           include/bonsai/util.h:831,634
//   build_parent_map             include/bonsai/util.h:766
//   update_lca_map               include/bonsai/feature_min.h:205
//   classify_seq                 include/bonsai/classifier.h:213
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <omp.h>
#include "bonsai/encoder.h"
#include "bonsai/classifier.h"
#include "bonsai/feature_min.h"
#include "bonsai/util.h"

using namespace bns;

namespace {

spvec_t make_gaps(unsigned k, const uint16_t *gaps) {
    if(!gaps) return spvec_t{};
    return spvec_t(gaps, gaps + (k - 1));
}

// api: 0 = string overload for_each(fn, str, l)        (what classify_seq calls)
//      1 = path-overload semantics on one record: canon ? for_each_canon : for_each_uncanon
//          (encoder.h:448-464; what the DB builder uses; spaced seeds work here)
//      2 = call by call: assign(), then next_canonicalized_minimizer() / next_minimizer() while has_next_kmer()
//          (encoder.h:201-206,594-628); the values from the first full window on, nothing filtered
template<typename Score, typename F>
void run_encoder(Encoder<Score> &enc, int api, const char *seq, uint64_t len, const F &fn, int iter_canon = 0) {
    if(api == 0) {
        enc.for_each(fn, seq, len);
    } else if(api == 2) {
        enc.assign(seq, len);
        const uint64_t wsz = (uint64_t)enc.sp_.w_ - enc.sp_.c_ + 1;
        uint64_t calls = 0;
        while(enc.has_next_kmer()) {
            const u64 m = iter_canon ? enc.next_canonicalized_minimizer() : enc.next_minimizer();
            if(++calls >= wsz) fn(m);
        }
    } else {
        enc.assign(seq, len);
        if(!enc.has_next_kmer()) return;
        if(enc.canonicalize()) {
            if(enc.sp_.unwindowed()) enc.for_each_canon_unwindowed(fn);
            else                     enc.for_each_canon_windowed(fn);
        } else {
            if(enc.sp_.unspaced()) {
                if(enc.sp_.unwindowed()) enc.for_each_uncanon_unspaced_unwindowed(fn);
                else                     enc.for_each_uncanon_unspaced_windowed(fn);
            } else enc.for_each_uncanon_spaced(fn);
        }
    }
}

template<typename Score>
int64_t encode_impl(unsigned k, unsigned w, const uint16_t *gaps, int canon, int api,
                    const char *seq, uint64_t len, uint64_t *out, uint64_t cap) {
    Spacer sp(k, w, make_gaps(k, gaps));
    Encoder<Score> enc(sp, (bool)canon);
    uint64_t n = 0;
    run_encoder(enc, api, seq, len, [&](u64 km) {
        if(n < cap) out[n] = km;
        ++n;
    }, canon);
    return (int64_t)n;
}

struct RefTax { khash_t(p) *m; };
struct RefDb  { khash_t(c) *m; };

template<typename Score>
void classify_impl(const RefDb *db, const RefTax *tax, unsigned k, unsigned w, const uint16_t *gaps,
                   int canon, int api, const char *bases, const uint64_t *offsets, uint64_t n_reads,
                   int paired, uint32_t *taxon_out, uint32_t *nhit_out, uint32_t *nmiss_out,
                   uint32_t *taxa_out, const uint64_t *taxa_offsets, int nthreads) {
    Spacer sp(k, w, make_gaps(k, gaps));
    const Encoder<Score> proto(sp, (bool)canon);
    const int inc = paired ? 2 : 1;
    const int64_t nrec = (int64_t)(n_reads / inc);
    if(nthreads <= 0) nthreads = omp_get_max_threads();
    #pragma omp parallel num_threads(nthreads)
    {
        Encoder<Score> enc(proto);                    // one Encoder copy per thread, classifier.h:258
        #pragma omp for schedule(dynamic, 1024)
        for(int64_t r = 0; r < nrec; ++r) {
            const uint64_t i = (uint64_t)r * inc;
            tax_counter hit_counts;
            uint32_t missing = 0, nhit = 0;
            uint32_t *tp = taxa_out ? taxa_out + taxa_offsets[r] : nullptr;
            auto fn = [&](u64 kmer) {                 // the lambda of classifier.h:225-229
                khiter_t ki;
                if((ki = kh_get(c, db->m, kmer)) == kh_end(db->m)) ++missing;
                else {
                    if(tp) tp[nhit] = kh_val(db->m, ki);
                    ++nhit;
                    hit_counts.add(kh_val(db->m, ki));
                }
            };
            for(int m = 0; m < inc; ++m)
                run_encoder(enc, api, bases + offsets[i + m], offsets[i + m + 1] - offsets[i + m], fn);
            taxon_out[r] = resolve_tree(hit_counts, tax->m);
            if(nhit_out)  nhit_out[r] = nhit;
            if(nmiss_out) nmiss_out[r] = missing;
        }
    }
}

} // namespace

extern "C" {

int bref_cast_saturates(void) {
    volatile double d = -5.5;
    return (uint64_t)d == ~uint64_t(0);
}
const char *bref_build_info(void) {
    return "reference dnbaker/bonsai@6741de9c, g++ " __VERSION__
#ifdef __AVX512F__
           ", AVX512F"
#else
           ", no-AVX512"
#endif
           ;
}
uint64_t bref_lex_score(uint64_t x)           { return lex_score(x); }
uint64_t bref_wang64(uint64_t x)              { return __ac_Wang64_hash(x); }
uint64_t bref_rc(uint64_t x, unsigned k)      { return reverse_complement(x, (uint8_t)k); }
uint64_t bref_canonical(uint64_t x, unsigned k) { return canonical_representation(x, (uint8_t)k); }

// Spacer geometry (spacer.h:59-71): returns c_, w_, unspaced(), unwindowed()
void bref_spacer(unsigned k, unsigned w, const uint16_t *gaps, uint32_t *c, uint32_t *wout,
                 int *unspaced, int *unwindowed) {
    Spacer sp(k, w, make_gaps(k, gaps));
    *c = sp.c_; *wout = sp.w_; *unspaced = sp.unspaced(); *unwindowed = sp.unwindowed();
}
// parse_spacing (spacer.h:29): returns number of entries written
int bref_parse_spacing(const char *s, unsigned k, uint16_t *out, int cap) {
    spvec_t v = parse_spacing(s, k);
    for(size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}

int64_t bref_encode(unsigned k, unsigned w, const uint16_t *gaps, int score, int canon, int api,
                    const char *seq, uint64_t len, uint64_t *out, uint64_t cap) {
    return score == 1 ? encode_impl<score::Entropy>(k, w, gaps, canon, api, seq, len, out, cap)
                      : encode_impl<score::Lex>(k, w, gaps, canon, api, seq, len, out, cap);
}

// ---- taxonomy ---------------------------------------------------------------------------
void *bref_tax_load(const char *nodes_dmp) {
    RefTax *t = new RefTax{build_parent_map(nodes_dmp)};
    return t;
}
void *bref_tax_from_pairs(const uint32_t *child, const uint32_t *parent, uint64_t n) {
    khash_t(p) *m = kh_init(p);
    int khr;
    for(uint64_t i = 0; i < n; ++i) {
        khint_t ki = kh_put(p, m, child[i], &khr);
        kh_val(m, ki) = parent[i];
    }
    khint_t ki = kh_put(p, m, 1, &khr);          // util.h:780-781
    kh_val(m, ki) = 0;
    return new RefTax{m};
}
uint64_t bref_tax_size(void *t) { return kh_size(((RefTax *)t)->m); }
uint64_t bref_tax_pairs(void *t, uint32_t *child, uint32_t *parent, uint64_t cap) {
    khash_t(p) *m = ((RefTax *)t)->m;
    uint64_t n = 0;
    for(khiter_t ki = kh_begin(m); ki != kh_end(m); ++ki)
        if(kh_exist(m, ki)) {
            if(n < cap) child[n] = kh_key(m, ki), parent[n] = kh_val(m, ki);
            ++n;
        }
    return n;
}
void bref_tax_free(void *t) { kh_destroy(p, ((RefTax *)t)->m); delete (RefTax *)t; }
uint32_t bref_lca(void *t, uint32_t a, uint32_t b) { return lca(((RefTax *)t)->m, a, b); }
uint32_t bref_resolve(void *t, const uint32_t *taxa, const uint16_t *counts, uint32_t n) {
    tax_counter hc;
    for(uint32_t i = 0; i < n; ++i) hc.add(taxa[i], counts[i]);
    return resolve_tree(hc, ((RefTax *)t)->m);
}

// ---- database ----------------------------------------------------------------------------
// DB = k-mer sets of each genome (path overload of for_each, as fill_set_genome does,
// feature_min.h:68-83) merged with update_lca_map (feature_min.h:205).
void *bref_db_build(void *tax, unsigned k, unsigned w, const uint16_t *gaps, int score, int canon,
                    int n_genomes, const char **paths, const uint32_t *taxids) {
    Spacer sp(k, w, make_gaps(k, gaps));
    khash_t(c) *kc = kh_init(c);
    for(int g = 0; g < n_genomes; ++g) {
        khash_t(all) *set = kh_init(all);
        if(score == 1) fill_set_genome<score::Entropy>(paths[g], sp, set, 0, nullptr, (bool)canon);
        else           fill_set_genome<score::Lex>(paths[g], sp, set, 0, nullptr, (bool)canon);
        update_lca_map(kc, set, ((RefTax *)tax)->m, taxids[g]);
        kh_destroy(all, set);
    }
    return new RefDb{kc};
}
void *bref_db_from_pairs(const uint64_t *keys, const uint32_t *vals, uint64_t n) {
    khash_t(c) *kc = kh_init(c);
    int khr;
    for(uint64_t i = 0; i < n; ++i) {
        khint_t ki = kh_put(c, kc, keys[i], &khr);
        kh_val(kc, ki) = vals[i];
    }
    return new RefDb{kc};
}
// Borrowed views of the raw khash arrays (khash64.h:213-219)
void bref_db_arrays(void *db, const uint64_t **keys, const uint32_t **vals, const uint32_t **flags,
                    uint64_t *n_buckets, uint64_t *size) {
    khash_t(c) *m = ((RefDb *)db)->m;
    *keys = m->keys; *vals = m->vals; *flags = m->flags; *n_buckets = m->n_buckets; *size = m->size;
}
// returns 1 and writes *val on hit, 0 on miss; *probes = slots inspected (for the p-bar figure)
int bref_db_get(void *db, uint64_t key, uint32_t *val) {
    khash_t(c) *m = ((RefDb *)db)->m;
    khint_t ki = kh_get(c, m, key);
    if(ki == kh_end(m)) return 0;
    *val = kh_val(m, ki);
    return 1;
}
void bref_db_free(void *db) { kh_destroy(c, ((RefDb *)db)->m); delete (RefDb *)db; }

// ---- classify: the CPU-baseline core loop --------------------------
// bases/offsets: reads concatenated, offsets[n_reads+1]. paired: mates interleaved (i, i+1),
// one record per pair (classifier.h:233-236). taxa_out (optional): per-record ordered hit list,
// record r starts at taxa_offsets[r].
void bref_classify(void *db, void *tax, unsigned k, unsigned w, const uint16_t *gaps, int score,
                   int canon, int api, const char *bases, const uint64_t *offsets, uint64_t n_reads,
                   int paired, uint32_t *taxon_out, uint32_t *nhit_out, uint32_t *nmiss_out,
                   uint32_t *taxa_out, const uint64_t *taxa_offsets, int nthreads) {
    if(score == 1)
        classify_impl<score::Entropy>((RefDb *)db, (RefTax *)tax, k, w, gaps, canon, api, bases, offsets,
                                      n_reads, paired, taxon_out, nhit_out, nmiss_out, taxa_out, taxa_offsets, nthreads);
    else
        classify_impl<score::Lex>((RefDb *)db, (RefTax *)tax, k, w, gaps, canon, api, bases, offsets,
                                  n_reads, paired, taxon_out, nhit_out, nmiss_out, taxa_out, taxa_offsets, nthreads);
}

// ---- classify_seq, unmodified, for the text format (classifier.h:213) ----------------------
// Works around SURVEY App. B-3/B-4 from the *outside*: the ClassifierGeneric storage is zeroed
// before construction (output_flag_ is never initialised) and bs.sam is pre-seeded.
// Returns a malloc'd NUL-terminated buffer the caller frees with bref_free; *len_out = length.
char *bref_classify_text(void *db, void *tax, unsigned k, unsigned w, const uint16_t *gaps, int canon,
                         int emit_all, int emit_fastq, int emit_kraken,
                         const char *bases, const uint64_t *offsets, const char *const *names,
                         const char *const *quals, uint64_t n_reads, int paired, uint64_t *len_out,
                         uint64_t *n_classified, uint64_t *n_unclassified) {
    using Cls = ClassifierGeneric<score::Lex>;
    alignas(Cls) static unsigned char storage[sizeof(Cls)];
    std::memset(storage, 0, sizeof(storage));
    Cls *c = new(storage) Cls(((RefDb *)db)->m, make_gaps(k, gaps), (u8)k, (uint16_t)w, 1,
                              (bool)emit_all, (bool)emit_fastq, (bool)emit_kraken, (bool)canon);
    Encoder<score::Lex> enc(c->enc_);
    std::vector<tax_t> taxa;
    std::string out;
    const int inc = paired ? 2 : 1;
    std::vector<bseq1_t> bs(inc);
    std::vector<std::string> seqs(inc), nm(inc), ql(inc);
    for(uint64_t i = 0; i + inc <= n_reads; i += inc) {
        for(int m = 0; m < inc; ++m) {
            seqs[m].assign(bases + offsets[i + m], offsets[i + m + 1] - offsets[i + m]);
            nm[m] = names[i + m];
            std::memset(&bs[m], 0, sizeof(bseq1_t));
            bs[m].name = nm[m].data();
            bs[m].seq = seqs[m].data();
            bs[m].l_seq = (int)seqs[m].size();
            if(quals && quals[i + m]) { ql[m] = quals[i + m]; bs[m].qual = ql[m].data(); }
        }
        bs[0].sam = (char *)std::malloc(16); bs[0].l_sam = 0;
        classify_seq(*c, enc, ((RefTax *)tax)->m, bs.data(), paired, taxa);
        out.append(bs[0].sam, bs[0].l_sam);
        std::free(bs[0].sam);
    }
    if(n_classified)   *n_classified = c->n_classified();
    if(n_unclassified) *n_unclassified = c->n_unclassified();
    c->~Cls();
    char *ret = (char *)std::malloc(out.size() + 1);
    std::memcpy(ret, out.data(), out.size());
    ret[out.size()] = 0;
    *len_out = out.size();
    return ret;
}
void bref_free(void *p) { std::free(p); }

} // extern "C"
