/* TEST INFRASTRUCTURE ONLY -- see bns_oracle.h. Plain-C restatement of the reference hot path.
 * Citations are file:line under /root/reference (dnbaker/bonsai @ 6741de9c). */
#include "bns_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * scalar helpers
 * ---------------------------------------------------------------------------------------- */

/* lex_score == FRev64 == CEIFused<CEIXOR<a>, CEIMul<b>, RotL<31>, CEIXOR<c>>
 * encoder.h:47,59 ; hll/include/sketch/hash.h:688-709,764-853 */
uint64_t bo_lex_score(uint64_t x) {
    x ^= UINT64_C(0x533f8c2151b20f97);
    x *= UINT64_C(0x9a98567ed20c127d);
    x = (x << 31) | (x >> 33);
    return x ^ UINT64_C(0x691a9d706391077a);
}

/* __ac_Wang64_hash, khash64.h:202-211 */
uint64_t bo_wang64(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

/* reverse_complement, kmerutil.h:83-90: swap 2-bit groups end for end, complement, shift down */
uint64_t bo_rc(uint64_t kmer, unsigned k) {
    uint64_t r = 0;
    for(int i = 0; i < 32; ++i) {              /* group i goes to group 31-i */
        r |= ((kmer >> (2 * i)) & 3u) << (2 * (31 - i));
    }
    k &= 0xffu;                                 /* uint8_t parameter */
    unsigned sh = 64u - (k << 1);
    return sh >= 64 ? 0 : (~r) >> sh;           /* k == 0 is never used */
}

/* canonical_representation, kmerutil.h:137-140 */
uint64_t bo_canonical(uint64_t kmer, unsigned k) {
    const uint64_t r = bo_rc(kmer, k);
    return kmer < r ? kmer : r;
}

/* (u64)double as x86-64 executes it. The C++ conversion is UB out of range; the two behaviours
 * a `-march=native` build of the reference can show are restated with defined arithmetic:
 *   SATURATE: vcvttsd2usi -- truncation if the truncated value is in [0, 2^64), else 2^64-1
 *   WRAP:     x < 2^63 ? (u64)(i64)cvttsd2si(x) : cvttsd2si(x - 2^63) ^ 2^63, where cvttsd2si
 *             gives 0x8000000000000000 for NaN or |trunc| >= 2^63 */
static uint64_t cvttsd2si_(double x) {
    if(!(x > -9223372036854775808.0 - 1.0 && x < 9223372036854775808.0)) {
        /* note: -2^63 itself is representable and converts exactly to 0x8000... as well */
        return UINT64_C(0x8000000000000000);
    }
    return (uint64_t)(int64_t)x;
}
uint64_t bo_cast_u64(double x, int cast_mode) {
    if(cast_mode == BO_CAST_SATURATE) {
        if(x != x) return BO_OVERFLOW;
        if(x <= -1.0) return BO_OVERFLOW;
        if(x >= 18446744073709551616.0) return BO_OVERFLOW;
        if(x < 0.0) return 0;                   /* (-1, 0) truncates to 0 */
        return (uint64_t)x;
    }
    if(x != x) return UINT64_C(0x8000000000000000);
    if(x < 9223372036854775808.0) return cvttsd2si_(x);
    return cvttsd2si_(x - 9223372036854775808.0) ^ UINT64_C(0x8000000000000000);
}
int bo_host_cast_mode(void) {
    volatile double d = -5.5;
    /* the probe the reference-driver prints; -5.5 wraps to 0xFFFF...FB or saturates to ~0 */
    return ((uint64_t)d == BO_OVERFLOW) ? BO_CAST_SATURATE : BO_CAST_WRAP;
}

/* alphabet.h:30-43,128 (DNA4) == kmerutil.h:36-48 (cstr_lut): A/a 0, C/c 1, G/g 2, T/t 3, else -1.
 * 'U' stays -1 (the alias code indexes with the *value*). Bytes >= 0x80 index the
 * reference table with a negative char (UB) -- treated as invalid. */
static inline int lut(char ch) {
    switch(ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

/* ------------------------------------------------------------------------------------------
 * Spacer, spacer.h:14-19,49-71 ; parse_spacing, spacer.h:29-47
 * ---------------------------------------------------------------------------------------- */
int bo_spacer_init(bo_spacer *sp, unsigned k, unsigned w, const uint16_t *gaps) {
    if(k < 1 || k > BO_MAXK) return -1;
    memset(sp, 0, sizeof(*sp));
    sp->k = k;
    uint32_t c = k;                                   /* comb_size: spaces.size() + 1 + sum */
    for(unsigned i = 0; i + 1 < k; ++i) {
        uint16_t g = gaps ? gaps[i] : 0;
        c += g;
        sp->s[i] = (uint16_t)(g + 1);                 /* "Convert differences into offsets" */
    }
    sp->c = c;
    sp->w = ((int)c > (int)w) ? c : w;                /* std::max((int)c_, (int)w) */
    sp->unspaced = 1;
    for(unsigned i = 0; i + 1 < k; ++i) if(sp->s[i] != 1) sp->unspaced = 0;
    sp->unwindowed = (sp->k == sp->w);
    return 0;
}
void bo_spacer_info(unsigned k, unsigned w, const uint16_t *gaps, uint32_t *c, uint32_t *wout, int *unspaced, int *unwindowed) {
    bo_spacer sp;
    bo_spacer_init(&sp, k, w, gaps);
    *c = sp.c; *wout = sp.w; *unspaced = sp.unspaced; *unwindowed = sp.unwindowed;
}
int bo_parse_spacing(const char *sss, unsigned k, uint16_t *out, int cap) {
    int n = 0;
    if(!sss || *sss == '\0') {
        for(unsigned i = 0; i + 1 < k; ++i, ++n) if(n < cap) out[n] = 0;
        return n;
    }
    char *ss = (char *)sss;
    for(; *ss; ++ss) {
        const int j = (int)strtoul(ss, &ss, 10);
        if(n < cap) out[n] = (uint16_t)j;
        ++n;
        if(*ss == 'x') {
            ss = strchr(ss, 'x') + 1;
            int m = (int)strtoul(ss, &ss, 10) - 1;
            if(m < 0) m = 0;
            for(int i = 0; i < m; ++i, ++n) if(n < cap) out[n] = (uint16_t)j;
        }
        ss = strchr(ss, ',');
        if(!ss) break;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * QueueMap, qmap.h:17-29,79-96. FIFO of the last W (el, score); the map's begin() is the minimum
 * under (score, el). Equal pairs are ref-counted there, so a plain scan of the FIFO is equivalent.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t *el, *sc;
    uint64_t wsz, head, n;
} qmap_t;
static void qmap_init(qmap_t *q, uint64_t wsz) {
    q->wsz = wsz ? wsz : 1;
    q->el = (uint64_t *)malloc(sizeof(uint64_t) * (q->wsz + 1));
    q->sc = (uint64_t *)malloc(sizeof(uint64_t) * (q->wsz + 1));
    q->head = q->n = 0;
}
static void qmap_free(qmap_t *q) { free(q->el); free(q->sc); }
static uint64_t qmap_min(const qmap_t *q) {
    uint64_t bi = q->head % (q->wsz + 1);
    for(uint64_t j = 1; j < q->n; ++j) {
        uint64_t i = (q->head + j) % (q->wsz + 1);
        if(q->sc[i] < q->sc[bi] || (q->sc[i] == q->sc[bi] && q->el[i] < q->el[bi])) bi = i;
    }
    return q->el[bi];
}
/* next_value, qmap.h:79-87 */
static uint64_t qmap_next(qmap_t *q, uint64_t el, uint64_t score) {
    uint64_t tail = (q->head + q->n) % (q->wsz + 1);
    q->el[tail] = el; q->sc[tail] = score; ++q->n;
    if(q->n > q->wsz) { q->head = (q->head + 1) % (q->wsz + 1); --q->n; }
    if(q->n == q->wsz) return qmap_min(q);
    return BO_OVERFLOW;
}
/* partially_full, qmap.h:93-96 */
static int qmap_partial(const qmap_t *q) { return q->n > 0 && q->n < q->wsz; }

/* ------------------------------------------------------------------------------------------
 * CircusEnt, entropy.h:9-53: counts of the last qsz pushed symbols;
 * value() = sum_c (n_c/qsz) * log(n_c/qsz), NOT_FULL (-1) until qsz symbols were pushed.
 * The sum runs in ska::flat_hash_map iteration order. With fibonacci hashing
 * (flat_hash_map.hpp:1268-1273) keys 0..3 sit in slots 0,4,1,6 of the 8-slot table the map has
 * grown to by the time a third distinct symbol is present (max_load_factor 0.5), i.e. the order
 * is A, G, C, T; with <= 2 symbols present the two-term sum is order-independent.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t q[BO_MAXK + 1];
    uint32_t cnt[4];
    uint32_t cqsz, qsz, head;
} circus_t;
static void circus_clear(circus_t *e) { memset(e->cnt, 0, sizeof(e->cnt)); e->cqsz = 0; e->head = 0; }
static void circus_init(circus_t *e, unsigned qsz) { e->qsz = qsz; circus_clear(e); }
static void circus_push(circus_t *e, int c) {
    ++e->cnt[c];
    if(e->cqsz == e->qsz) {
        --e->cnt[e->q[e->head]];
        e->q[e->head] = (uint8_t)c;
        e->head = (e->head + 1) % e->qsz;
    } else {
        e->q[(e->head + e->cqsz) % e->qsz] = (uint8_t)c;
        ++e->cqsz;
    }
}
static double circus_value(const circus_t *e) {
    if(e->cqsz < e->qsz) return -1.;
    const double qi = 1. / e->qsz;
    static const int order[4] = {0, 2, 1, 3};
    double s = 0.;
    for(int i = 0; i < 4; ++i) {
        const uint32_t n = e->cnt[order[i]];
        if(n) s = s + (double)n * qi * log((double)n * qi);
    }
    return s;
}

/* ------------------------------------------------------------------------------------------
 * Encoder, encoder.h
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const bo_spacer *sp;
    const char *s;
    uint64_t l, pos;
    int score, canon, cast_mode;
    qmap_t q;
    circus_t ent;
} enc_t;

/* Encoder::kmer(start), encoder.h:547-592 (DNA branch). The entropy tracker is cleared and then
 * sees the k-1 symbols after the first one (the first base is never pushed, :550,:566-575). */
static uint64_t enc_kmer(enc_t *e, uint64_t start) {
    const bo_spacer *sp = e->sp;
    if(e->l < sp->c) return BO_OVERFLOW;
    int c0 = lut(e->s[start]);
    uint64_t km = (uint64_t)(int64_t)c0;              /* KmerT(int8_t(-1)) == ~0 */
    if(km == BO_OVERFLOW) return BO_OVERFLOW;
    if(e->score == BO_SCORE_ENTROPY) circus_clear(&e->ent);
    for(unsigned i = 0; i + 1 < sp->k; ++i) {
        start += sp->s[i];
        int nc = lut(e->s[start]);
        if(nc < 0) return BO_OVERFLOW;
        km = (km << 2) | (uint64_t)nc;
        if(e->score == BO_SCORE_ENTROPY) circus_push(&e->ent, nc);
    }
    return km;
}
static int enc_has_next(const enc_t *e) { return (e->pos + e->sp->c - 1) < e->l; }   /* encoder.h:594-597 */

/* scorer_(k, getdata()): score::Lex -> lex_score ; score::Entropy -> ent_score, encoder.h:55-59,76-83 */
static uint64_t enc_score(enc_t *e, uint64_t km) {
    if(e->score == BO_SCORE_ENTROPY)
        return bo_cast_u64((double)km / (circus_value(&e->ent) + 1e-4), e->cast_mode);
    return bo_lex_score(km);
}

static uint64_t kmask(unsigned k) { return BO_OVERFLOW >> (64 - (k << 1)); }         /* rhtraits.h:52-56 */

/* for_each_uncanon_unspaced_unwindowed, encoder.h:240-272 (+ canon wrapper :218-232) */
static void fe_unspaced_unwindowed(enc_t *e, int canon, bo_kmer_fn fn, void *ctx) {
    const unsigned k = e->sp->k;
    const uint64_t mask = kmask(k);
    uint64_t min; unsigned filled;
loop_start:
    min = 0; filled = 0;
    while(e->pos < e->l) {
        while(filled < k && e->pos < e->l) {
            int nv = lut(e->s[e->pos]);
            ++e->pos;
            if(nv < 0) goto loop_start;
            min = (min * 4) | (uint64_t)nv;
            ++filled;
        }
        if(filled == k) {
            min &= mask;
            fn(canon ? bo_canonical(min, k) : min, ctx);
            --filled;
        }
    }
}

/* for_each_uncanon_unspaced_windowed, encoder.h:273-306. The restart test is on the OR-ed word
 * (`(min |= lut) == ENCODE_OVERFLOW`, :283): an invalid base sign-extends to ~0, and for k == 32
 * thirty-two T's do too (the `!= 'T'` guard compares a code with a letter and never fires).
 * The queue survives restarts; a non-empty, non-full queue is flushed at the end (:304-305). */
static void fe_unspaced_windowed(enc_t *e, bo_kmer_fn fn, void *ctx) {
    const unsigned k = e->sp->k;
    const uint64_t mask = kmask(k);
    uint64_t min, kmer; unsigned filled;
windowed_loop_start:
    min = 0; filled = 0;
    while(e->pos < e->l) {
        while(filled < k && e->pos < e->l) {
            min *= 4;
            min |= (uint64_t)(int64_t)lut(e->s[e->pos++]);
            if(min == BO_OVERFLOW) goto windowed_loop_start;
            ++filled;
        }
        if(filled == k) {
            min &= mask;
            if((kmer = qmap_next(&e->q, min, enc_score(e, min))) != BO_OVERFLOW) fn(kmer, ctx);
            --filled;
        }
    }
    if(qmap_partial(&e->q)) fn(qmap_min(&e->q), ctx);
}

/* for_each_uncanon_unspaced_windowed_entropy_, encoder.h:307-346 (+ canon wrapper :347-353):
 * rolling k-mer, rolling symbol counts cleared at every restart; score = kmer / (H + .001) cast to
 * u64 (:337); selection on the forward k-mer, canonicalisation (if any) only on emit. */
static void fe_unspaced_windowed_entropy(enc_t *e, int canon, bo_kmer_fn fn, void *ctx) {
    const unsigned k = e->sp->k;
    const uint64_t mask = kmask(k);
    uint64_t min, kmer; unsigned filled;
windowed_loop_start:
    circus_clear(&e->ent);
    filled = 0; min = 0;
    while(e->pos < e->l) {
        while(filled < k && e->pos < e->l) {
            int nc = lut(e->s[e->pos++]);
            if(nc < 0) goto windowed_loop_start;
            min = (4 * min) | (uint64_t)nc;
            circus_push(&e->ent, nc);
            ++filled;
        }
        if(filled == k) {
            min &= mask;
            const uint64_t sc = bo_cast_u64((double)min / (circus_value(&e->ent) + .001), e->cast_mode);
            if((kmer = qmap_next(&e->q, min, sc)) != BO_OVERFLOW) fn(canon ? bo_canonical(kmer, k) : kmer, ctx);
            --filled;
        }
    }
    if(qmap_partial(&e->q)) {
        kmer = qmap_min(&e->q);
        fn(canon ? bo_canonical(kmer, k) : kmer, ctx);
    }
}

/* for_each_canon_windowed / next_canonicalized_minimizer, encoder.h:211-217,622-628:
 * invalid k-mer (~0) canonicalises to 0 and is pushed like any other; no tail flush. */
static void fe_canon_windowed(enc_t *e, bo_kmer_fn fn, void *ctx) {
    while(enc_has_next(e)) {
        uint64_t nk = enc_kmer(e, e->pos++);
        nk = bo_canonical(nk, e->sp->k);
        const uint64_t m = qmap_next(&e->q, nk, enc_score(e, nk));
        if(m != BO_OVERFLOW) fn(m, ctx);
    }
}
/* for_each_uncanon_spaced / next_minimizer, encoder.h:233-239,616-621 */
static void fe_uncanon_spaced(enc_t *e, bo_kmer_fn fn, void *ctx) {
    while(enc_has_next(e)) {
        const uint64_t k = enc_kmer(e, e->pos++);
        const uint64_t m = qmap_next(&e->q, k, enc_score(e, k));
        if(m != BO_OVERFLOW) fn(m, ctx);
    }
}
/* spaced branch of for_each_canon_unwindowed, encoder.h:226-231 (unreachable through the public
 * constructors because a spaced Spacer is never "unwindowed" unless c == k; kept for completeness) */
static void fe_canon_unwindowed_spaced(enc_t *e, bo_kmer_fn fn, void *ctx) {
    while(enc_has_next(e)) {
        const uint64_t m = enc_kmer(e, e->pos++);
        if(m != BO_OVERFLOW) fn(bo_canonical(m, e->sp->k), ctx);
    }
}

void bo_for_each(const bo_spacer *sp, int score, int canon, int api, int cast_mode,
                 const char *seq, uint64_t len, bo_kmer_fn fn, void *ctx) {
    enc_t e;
    e.sp = sp; e.s = seq; e.l = len; e.pos = 0; e.score = score; e.cast_mode = cast_mode;
    /* ctor, encoder.h:148-150: a spaced seed switches canonicalisation off */
    e.canon = (canon && sp->unspaced) ? 1 : 0;
    qmap_init(&e.q, (uint64_t)sp->w - sp->c + 1);               /* encoder.h:142 */
    circus_init(&e.ent, sp->k);
    /* assign(), encoder.h:201-206, then for_each(), :416-442 */
    if(!enc_has_next(&e)) { qmap_free(&e.q); return; }
    if(api == BO_API_ITER) {
        /* next_canonicalized_minimizer (encoder.h:622-628; canonicalises whatever canonicalize_ says) or next_minimizer
         * (:616-621), one call per position; the calls made before the window fills return ~0 and are not reported */
        const uint64_t wsz = (uint64_t)sp->w - sp->c + 1;
        uint64_t calls = 0;
        while(enc_has_next(&e)) {
            uint64_t nk = enc_kmer(&e, e.pos++);
            if(canon) nk = bo_canonical(nk, sp->k);
            const uint64_t m = qmap_next(&e.q, nk, enc_score(&e, nk));
            if(++calls >= wsz) fn(m, ctx);
        }
    } else if(api == BO_API_STRING) {
        if(e.canon) {
            if(sp->unwindowed) {
                if(sp->unspaced) fe_unspaced_unwindowed(&e, 1, fn, ctx);
                else             fe_canon_unwindowed_spaced(&e, fn, ctx);
            } else if(score == BO_SCORE_ENTROPY && sp->unspaced) fe_unspaced_windowed_entropy(&e, 1, fn, ctx);
            else fe_canon_windowed(&e, fn, ctx);
        } else if(sp->unspaced) {
            if(sp->unwindowed) fe_unspaced_unwindowed(&e, 0, fn, ctx);
            else if(score == BO_SCORE_ENTROPY) fe_unspaced_windowed_entropy(&e, 0, fn, ctx);
            else fe_unspaced_windowed(&e, fn, ctx);
        } else {
            /* encoder.h:437-440: `if(canonicalize_) for_each_uncanon_spaced(func);` inside the
             * !canonicalize_ branch -- dead: spaced seeds emit nothing through this overload */
        }
    } else {
        /* for_each_canon / for_each_uncanon (kseq overloads), encoder.h:448-464 */
        if(e.canon) {
            if(sp->unwindowed) {
                if(sp->unspaced) fe_unspaced_unwindowed(&e, 1, fn, ctx);
                else             fe_canon_unwindowed_spaced(&e, fn, ctx);
            } else fe_canon_windowed(&e, fn, ctx);
        } else if(sp->unspaced) {
            if(sp->unwindowed) fe_unspaced_unwindowed(&e, 0, fn, ctx);
            else               fe_unspaced_windowed(&e, fn, ctx);
        } else fe_uncanon_spaced(&e, fn, ctx);
    }
    qmap_free(&e.q);
}

typedef struct { uint64_t *out, cap, n; } collect_t;
static void collect_fn(uint64_t km, void *ctx) {
    collect_t *c = (collect_t *)ctx;
    if(c->n < c->cap) c->out[c->n] = km;
    ++c->n;
}
int64_t bo_encode(unsigned k, unsigned w, const uint16_t *gaps, int score, int canon, int api, int cast_mode,
                  const char *seq, uint64_t len, uint64_t *out, uint64_t cap) {
    bo_spacer sp;
    if(bo_spacer_init(&sp, k, w, gaps)) return -1;
    collect_t c = {out, cap, 0};
    bo_for_each(&sp, score, canon, api, cast_mode, seq, len, collect_fn, &c);
    return (int64_t)c.n;
}

/* ------------------------------------------------------------------------------------------
 * khash (klib 0.2.8 widened to 64-bit khint_t), khash64.h:169-177,198-263,327-372.
 * Generic over the hash function: khash_t(c) uses Wang64 (util.h:160 via KHASH_MAP_INIT_INT64 ->
 * __ac_Wang64_hash... see khash64.h:651-656), khash_t(p) the identity (kh_int_hash_func).
 * ---------------------------------------------------------------------------------------- */
#define FL_ISEMPTY(f, i)  ((f[(i) >> 4] >> (((i) & 0xfU) << 1)) & 2)
#define FL_ISDEL(f, i)    ((f[(i) >> 4] >> (((i) & 0xfU) << 1)) & 1)
#define FL_ISEITHER(f, i) ((f[(i) >> 4] >> (((i) & 0xfU) << 1)) & 3)
#define FL_SET_BOTH_FALSE(f, i) (f[(i) >> 4] &= ~(3u << (((i) & 0xfU) << 1)))
#define FL_FSIZE(m) ((m) < 16 ? 1 : (m) >> 4)
static const double HASH_UPPER = 0.77;                       /* khash64.h:198 */

typedef struct {
    uint64_t n_buckets, size, n_occupied, upper_bound;
    uint32_t *flags;
    uint64_t *keys;      /* khash_t(p) stores its u32 keys widened; layout is not observable */
    uint32_t *vals;
    int wang;            /* 1: Wang64 (khash_t(c)), 0: identity (khash_t(p)) */
    int borrowed;
} kh_t;

static inline uint64_t kh_hash(const kh_t *h, uint64_t key) { return h->wang ? bo_wang64(key) : (uint64_t)(uint32_t)key; }

/* kh_get, khash64.h:250-263: triangular probing, stop at an empty slot */
static uint64_t kh_get_(const kh_t *h, uint64_t key) {
    if(!h->n_buckets) return 0;
    const uint64_t mask = h->n_buckets - 1;
    uint64_t i = kh_hash(h, key) & mask, last = i, step = 0;
    while(!FL_ISEMPTY(h->flags, i) && (FL_ISDEL(h->flags, i) || h->keys[i] != key)) {
        i = (i + (++step)) & mask;
        if(i == last) return h->n_buckets;
    }
    return FL_ISEITHER(h->flags, i) ? h->n_buckets : i;
}
static uint64_t kh_probe_count_(const kh_t *h, uint64_t key) {
    if(!h->n_buckets) return 0;
    const uint64_t mask = h->n_buckets - 1;
    uint64_t i = kh_hash(h, key) & mask, last = i, step = 0, n = 1;
    while(!FL_ISEMPTY(h->flags, i) && (FL_ISDEL(h->flags, i) || h->keys[i] != key)) {
        i = (i + (++step)) & mask; ++n;
        if(i == last) break;
    }
    return n;
}
static void kh_resize_(kh_t *h, uint64_t new_n) {
    /* kroundup64 + minimum 4, khash64.h:268-270; rebuilt out of place (the reference kicks
     * keys around in place, :290-320 -- the resulting slot assignment is not observable) */
    uint64_t nb = 4;
    while(nb < new_n) nb <<= 1;
    if(h->size >= (uint64_t)(nb * HASH_UPPER + 0.5)) return;
    kh_t o = *h;
    h->n_buckets = nb;
    h->flags = (uint32_t *)malloc(FL_FSIZE(nb) * sizeof(uint32_t));
    memset(h->flags, 0xaa, FL_FSIZE(nb) * sizeof(uint32_t));
    h->keys = (uint64_t *)malloc(nb * sizeof(uint64_t));
    h->vals = (uint32_t *)malloc(nb * sizeof(uint32_t));
    const uint64_t mask = nb - 1;
    for(uint64_t j = 0; j < o.n_buckets; ++j) {
        if(FL_ISEITHER(o.flags, j)) continue;
        uint64_t i = kh_hash(h, o.keys[j]) & mask, step = 0;
        while(!FL_ISEMPTY(h->flags, i)) i = (i + (++step)) & mask;
        FL_SET_BOTH_FALSE(h->flags, i);
        h->keys[i] = o.keys[j]; h->vals[i] = o.vals[j];
    }
    h->n_occupied = h->size;
    h->upper_bound = (uint64_t)(nb * HASH_UPPER + 0.5);
    if(!o.borrowed) { free(o.flags); free(o.keys); free(o.vals); }
    h->borrowed = 0;
}
/* kh_put, khash64.h:327-372 (no deletions ever happen on this path) */
static uint64_t kh_put_(kh_t *h, uint64_t key, int *ret) {
    if(h->n_occupied >= h->upper_bound) {
        if(h->n_buckets > (h->size << 1)) kh_resize_(h, h->n_buckets - 1);
        else kh_resize_(h, h->n_buckets + 1);
    }
    const uint64_t mask = h->n_buckets - 1;
    uint64_t i = kh_hash(h, key) & mask, step = 0;
    while(!FL_ISEMPTY(h->flags, i) && h->keys[i] != key) i = (i + (++step)) & mask;
    if(FL_ISEMPTY(h->flags, i)) {
        h->keys[i] = key;
        FL_SET_BOTH_FALSE(h->flags, i);
        ++h->size; ++h->n_occupied;
        *ret = 1;
    } else *ret = 0;
    return i;
}
static kh_t *kh_new(int wang) {
    kh_t *h = (kh_t *)calloc(1, sizeof(kh_t));
    h->wang = wang;
    return h;
}
static void kh_free_(kh_t *h) {
    if(!h) return;
    if(!h->borrowed) { free(h->flags); free(h->keys); free(h->vals); }
    free(h);
}

/* ------------------------------------------------------------------------------------------
 * taxonomy: khash_t(p) child -> parent. build_parent_map, util.h:766-785
 * ---------------------------------------------------------------------------------------- */
void *bo_tax_from_pairs(const uint32_t *child, const uint32_t *parent, uint64_t n) {
    kh_t *m = kh_new(0);
    int r;
    uint64_t ki;
    for(uint64_t i = 0; i < n; ++i) { ki = kh_put_(m, child[i], &r); m->vals[ki] = parent[i]; }
    ki = kh_put_(m, 1, &r);                               /* "Root of the tree", util.h:780-781 */
    m->vals[ki] = 0;
    return m;
}
void *bo_tax_load(const char *fn) {
    FILE *fp = fopen(fn, "r");
    if(!fp) return NULL;
    kh_t *m = kh_new(0);
    int r;
    char *line = NULL; size_t cap = 0; ssize_t len;
    while((len = getline(&line, &cap, fp)) >= 0) {
        if(len && line[len - 1] == '\n') line[len - 1] = 0;   /* std::getline strips the newline */
        switch(line[0]) { case '\n': case '\0': case '#': continue; }
        uint64_t ki = kh_put_(m, (uint32_t)atoi(line), &r);
        const char *p = strchr(line, '|');
        m->vals[ki] = p ? (uint32_t)atoi(p + 2) : (uint32_t)-1;
    }
    free(line);
    fclose(fp);
    { uint64_t ki = kh_put_(m, 1, &r); m->vals[ki] = 0; }
    if(m->size < 2) { kh_free_(m); return NULL; }         /* RUNTIME_ERROR, util.h:782 */
    return m;
}
uint64_t bo_tax_size(void *t) { return ((kh_t *)t)->size; }
uint64_t bo_tax_pairs(void *t, uint32_t *child, uint32_t *parent, uint64_t cap) {
    kh_t *m = (kh_t *)t; uint64_t n = 0;
    for(uint64_t i = 0; i < m->n_buckets; ++i)
        if(!FL_ISEITHER(m->flags, i)) {
            if(n < cap) child[n] = (uint32_t)m->keys[i], parent[n] = m->vals[i];
            ++n;
        }
    return n;
}
void bo_tax_free(void *t) { kh_free_((kh_t *)t); }

/* linear::set<tax_t>, linear/linear.h:83-100: insertion-ordered, find = linear scan */
typedef struct { uint32_t *d; uint32_t n, m; } lset_t;
static int lset_has(const lset_t *s, uint32_t v) { for(uint32_t i = 0; i < s->n; ++i) if(s->d[i] == v) return 1; return 0; }
static void lset_insert(lset_t *s, uint32_t v) {
    if(lset_has(s, v)) return;
    if(s->n == s->m) { s->m = s->m ? s->m * 2 : 8; s->d = (uint32_t *)realloc(s->d, s->m * sizeof(uint32_t)); }
    s->d[s->n++] = v;
}

/* lca, util.h:634-663 */
static uint32_t lca_(const kh_t *map, uint32_t a, uint32_t b) {
    if(a == b) return a;
    if(b == 0) return a;
    if(a == 0) return b;
    lset_t nodes = {0, 0, 0};
    uint64_t ki;
    uint32_t ret = 1;
    while(a) {
        lset_insert(&nodes, a);
        if((ki = kh_get_(map, a)) == map->n_buckets) { ret = (uint32_t)-1; goto done; }
        a = map->vals[ki];
    }
    while(b) {
        if(lset_has(&nodes, b)) { ret = b; goto done; }
        if((ki = kh_get_(map, b)) == map->n_buckets) { ret = (uint32_t)-1; goto done; }
        b = map->vals[ki];
    }
done:
    free(nodes.d);
    return ret;
}
uint32_t bo_lca(void *t, uint32_t a, uint32_t b) { return lca_((kh_t *)t, a, b); }

/* linear::counter<tax_t, u16>, linear/linear.h:182-244: first-seen order, u16 counts */
typedef struct { uint32_t *keys; uint16_t *vals; uint32_t n, m; } counter_t;
static void counter_add(counter_t *c, uint32_t key, uint16_t inc) {
    for(uint32_t i = 0; i < c->n; ++i) if(c->keys[i] == key) { c->vals[i] = (uint16_t)(c->vals[i] + inc); return; }
    if(c->n == c->m) {
        c->m = c->m ? c->m * 2 : 8;
        c->keys = (uint32_t *)realloc(c->keys, c->m * sizeof(uint32_t));
        c->vals = (uint16_t *)realloc(c->vals, c->m * sizeof(uint16_t));
    }
    c->keys[c->n] = key; c->vals[c->n] = inc; ++c->n;
}
static uint16_t counter_count(const counter_t *c, uint32_t key) {
    for(uint32_t i = 0; i < c->n; ++i) if(c->keys[i] == key) return c->vals[i];
    return 0;
}

/* resolve_tree, util.h:831-869. A taxon missing from the parent map is UB in the reference
 * (kh_val at kh_end, SURVEY B-9); here the walk just stops (never exercised by the tests). */
static uint32_t resolve_tree_(const counter_t *hc, const kh_t *pm) {
    lset_t max_taxa = {0, 0, 0};
    uint32_t max_taxon = 0, max_score = 0;
    for(uint32_t i = 0; i < hc->n; ++i) {
        uint32_t taxon = hc->keys[i], node = taxon, score = 0;
        while(node) {
            score += counter_count(hc, node);
            uint64_t ki = kh_get_(pm, node);
            if(ki == pm->n_buckets) break;
            node = pm->vals[ki];
        }
        if(score > max_score) {
            max_taxa.n = 0;
            max_score = score;
            max_taxon = taxon;
        } else if(score == max_score) {
            if(max_taxa.n == 0) lset_insert(&max_taxa, max_taxon);
            lset_insert(&max_taxa, taxon);
        }
    }
    if(max_taxa.n) {
        max_taxon = max_taxa.d[0];
        for(uint32_t i = 1; i < max_taxa.n; ++i) max_taxon = lca_(pm, max_taxon, max_taxa.d[i]);
    }
    free(max_taxa.d);
    return max_taxon;
}
uint32_t bo_resolve(void *t, const uint32_t *taxa, const uint16_t *counts, uint32_t n) {
    counter_t hc = {0, 0, 0, 0};
    for(uint32_t i = 0; i < n; ++i) counter_add(&hc, taxa[i], counts[i]);
    uint32_t r = resolve_tree_(&hc, (kh_t *)t);
    free(hc.keys); free(hc.vals);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * database: khash_t(c) k-mer -> taxid
 * ---------------------------------------------------------------------------------------- */
void *bo_db_new(void) { return kh_new(1); }
void *bo_db_from_pairs(const uint64_t *keys, const uint32_t *vals, uint64_t n) {
    kh_t *m = kh_new(1);
    int r;
    for(uint64_t i = 0; i < n; ++i) { uint64_t ki = kh_put_(m, keys[i], &r); m->vals[ki] = vals[i]; }
    return m;
}
void *bo_db_from_arrays(const uint64_t *keys, const uint32_t *vals, const uint32_t *flags, uint64_t n_buckets) {
    kh_t *m = kh_new(1);
    m->keys = (uint64_t *)keys; m->vals = (uint32_t *)vals; m->flags = (uint32_t *)flags;
    m->n_buckets = n_buckets; m->borrowed = 1;
    for(uint64_t i = 0; i < n_buckets; ++i) if(!FL_ISEITHER(flags, i)) ++m->size;
    m->n_occupied = m->size; m->upper_bound = (uint64_t)(n_buckets * HASH_UPPER + 0.5);
    return m;
}
typedef struct { kh_t *set; } setctx_t;
static void set_fn(uint64_t km, void *ctx) {           /* fill_set_genome's lambda, feature_min.h:73-79 */
    int r;
    kh_put_(((setctx_t *)ctx)->set, km, &r);
}
/* fill_set_genome (feature_min.h:68-83) over the records of one genome through the path overload,
 * then update_lca_map (feature_min.h:205-228): insert with taxid, or replace by lca(tax, taxid, old) */
void bo_db_add_genome(void *db, void *tax, unsigned k, unsigned w, const uint16_t *gaps, int score, int canon,
                      int cast_mode, const char *bases, const uint64_t *offsets, uint64_t n_records, uint32_t taxid) {
    bo_spacer sp;
    if(bo_spacer_init(&sp, k, w, gaps)) return;
    kh_t *kc = (kh_t *)db;
    setctx_t sc = {kh_new(1)};
    for(uint64_t r = 0; r < n_records; ++r)
        bo_for_each(&sp, score, canon, BO_API_PATH, cast_mode, bases + offsets[r], offsets[r + 1] - offsets[r], set_fn, &sc);
    int ret;
    for(uint64_t i = 0; i < sc.set->n_buckets; ++i) {
        if(FL_ISEITHER(sc.set->flags, i)) continue;
        const uint64_t key = sc.set->keys[i];
        uint64_t k2 = kh_get_(kc, key);
        if(k2 == kc->n_buckets) {
            k2 = kh_put_(kc, key, &ret);
            kc->vals[k2] = taxid;
        } else if(kc->vals[k2] != taxid) {
            kc->vals[k2] = lca_((kh_t *)tax, taxid, kc->vals[k2]);
        }
    }
    kh_free_(sc.set);
}
void bo_db_arrays(void *db, const uint64_t **keys, const uint32_t **vals, const uint32_t **flags,
                  uint64_t *n_buckets, uint64_t *size) {
    kh_t *m = (kh_t *)db;
    *keys = m->keys; *vals = m->vals; *flags = m->flags; *n_buckets = m->n_buckets; *size = m->size;
}
int bo_db_get(void *db, uint64_t key, uint32_t *val) {
    kh_t *m = (kh_t *)db;
    uint64_t ki = kh_get_(m, key);
    if(ki == m->n_buckets) return 0;
    *val = m->vals[ki];
    return 1;
}
uint64_t bo_db_probe_count(void *db, uint64_t key) { return kh_probe_count_((kh_t *)db, key); }
void bo_db_free(void *db) { kh_free_((kh_t *)db); }

/* ------------------------------------------------------------------------------------------
 * classify_seq core, classifier.h:213-238 (taxid only; the CPU-baseline loop of BASELINE.md 3)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const kh_t *db;
    counter_t hc;
    uint32_t missing, nhit;
    uint32_t *taxa; uint32_t taxa_cap;     /* std::vector<tax_t> taxa */
    int keep_taxa;
} cls_ctx_t;
static void cls_fn(uint64_t kmer, void *ctx_) {           /* the lambda at classifier.h:225-229 */
    cls_ctx_t *c = (cls_ctx_t *)ctx_;
    uint64_t ki = kh_get_(c->db, kmer);
    if(ki == c->db->n_buckets) { ++c->missing; return; }
    const uint32_t v = c->db->vals[ki];
    if(c->keep_taxa) {
        if(c->nhit == c->taxa_cap) { c->taxa_cap = c->taxa_cap ? c->taxa_cap * 2 : 256; c->taxa = (uint32_t *)realloc(c->taxa, c->taxa_cap * sizeof(uint32_t)); }
        c->taxa[c->nhit] = v;
    }
    ++c->nhit;
    counter_add(&c->hc, v, 1);
}

void bo_classify(void *db, void *tax, unsigned k, unsigned w, const uint16_t *gaps, int score, int canon, int api,
                 int cast_mode, const char *bases, const uint64_t *offsets, uint64_t n_reads, int paired,
                 uint32_t *taxon_out, uint32_t *nhit_out, uint32_t *nmiss_out,
                 uint32_t *taxa_out, const uint64_t *taxa_offsets, int nthreads) {
    bo_spacer sp;
    if(bo_spacer_init(&sp, k, w, gaps)) return;
    const int inc = paired ? 2 : 1;
    const int64_t nrec = (int64_t)(n_reads / inc);
#ifdef _OPENMP
    if(nthreads <= 0) nthreads = omp_get_max_threads();
#else
    (void)nthreads;
#endif
    #pragma omp parallel num_threads(nthreads)
    {
        cls_ctx_t c;
        memset(&c, 0, sizeof(c));
        c.db = (const kh_t *)db;
        c.keep_taxa = taxa_out != NULL;
        #pragma omp for schedule(dynamic, 1024)
        for(int64_t r = 0; r < nrec; ++r) {
            const uint64_t i = (uint64_t)r * inc;
            c.hc.n = 0; c.missing = 0; c.nhit = 0;
            for(int m = 0; m < inc; ++m)
                bo_for_each(&sp, score, canon, api, cast_mode, bases + offsets[i + m], offsets[i + m + 1] - offsets[i + m], cls_fn, &c);
            taxon_out[r] = resolve_tree_(&c.hc, (const kh_t *)tax);
            if(nhit_out)  nhit_out[r] = c.nhit;
            if(nmiss_out) nmiss_out[r] = c.missing;
            if(taxa_out)  memcpy(taxa_out + taxa_offsets[r], c.taxa, c.nhit * sizeof(uint32_t));
        }
        free(c.hc.keys); free(c.hc.vals); free(c.taxa);
    }
}

/* ------------------------------------------------------------------------------------------
 * text emitters, classifier.h:30-129,232-246 ; integer printing kspp/ks.h:318-372
 * ---------------------------------------------------------------------------------------- */
typedef struct { char *s; size_t l, m; } sbuf_t;
static void sb_putn(sbuf_t *b, const char *p, size_t n) {
    if(b->l + n + 1 > b->m) { b->m = (b->l + n + 1) * 2; b->s = (char *)realloc(b->s, b->m); }
    memcpy(b->s + b->l, p, n); b->l += n; b->s[b->l] = 0;
}
static void sb_putc(sbuf_t *b, char c) { sb_putn(b, &c, 1); }
static void sb_putu(sbuf_t *b, uint32_t x) { char t[16]; int n = snprintf(t, sizeof t, "%u", x); sb_putn(b, t, (size_t)n); }
static void sb_puti(sbuf_t *b, long x) { char t[32]; int n = snprintf(t, sizeof t, "%ld", x); sb_putn(b, t, (size_t)n); }

static void append_taxa_run(uint32_t last, uint32_t run, sbuf_t *b) {      /* classifier.h:30-43 */
    if(last == 0) sb_putc(b, 'U');
    else if(last == (uint32_t)-1) sb_putc(b, 'A');
    else sb_putu(b, last);
    sb_putc(b, ':'); sb_putu(b, run); sb_putc(b, '\t');
}
static void append_taxa_runs(uint32_t taxon, const uint32_t *taxa, uint32_t n, sbuf_t *b) {   /* :46-61 */
    if(taxon) {
        uint32_t last = taxa[0], run = 1;
        for(uint32_t i = 1; i != n; ++i) {
            if(taxa[i] == last) ++run;
            else { append_taxa_run(last, run, b); last = taxa[i]; run = 1; }
        }
        append_taxa_run(last, run, b);
        b->s[b->l - 1] = '\n';
    } else sb_putn(b, "0:0\n", 4);
}
static void append_counts(uint32_t count, char ch, sbuf_t *b) {              /* :63-70 */
    if(count) { sb_putc(b, ch); sb_putc(b, ':'); sb_putu(b, count); sb_putc(b, '\t'); }
}

char *bo_classify_text(void *db, void *tax, unsigned k, unsigned w, const uint16_t *gaps, int canon,
                       int emit_all, int emit_fastq, int emit_kraken,
                       const char *bases, const uint64_t *offsets, const char *const *names,
                       const char *const *quals, uint64_t n_reads, int paired, uint64_t *len_out,
                       uint64_t *n_classified, uint64_t *n_unclassified) {
    bo_spacer sp;
    if(bo_spacer_init(&sp, k, w, gaps)) return NULL;
    const int cast_mode = bo_host_cast_mode();
    const int inc = paired ? 2 : 1;
    const int flag = (emit_kraken ? 1 : 0) | (emit_fastq ? 2 : 0) | (emit_all ? 4 : 0);   /* :23-27 */
    sbuf_t out = {NULL, 0, 0};
    sb_putn(&out, "", 0);
    uint64_t ncls[2] = {0, 0};
    cls_ctx_t c;
    memset(&c, 0, sizeof(c));
    c.db = (const kh_t *)db; c.keep_taxa = 1;
    for(uint64_t i = 0; i + inc <= n_reads; i += inc) {
        c.hc.n = 0; c.missing = 0; c.nhit = 0;
        const uint64_t l0 = offsets[i + 1] - offsets[i];
        bo_for_each(&sp, BO_SCORE_LEX, canon, BO_API_STRING, cast_mode, bases + offsets[i], l0, cls_fn, &c);
        /* classifier.h:232: unsigned ambig_count(bs->l_seq - enc.sp_.c_ + 1 - taxa.size() - missing_count) */
        uint32_t ambig = (uint32_t)((uint64_t)(uint32_t)((uint32_t)(int)l0 - sp.c + 1) - (uint64_t)c.nhit - c.missing);
        uint64_t l1 = 0;
        if(paired) {
            l1 = offsets[i + 2] - offsets[i + 1];
            bo_for_each(&sp, BO_SCORE_LEX, canon, BO_API_STRING, cast_mode, bases + offsets[i + 1], l1, cls_fn, &c);
            /* :235 (re-subtracts the cumulative counts; reporting bug kept) */
            ambig += (uint32_t)((uint64_t)(uint32_t)((uint32_t)(int)l1 - (sp.c - 1)) - (uint64_t)c.nhit - c.missing);
        }
        const uint32_t taxon = resolve_tree_(&c.hc, (const kh_t *)tax);
        ++ncls[!taxon];
        if(!(emit_all || taxon)) continue;
        const char *seq0 = bases + offsets[i];
        if(flag & 2) {                                   /* any FASTQ combination, :241-242 */
            /* append_fastq_classification, classifier.h:72-108 */
            sb_putn(&out, names[i], strlen(names[i]));
            sb_putc(&out, ' ');
            const size_t cms = out.l;
            sb_putc(&out, taxon == 0 ? 'U' : 'C'); sb_putc(&out, '\t');
            sb_putu(&out, taxon); sb_putc(&out, '\t');
            sb_puti(&out, (long)(int)l0); sb_putc(&out, '\t');
            append_counts(c.missing, 'M', &out);
            append_counts(ambig, 'A', &out);
            if(emit_kraken) append_taxa_runs(taxon, c.taxa, c.nhit, &out);
            else out.s[out.l - 1] = '\n';
            const size_t cme = out.l;
            sb_putn(&out, seq0, l0);
            sb_putn(&out, "\n+\n", 3);
            sb_putn(&out, (quals && quals[i]) ? quals[i] : seq0, l0);
            sb_putc(&out, '\n');
            if(paired) {
                sb_putn(&out, names[i + 1], strlen(names[i + 1]));
                sb_putc(&out, ' ');
                char *cm = (char *)malloc(cme - cms);
                memcpy(cm, out.s + cms, cme - cms);
                sb_putn(&out, cm, cme - cms);
                free(cm);
                sb_putc(&out, '\n');
                const char *seq1 = bases + offsets[i + 1];
                sb_putn(&out, seq1, l1);
                sb_putn(&out, "\n+\n", 3);
                sb_putn(&out, (quals && quals[i + 1]) ? quals[i + 1] : seq1, l1);
                sb_putc(&out, '\n');
            }
        } else if(flag & 1) {                            /* KRAKEN, EMIT_ALL|KRAKEN, :243-244 */
            /* append_kraken_classification, classifier.h:112-129 */
            sb_putc(&out, taxon ? 'C' : 'U'); sb_putc(&out, '\t');
            sb_putn(&out, names[i], strlen(names[i])); sb_putc(&out, '\t');
            sb_putu(&out, taxon); sb_putc(&out, '\t');
            sb_puti(&out, (long)(int)l0); sb_putc(&out, '\t');
            append_counts(c.missing, 'M', &out);
            append_counts(ambig, 'A', &out);
            append_taxa_runs(taxon, c.taxa, c.nhit, &out);
        }
        /* flag == 0 or EMIT_ALL alone: the switch matches no case -> no text */
    }
    free(c.hc.keys); free(c.hc.vals); free(c.taxa);
    if(n_classified)   *n_classified = ncls[0];
    if(n_unclassified) *n_unclassified = ncls[1];
    *len_out = out.l;
    return out.s;
}
void bo_free(void *p) { free(p); }
