/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C restatement of the reference's (dnbaker/bonsai @ 6741de9c) `classify` hot path:
 * Spacer, Encoder::for_each and its mode bodies, QueueMap, CircusEnt, lex_score, khash kh_get/kh_put,
 * linear::counter, resolve_tree, lca, build_parent_map, update_lca_map and the Kraken/FASTQ text
 * emitters. Every function cites the reference file:line it follows (paths relative to
 * /root/reference). It is pinned against (a) the reference's own known-answer tests
 * (test/encoding.cpp:84,122,146-147,194), (b) golden vectors produced by the unmodified reference
 * headers compiled here (oracle/ref_driver.cpp -> oracle/_ref, tests/golden/make_golden.py) and
 * (c) live differential tests against oracle/_ref whenever that library is present.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 */
#ifndef BNS_ORACLE_H
#define BNS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BO_MAXK 32
#define BO_OVERFLOW (~(uint64_t)0)      /* ENCODE_OVERFLOW, encoder.h:119 */

enum { BO_SCORE_LEX = 0, BO_SCORE_ENTROPY = 1 };
/* api 0: Encoder::for_each(fn, str, l) dispatch (encoder.h:416-442) -- what classify_seq calls
 * api 1: the path/kseq overloads applied to one record (encoder.h:448-464) -- what the DB builder calls */
/* api 2: the call-by-call surface (encoder.h:201-206,594-628): assign(), then next_canonicalized_minimizer() (canon) or
 *        next_minimizer() once per position while has_next_kmer(); every return value from the first full window on is
 *        reported, ENCODE_OVERFLOW included (the W-1 calls before it return ENCODE_OVERFLOW by construction) */
enum { BO_API_STRING = 0, BO_API_PATH = 1, BO_API_ITER = 2 };
/* (u64) of an out-of-range double is UB in C++; x86-64 gives one of two behaviours:
 * SATURATE = AVX-512 vcvttsd2usi, WRAP = pre-AVX-512 cvttsd2si sequence. */
enum { BO_CAST_SATURATE = 0, BO_CAST_WRAP = 1 };

typedef struct {
    uint32_t k, c, w;               /* spacer.h:53-55 */
    uint16_t s[BO_MAXK];            /* offsets = gap + 1, spacer.h:65 */
    int unspaced, unwindowed;       /* spacer.h:82-87 */
} bo_spacer;

typedef void (*bo_kmer_fn)(uint64_t kmer, void *ctx);

uint64_t bo_lex_score(uint64_t x);
uint64_t bo_wang64(uint64_t x);
uint64_t bo_rc(uint64_t x, unsigned k);
uint64_t bo_canonical(uint64_t x, unsigned k);
uint64_t bo_cast_u64(double x, int cast_mode);
int  bo_host_cast_mode(void);
int  bo_spacer_init(bo_spacer *sp, unsigned k, unsigned w, const uint16_t *gaps);
void bo_spacer_info(unsigned k, unsigned w, const uint16_t *gaps, uint32_t *c, uint32_t *wout, int *unspaced, int *unwindowed);
int  bo_parse_spacing(const char *s, unsigned k, uint16_t *out, int cap);

void bo_for_each(const bo_spacer *sp, int score, int canon, int api, int cast_mode,
                 const char *seq, uint64_t len, bo_kmer_fn fn, void *ctx);
int64_t bo_encode(unsigned k, unsigned w, const uint16_t *gaps, int score, int canon, int api, int cast_mode,
                  const char *seq, uint64_t len, uint64_t *out, uint64_t cap);

/* taxonomy (khash_t(p)) */
void *bo_tax_load(const char *nodes_dmp);
void *bo_tax_from_pairs(const uint32_t *child, const uint32_t *parent, uint64_t n);
uint64_t bo_tax_size(void *t);
uint64_t bo_tax_pairs(void *t, uint32_t *child, uint32_t *parent, uint64_t cap);
void bo_tax_free(void *t);
uint32_t bo_lca(void *t, uint32_t a, uint32_t b);
uint32_t bo_resolve(void *t, const uint32_t *taxa, const uint16_t *counts, uint32_t n);

/* database (khash_t(c)) */
void *bo_db_new(void);
void *bo_db_from_pairs(const uint64_t *keys, const uint32_t *vals, uint64_t n);
/* wrap caller-owned raw khash arrays (no copy) */
void *bo_db_from_arrays(const uint64_t *keys, const uint32_t *vals, const uint32_t *flags, uint64_t n_buckets);
/* one genome record set -> update_lca_map; seqs are records (contigs) of ONE genome */
void bo_db_add_genome(void *db, void *tax, unsigned k, unsigned w, const uint16_t *gaps, int score, int canon,
                      int cast_mode, const char *bases, const uint64_t *offsets, uint64_t n_records, uint32_t taxid);
void bo_db_arrays(void *db, const uint64_t **keys, const uint32_t **vals, const uint32_t **flags,
                  uint64_t *n_buckets, uint64_t *size);
int  bo_db_get(void *db, uint64_t key, uint32_t *val);
uint64_t bo_db_probe_count(void *db, uint64_t key);
void bo_db_free(void *db);

void bo_classify(void *db, void *tax, unsigned k, unsigned w, const uint16_t *gaps, int score, int canon, int api,
                 int cast_mode, const char *bases, const uint64_t *offsets, uint64_t n_reads, int paired,
                 uint32_t *taxon_out, uint32_t *nhit_out, uint32_t *nmiss_out,
                 uint32_t *taxa_out, const uint64_t *taxa_offsets, int nthreads);

char *bo_classify_text(void *db, void *tax, unsigned k, unsigned w, const uint16_t *gaps, int canon,
                       int emit_all, int emit_fastq, int emit_kraken,
                       const char *bases, const uint64_t *offsets, const char *const *names,
                       const char *const *quals, uint64_t n_reads, int paired, uint64_t *len_out,
                       uint64_t *n_classified, uint64_t *n_unclassified);
void bo_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
