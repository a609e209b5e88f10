/* bonsai_b200.h -- C ABI of libbonsai_b200.so: the B200-native `classify` hot path of dnbaker/bonsai.
 *
 * The reference has no FFI layer; its boundary for this path is the C++ template surface in
 * namespace bns. Each entry point below names the reference interface it replaces
 * (file:line under the reference tree @ 6741de9c). include/bonsai_b200/bonsai.hpp re-creates those C++
 * names (Spacer, Encoder<Score>::for_each, ClassifierGeneric, classify_seqs, process_dataset, Database)
 * on top of this ABI, and INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: every call returns 0 (BNS_OK) or a negative BNS_E_* code and never exits or throws;
 * bns_b200_last_error() gives the message. All buffers are caller-owned. Host-pointer calls block
 * until their outputs are written. A context is bound to one CUDA device (one process per GPU);
 * it is thread-compatible (one call at a time per context). There is NO CPU fallback: without a
 * usable CUDA device bns_b200_open() fails with BNS_E_CUDA.
 */
#ifndef BONSAI_B200_H
#define BONSAI_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNS_B200_ABI_VERSION 1
#define BNS_MAX_K 32

enum {
    BNS_OK = 0,
    BNS_E_INVAL = -1,      /* bad argument / unsupported configuration */
    BNS_E_CUDA = -2,       /* CUDA runtime error (no device, launch failure, ...) */
    BNS_E_NOMEM = -3,      /* host or device allocation failed */
    BNS_E_STATE = -4,      /* call order: table / taxonomy not loaded yet */
    BNS_E_TAXONOMY = -5,   /* malformed taxonomy (cycle) */
    BNS_E_CAPACITY = -6,   /* an output buffer the caller sized is too small */
    BNS_E_IO = -7          /* file could not be read / parsed */
};

/* score::Lex / score::Entropy -- include/bonsai/encoder.h:76-89 */
enum { BNS_SCORE_LEX = 0, BNS_SCORE_ENTROPY = 1 };
/* Which overload family of Encoder is reproduced:
 *   BNS_API_STRING  Encoder::for_each(fn, const char*, u64)          encoder.h:416-442 (classify_seq uses it;
 *                   spaced seeds emit nothing there -- reference quirk)
 *   BNS_API_PATH    Encoder::for_each_canon / for_each_uncanon on one record, encoder.h:448-464 (what the
 *                   database builder uses; spaced seeds work: for_each_uncanon_spaced, encoder.h:233) */
/*   BNS_API_ITER    the call-by-call surface: assign(), then next_canonicalized_minimizer() (canonicalize != 0) or
 *                   next_minimizer() (== next_kmer() when the window is the comb) once per position while
 *                   has_next_kmer(), encoder.h:201-206,594-628. Nothing is filtered: the stream holds the value of every
 *                   call from the first full window on (the W-1 calls before it return ENCODE_OVERFLOW by
 *                   construction, qmap.h:79-87), ENCODE_OVERFLOW included where an invalid k-mer wins its window.
 *                   `canonicalize` is taken as given here, also for spaced seeds (next_canonicalized_minimizer does not
 *                   look at canonicalize_). Encode only: the classify and build calls refuse this configuration. */
enum { BNS_API_STRING = 0, BNS_API_PATH = 1, BNS_API_ITER = 2 };
/* The entropy score casts a negative double to u64 (encoder.h:337, :55-58), which is UB; an x86-64 build of
 * the reference saturates (AVX-512 vcvttsd2usi) or wraps (cvttsd2si sequence) depending on -march. */
enum { BNS_CAST_SATURATE = 0, BNS_CAST_WRAP = 1 };

/* Spacer(k, w, gaps) + Encoder(sp, canonicalize) -- include/bonsai/spacer.h:59-71, encoder.h:133-153 */
typedef struct bns_b200_config {
    uint32_t k;                 /* k-mer length, 1..32 */
    uint32_t w;                 /* window; the Spacer uses max(comb, w) */
    uint16_t gaps[BNS_MAX_K];   /* gaps[0..k-2], 0 = adjacent (spvec_t before the +1 of spacer.h:65) */
    uint32_t score;             /* BNS_SCORE_* */
    uint32_t canonicalize;      /* requested; switched off for spaced seeds as encoder.h:148-150 does */
    uint32_t api;               /* BNS_API_* */
    uint32_t entropy_cast;      /* BNS_CAST_* */
    int32_t  device;            /* CUDA ordinal, -1 = current device */
    uint32_t n_gpus;            /* bns_b200_open_multi: contexts to open when its n_gpus argument is 0 (0 = every visible device) */
    uint32_t host_pack_threads; /* host-buffer classify calls pack the bases to 2 bits on this many worker threads of the caller's machine
                                 * (chunk by chunk, next to chunks that cross as ASCII: 38 instead of 150 bytes per 150 bp read over PCIe).
                                 * 0 = the process's share of the hardware threads, 0xffffffff = off (every chunk crosses as ASCII) */
    uint32_t reserved[5];
} bns_b200_config;

typedef struct bns_b200_ctx bns_b200_t;

typedef struct bns_b200_table_info {
    uint64_t n_keys;            /* entries resident */
    uint64_t n_buckets;         /* 32-byte buckets of 4 slots (power of two) */
    uint64_t bytes;             /* device bytes of the slot array */
    uint32_t bucket_bits, val_bits, n_values, max_disp;
    uint64_t n_displaced;       /* entries not in their home bucket */
    uint64_t n_overflowed;      /* home buckets with the overflow mark */
    uint32_t layout;            /* 0 hash (bucket = mix64), 1 minimizer (group of four 64-byte units by the k-mer's 16-mer minimizer) */
    uint32_t disp_bits;         /* width of the slot's displacement field */
    uint64_t n_stash;           /* layout 1: entries kept in the overflow stash (keys whose probe chain was full) */
} bns_b200_table_info;

typedef struct bns_b200_stats {
    uint64_t n_classified, n_unclassified;   /* ClassifierGeneric::n_classified()/n_unclassified(), classifier.h:170-171 */
    uint64_t kernel_launches;                /* kernels of this library launched since open */
    uint64_t reads_processed, bases_processed;
    uint64_t h2d_bytes, d2h_bytes;
    double   last_kernel_ms;                 /* device time of the last classify/encode kernel (CUDA events) */
    double   kernel_ms_total;
} bns_b200_stats;

const char *bns_b200_version(void);
const char *bns_b200_strerror(int code);
const char *bns_b200_last_error(const bns_b200_t *ctx);

/* ClassifierGeneric(map, spaces, k, wsz, ...) + Encoder copy per worker -- classifier.h:155-166,258 */
int  bns_b200_open(const bns_b200_config *cfg, bns_b200_t **out);
void bns_b200_close(bns_b200_t *ctx);
/* Spacer geometry after construction: comb c_, window w_, unspaced(), unwindowed(), effective canonicalize */
int  bns_b200_geometry(const bns_b200_t *ctx, uint32_t *c, uint32_t *w, int *unspaced, int *unwindowed, int *canon);
/* Upper bound on k-mers Encoder::for_each emits for one sequence of `len` bases */
uint64_t bns_b200_encode_bound(const bns_b200_t *ctx, uint64_t len);

/* ---- database: khash_t(c) -> device table ------------------------------------------------------
 * load_table takes the raw khash arrays the reference holds (struct kh_c_t, include/bonsai/khash64.h:213-219;
 * flags: 2 bits per bucket, bit1 empty / bit0 deleted, :169-177) exactly as Database<khash_t(c)>::db_ exposes
 * them (include/bonsai/database.h:17-31). Only kh_get's result (hit/miss + value, khash64.h:250-263) is
 * preserved; the device layout is this library's own. */
int bns_b200_load_table(bns_b200_t *ctx, const uint64_t *keys, const uint32_t *vals, const uint32_t *flags, uint64_t n_buckets);
/* same, from a dense list of distinct keys */
int bns_b200_load_pairs(bns_b200_t *ctx, const uint64_t *keys, const uint32_t *vals, uint64_t n);
/* same, keys/vals already in device memory (values must come from `values[n_values]`) */
int bns_b200_load_pairs_device(bns_b200_t *ctx, const uint64_t *d_keys, const uint32_t *d_vals, uint64_t n,
                               const uint32_t *values, uint32_t n_values);
int bns_b200_table_info_get(const bns_b200_t *ctx, bns_b200_table_info *info);
/* kh_get + kh_val over a batch of keys (host pointers): found_out[i] = 1/0, vals_out[i] = value if found */
int bns_b200_lookup_batch(bns_b200_t *ctx, const uint64_t *keys, uint64_t n, uint32_t *vals_out, uint8_t *found_out);

/* number of 32-byte buckets (DRAM sectors) the probes of `keys` touch in total: the p-bar of SURVEY 8(d) */
int bns_b200_lookup_sectors(bns_b200_t *ctx, const uint64_t *keys, uint64_t n, uint64_t *sectors_out);

/* ---- database construction on the device: `bonsai build` ----------------------------------------------------
 * fill_set_genome + update_lca_map (include/bonsai/feature_min.h:68-83,205-228): the k-mer / minimizer SET of every
 * genome (encoder = this context's configuration, normally BNS_API_PATH) is inserted with the genome's taxid, or merged
 * with lca(tax, taxid, old) where the key already exists. Needs the taxonomy. `taxids` announces every taxid that
 * add_genome will use (their ancestors become the value dictionary); max_kmers bounds the distinct keys
 * (BNS_E_CAPACITY from build_finish means: begin again with a larger bound). */
int bns_b200_build_begin(bns_b200_t *ctx, uint64_t max_kmers, const uint32_t *taxids, uint32_t n_taxids);
int bns_b200_build_add_genome(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_records, uint32_t taxid);
int bns_b200_build_finish(bns_b200_t *ctx);
/* the resident table as (key, value) pairs in no particular order; cap = 0 only reports the count in *n_out */
int bns_b200_table_dump(bns_b200_t *ctx, uint64_t *keys_out, uint32_t *vals_out, uint64_t cap, uint64_t *n_out);
/* change the Spacer / Encoder configuration of a context, keeping its table and taxonomy (a DB is minimised with one
 * encoder and queried with another: bin/bonsai.cpp:152) */
int bns_b200_reconfigure(bns_b200_t *ctx, const bns_b200_config *cfg);

/* ---- taxonomy: khash_t(p) child -> parent ---------------------------------------------------------
 * build_parent_map, include/bonsai/util.h:766-785 (taxid 1 is forced to parent 0) */
int bns_b200_load_taxonomy(bns_b200_t *ctx, const uint32_t *child, const uint32_t *parent, uint64_t n);
int bns_b200_load_taxonomy_file(bns_b200_t *ctx, const char *nodes_dmp);
/* resolve_tree(counter, parent_map), util.h:831-869, over a batch of (taxid,count) lists: list r is
 * taxa[offsets[r]..offsets[r+1]) with u16 counts as linear::counter<tax_t,u16> keeps them */
int bns_b200_resolve_batch(bns_b200_t *ctx, const uint32_t *taxa, const uint16_t *counts, const uint64_t *offsets,
                           uint64_t n_lists, uint32_t *taxon_out);

/* ---- replication: one broadcast at load ------------------------------------------------
 * Rank 0 loads table + taxonomy and exports the 128-byte header; the host plumbing broadcasts it, non-root ranks
 * allocate from it (db_alloc_from_header), every rank lists its device segments (db_segments), the plumbing
 * broadcasts each segment (NCCL / torch.distributed) and non-root ranks call db_commit. */
typedef struct bns_b200_db_header { uint64_t words[16]; } bns_b200_db_header;
int bns_b200_db_export_header(const bns_b200_t *ctx, bns_b200_db_header *hdr);
int bns_b200_db_alloc_from_header(bns_b200_t *ctx, const bns_b200_db_header *hdr);
/* device segments that make up the database: fills up to `cap` (ptr, bytes) pairs (cap >= 5), returns the count in *n */
int bns_b200_db_segments(const bns_b200_t *ctx, void **dev_ptrs, uint64_t *bytes, int cap, int *n);
int bns_b200_db_commit(bns_b200_t *ctx);

/* ---- several GPUs in ONE process (reads dealt chunk by chunk to the GPUs, database replicated once) ------
 * open_multi opens one context per device (devices == NULL: ordinals 0..n_gpus-1; n_gpus == 0: cfg->n_gpus, and if that is 0
 * too every visible device) with the same Spacer / Encoder configuration; out[] receives n contexts and *n_out their count.
 * replicate copies the database (table + taxonomy) of handles[root] into the other contexts: ncclCommInitAll over their
 * devices and ONE ncclBroadcast per device segment (slots, value dictionary, val_info, node_info) over NVLink. After it
 * every context classifies on its own; each is driven by one host thread at a time (process_dataset, classifier.h:296,
 * deals its chunks round-robin). There is no per-batch collective. */
int bns_b200_open_multi(const bns_b200_config *cfg, int n_gpus, const int *devices, bns_b200_t **out, int *n_out);
int bns_b200_replicate(bns_b200_t *const *handles, int n, int root);
void bns_b200_close_multi(bns_b200_t **handles, int n);

/* ---- Encoder<Score>::for_each(fn, str, len) over a batch -- encoder.h:416 ------------------------------
 * bases: all sequences concatenated (ASCII); offsets[n+1]. Sequence r's k-mers are written in emission order
 * to kmers_out[out_offsets[r] ...], at most out_offsets[r+1]-out_offsets[r] of them (BNS_E_CAPACITY if more
 * were produced); counts_out[r] = number emitted. */
int bns_b200_encode_batch(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_seqs,
                          uint64_t *kmers_out, const uint64_t *out_offsets, uint32_t *counts_out);

/* ---- classify_seqs / classify_seq core -- classifier.h:213-238,269-289 ---------------------------------
 * One record per read (paired = 0) or per interleaved mate pair (paired = 1, classifier.h:233-236,257):
 *   taxon_out[r]  resolve_tree result (0 = unclassified)
 *   n_hit_out[r]  taxa.size()       n_missing_out[r]  missing_count     (either may be NULL)
 *   taxa_out      optional ordered per-k-mer hit taxids (std::vector<tax_t> taxa, classifier.h:228) of record
 *                 r at taxa_out[taxa_offsets[r] ...]; NULL to skip. */
int bns_b200_classify_batch(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_reads, int paired,
                            uint32_t *taxon_out, uint32_t *n_hit_out, uint32_t *n_missing_out,
                            uint32_t *taxa_out, const uint64_t *taxa_offsets);
/* As above plus mate1_kmers_out[r] (optional): the number of k-mers the FIRST mate of record r produced, which the
 * reference's first ambig_count term needs (classifier.h:232 is evaluated before the second mate is encoded). */
int bns_b200_classify_batch_ex(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_reads, int paired,
                               uint32_t *taxon_out, uint32_t *n_hit_out, uint32_t *n_missing_out,
                               uint32_t *taxa_out, const uint64_t *taxa_offsets, uint32_t *mate1_kmers_out);
/* The ordered hit list run-length encoded on the device (what the Kraken-style run lists print, classifier.h:46-61):
 * record r's runs are runs_out[run_pos_out[r] .. + n_runs_out[r]), each (taxid << 32 | run length), in k-mer order.
 * runs_cap entries are available in runs_out; *n_runs_total_out receives the entries used (the runs of a chunk may leave
 * small unused stretches between records; every record's runs are contiguous at run_pos_out[r]). BNS_E_CAPACITY if the runs
 * do not fit: nothing is counted (classified / unclassified stay as they were), *n_runs_total_out holds a lower bound on the
 * entries needed, and the call can be repeated with a larger buffer (bases + 2 entries per record + 2^21 always suffice).
 * 8 bytes per run cross PCIe instead of 4 per window slot. */
int bns_b200_classify_batch_runs(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_reads, int paired,
                                 uint32_t *taxon_out, uint32_t *n_hit_out, uint32_t *n_missing_out, uint32_t *mate1_kmers_out,
                                 uint64_t *runs_out, uint64_t runs_cap, uint64_t *run_pos_out, uint32_t *n_runs_out,
                                 uint64_t *n_runs_total_out);
/* Worker threads the host-buffer classify calls pack bases with (see bns_b200_config.host_pack_threads); 0 = every chunk
 * crosses PCIe as ASCII. The second call reports the current number. */
int bns_b200_set_host_pack_threads(bns_b200_t *ctx, uint32_t n_threads);
int bns_b200_host_pack_threads(const bns_b200_t *ctx);
/* Same with every buffer resident on the context's device; asynchronous on `stream` (a cudaStream_t). Window of record r in
 * d_taxa: at least (bases of the record) + 2 entries. Nothing is read back: conditions the host-pointer calls report as
 * BNS_E_CAPACITY (a record of 2^32-1 bases or more, more than 65536 distinct taxa in one record) are latched on the device
 * and returned -- and cleared -- by bns_b200_device_status(), which waits for the context's device to go idle. */
int bns_b200_classify_device(bns_b200_t *ctx, const char *d_bases, const uint64_t *d_offsets, uint64_t n_reads, int paired,
                             uint32_t *d_taxon, uint32_t *d_n_hit, uint32_t *d_n_missing,
                             uint32_t *d_taxa, const uint64_t *d_taxa_offsets, void *stream);

/* classify_device with the run lists of bns_b200_classify_batch_runs, every buffer on the device: d_runs (runs_cap entries; bases
 * + 2 per record + 2^21 always suffice), d_run_pos / d_n_runs per record, *d_runs_total = entries used (zeroed by the call).
 * For the encoders whose runs the lean kernel produces (every k-mer, no window: what `bonsai classify` runs); BNS_E_INVAL else. */
int bns_b200_classify_device_runs(bns_b200_t *ctx, const char *d_bases, const uint64_t *d_offsets, uint64_t n_reads, int paired,
                                  uint32_t *d_taxon, uint32_t *d_n_hit, uint32_t *d_n_missing, uint64_t *d_runs, uint64_t runs_cap,
                                  uint64_t *d_run_pos, uint32_t *d_n_runs, uint64_t *d_runs_total, void *stream);
int bns_b200_device_status(bns_b200_t *ctx);
int bns_b200_sync(bns_b200_t *ctx);
int bns_b200_stats_get(const bns_b200_t *ctx, bns_b200_stats *out);
int bns_b200_stats_reset(bns_b200_t *ctx);

/* pinned host buffers for ingest rings (kseq -> bseq1_t batches, include/bonsai/kseq_declare.h:112-145) */
int bns_b200_host_alloc(void **ptr, size_t bytes);
int bns_b200_host_free(void *ptr);

/* measurement helper: independent 32-byte loads at uniformly random buckets of the resident table;
 * returns device milliseconds for n_loads loads (the "random-gather ceiling" of SURVEY 8d) */
int bns_b200_bench_gather(bns_b200_t *ctx, uint64_t n_loads, uint64_t seed, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* BONSAI_B200_H */
