// TEST SCAFFOLD, never linked into the product: a dummy stand-in for the C ABI of include/bonsai_b200.h so that the CLI's HOST
// pipeline (ingest -> pinned batches -> text emitters -> ordered output, include/bonsai_b200/bonsai.hpp) can be exercised by the
// CPU test suite. It classifies nothing: "taxon", hit and missing counts and hit lists are an arbitrary deterministic function
// of each record's bases, which is all the test needs -- every way of reading the same records (kseq, parallel index of plain /
// gzip / BGZF files, any window and batch size, any -p) must produce byte-identical text. Built only by
// tests/test_cli_cpu.py::test_cli_host_pipeline_ingest_kinds, together with bonsai_b200/csrc/cli/bonsai_main.cpp, into a temp dir.
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../include/bonsai_b200.h"
struct bns_b200_ctx { uint64_t ncls = 0, nun = 0; uint32_t k = 31, canon = 0, api = 0; };
static uint64_t hsh(const char *p, size_t n) { uint64_t h = 1469598103934665603ull; for(size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)p[i]) * 1099511628211ull; return h; }
extern "C" {
const char *bns_b200_last_error(const bns_b200_t *) { return "stub"; }
int bns_b200_open(const bns_b200_config *cfg, bns_b200_t **out) { *out = new bns_b200_ctx; (*out)->k = cfg->k; (*out)->canon = cfg->canonicalize; (*out)->api = cfg->api; return 0; }
void bns_b200_close(bns_b200_t *c) { delete c; }
int bns_b200_load_table(bns_b200_t *, const uint64_t *, const uint32_t *, const uint32_t *, uint64_t) { return 0; }
int bns_b200_replicate(bns_b200_t *const *, int, int) { return 0; }        // every dummy context already "holds" the database
int bns_b200_load_taxonomy(bns_b200_t *, const uint32_t *, const uint32_t *, uint64_t) { return 0; }
int bns_b200_stats_get(const bns_b200_t *c, bns_b200_stats *s) { memset(s, 0, sizeof *s); s->n_classified = c->ncls; s->n_unclassified = c->nun; return 0; }
int bns_b200_host_alloc(void **p, size_t n) { *p = malloc(n); return *p ? 0 : -3; }
int bns_b200_host_free(void *p) { free(p); return 0; }
uint64_t bns_b200_encode_bound(const bns_b200_t *c, uint64_t len) { return len >= c->k ? len - c->k + 1 : 0; }
// not an encoder: one arbitrary word per position (a hash of the k bases there, the canonicalize flag and the API selector) for the
// record-overload configurations, so that the mirror's batching and context selection can be checked; the other selectors refuse
int bns_b200_encode_batch(bns_b200_t *c, const char *bases, const uint64_t *offs, uint64_t n, uint64_t *out, const uint64_t *ooffs, uint32_t *counts) {
    if(c->api != BNS_API_PATH) return -2;
    for(uint64_t r = 0; r < n; ++r) {
        const uint64_t len = offs[r + 1] - offs[r], m = len >= c->k ? len - c->k + 1 : 0;
        if(m > ooffs[r + 1] - ooffs[r]) return BNS_E_CAPACITY;
        for(uint64_t p = 0; p < m; ++p) out[ooffs[r] + p] = hsh(bases + offs[r] + p, c->k) * 2 + c->canon;
        counts[r] = (uint32_t)m;
    }
    return 0;
}
int bns_b200_build_begin(bns_b200_t *, uint64_t, const uint32_t *, uint32_t) { return -2; }
int bns_b200_build_add_genome(bns_b200_t *, const char *, const uint64_t *, uint64_t, uint32_t) { return -2; }
int bns_b200_build_finish(bns_b200_t *) { return -2; }
int bns_b200_table_dump(bns_b200_t *, uint64_t *, uint32_t *, uint64_t, uint64_t *) { return -2; }
// a deterministic function of each record's bases
static void one(bns_b200_t *c, const char *bases, const uint64_t *offs, uint64_t r, int paired, uint32_t *taxon, uint32_t *nhit, uint32_t *nmiss, uint32_t *m1,
                std::vector<uint32_t> &hits) {
    const int inc = paired ? 2 : 1; hits.clear();
    uint32_t miss = 0, mate1 = 0;
    uint64_t h = 0;
    for(int m = 0; m < inc; ++m) {
        const char *s = bases + offs[r * inc + m]; const size_t n = offs[r * inc + m + 1] - offs[r * inc + m];
        h = h * 31 + hsh(s, n);
        const size_t win = n >= c->k ? n - c->k + 1 : 0;
        for(size_t i = 0; i < win; ++i) { const uint64_t x = hsh(s + i, 8) % 7; if(x < 3) hits.push_back(x < 2 ? 11 : 12); else if(x < 6) ++miss; }
        if(m == 0) mate1 = (uint32_t)hits.size() + miss;
    }
    taxon[r] = hits.empty() ? 0 : (h % 5 == 0 ? 0 : (h % 2 ? 11 : 10));
    if(taxon[r]) ++c->ncls; else ++c->nun;
    nhit[r] = (uint32_t)hits.size(); nmiss[r] = miss; if(m1) m1[r] = mate1;
}
int bns_b200_classify_batch_ex(bns_b200_t *c, const char *bases, const uint64_t *offs, uint64_t n_reads, int paired, uint32_t *taxon, uint32_t *nhit, uint32_t *nmiss,
                               uint32_t *taxa, const uint64_t *taxa_offs, uint32_t *m1) {
    const uint64_t nrec = n_reads / (paired ? 2 : 1); std::vector<uint32_t> hits;
    for(uint64_t r = 0; r < nrec; ++r) { one(c, bases, offs, r, paired, taxon, nhit, nmiss, m1, hits); if(taxa) std::copy(hits.begin(), hits.end(), taxa + taxa_offs[r]); }
    return 0;
}
int bns_b200_classify_batch(bns_b200_t *c, const char *bases, const uint64_t *offs, uint64_t n_reads, int paired, uint32_t *taxon, uint32_t *nhit, uint32_t *nmiss,
                            uint32_t *taxa, const uint64_t *taxa_offs) { return bns_b200_classify_batch_ex(c, bases, offs, n_reads, paired, taxon, nhit, nmiss, taxa, taxa_offs, nullptr); }
int bns_b200_classify_batch_runs(bns_b200_t *c, const char *bases, const uint64_t *offs, uint64_t n_reads, int paired, uint32_t *taxon, uint32_t *nhit, uint32_t *nmiss, uint32_t *m1,
                                 uint64_t *runs, uint64_t cap, uint64_t *run_pos, uint32_t *nruns, uint64_t *total) {
    const uint64_t nrec = n_reads / (paired ? 2 : 1); std::vector<uint32_t> hits; uint64_t t = 0;
    const uint64_t c0 = c->ncls, u0 = c->nun;
    for(uint64_t r = 0; r < nrec; ++r) {
        one(c, bases, offs, r, paired, taxon, nhit, nmiss, m1, hits);
        run_pos[r] = t; uint32_t nr = 0;
        for(size_t i = 0; i < hits.size();) { size_t j = i; while(j < hits.size() && hits[j] == hits[i]) ++j;
            if(t >= cap) { c->ncls = c0; c->nun = u0; return BNS_E_CAPACITY; }
            runs[t++] = (uint64_t)hits[i] << 32 | (j - i); ++nr; i = j; }
        nruns[r] = nr;
    }
    *total = t; return 0;
}
}
