// bns_api.cu -- host side of libbonsai_b200.so: the C ABI of include/bonsai_b200.h.
//
// Owns one CUDA device per context: the bucketised k-mer table, the Euler-tour taxonomy arrays, a small
// ring of stream slots through which host batches are pipelined (H2D copy / kernel / D2H copy overlap
// across slots), and the counters ClassifierGeneric keeps (classifier.h:138,170-171).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include "../../include/bonsai_b200.h"
#include "bns_device.cuh"
#include "bns_host_util.h"
#include "bns_kernels.h"
#include "bns_pack.h"

using namespace bns;

namespace {

constexpr int N_SLOTS = 10;                 // stream slots; the packing call draws on all of them (small chunks, many in flight)
constexpr int N_RING = 3;                   // ... the other host-buffer calls pipeline their (larger) chunks over the first three
constexpr u64 PACK_CHUNK_READS = 1ull << 18;  // reads per chunk of the packing calls (both lanes draw from one chunk list)
constexpr u64 PACK_MIN_BASES = 16ull << 20;   // smaller batches are not worth waking the worker threads for
constexpr u64 CHUNK_BASES = 96ull << 20;      // bases per pipelined chunk
constexpr u64 CHUNK_READS = 1ull << 20;
constexpr double TARGET_LOAD_BIG = 1.75;
constexpr double TARGET_LOAD = 1.25;          // max entries per 4-slot bucket: second-sector probes stay ~1 % (2.2/bucket measured 14.6 %)

struct Slot {
    cudaStream_t st = nullptr;
    char *d_bases = nullptr;       size_t cap_bases = 0;
    u64 *d_offsets = nullptr;      size_t cap_offsets = 0;
    u32 *d_out = nullptr;          size_t cap_out = 0;       // taxon | nhit | nmiss (3 * reads)
    u32 *d_taxa = nullptr;         size_t cap_taxa = 0;
    u64 *d_taxa_offsets = nullptr; size_t cap_taxa_offsets = 0;
    u64 *d_kmers = nullptr;        size_t cap_kmers = 0;
    u64 *d_out_offsets = nullptr;  size_t cap_out_offsets = 0;
    u32 *d_defer = nullptr;        size_t cap_defer = 0;     // records the lean windowed kernel leaves to the generic one
    u64 *d_runs = nullptr;         size_t cap_runs = 0;      // run-length encoded hit lists (classify_batch_runs)
    u64 *d_run_pos = nullptr;      size_t cap_run_pos = 0;
    u32 *d_nruns = nullptr;        size_t cap_nruns = 0;
    unsigned long long *d_defer_cnt = nullptr;                // [0] records left to the second pass  [1] run-buffer entries handed out
    cudaEvent_t ka = nullptr, kb = nullptr;                   // around the kernels of the chunk in flight (stats.kernel_ms_total)
    bool k_timed = false;
    // host-packed chunks (bns_pack.h): pinned staging the worker threads pack into, and the device copies next to d_bases
    uint16_t *h_units = nullptr;   size_t cap_h_units = 0;
    u32 *h_susp = nullptr;         size_t cap_h_susp = 0;
    u64 *h_exc = nullptr;          size_t cap_h_exc = 0;
    u32 *d_susp = nullptr;         size_t cap_d_susp = 0;
    u64 *d_exc = nullptr;          size_t cap_d_exc = 0;
    cudaEvent_t h2d_done = nullptr;                           // behind the input copies of the chunk in flight
    bool busy = false, raw = false;
    bool reserved = false;                                    // being packed into / packed and not queued yet: the stream is idle, the staging is not
};

template <class T>
int ensure(T *&p, size_t &cap, size_t need) {
    if(need <= cap) return BNS_OK;
    if(p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = need + need / 4 + 256;
    if(cudaMalloc((void **)&p, want * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return BNS_E_NOMEM; }
    cap = want;
    return BNS_OK;
}
// chunk sizes in whole waves of the lean kernel (every warp of every SM one batch): no warp idles while others run a last batch
u64 whole_waves(u64 reads, int n_sm) {
    const u64 w = lean_wave_reads(n_sm);
    return reads >= w ? reads / w * w : reads;
}
template <class T>
int ensure_pinned(T *&p, size_t &cap, size_t need) {
    if(need <= cap) return BNS_OK;
    if(p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = need + need / 4 + 256;
    if(cudaHostAlloc((void **)&p, want * sizeof(T), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return BNS_E_NOMEM; }
    cap = want;
    return BNS_OK;
}

}  // namespace

struct bns_b200_ctx {
    int device = 0;
    int n_sm = 148;
    bns_b200_config cfg{};
    EncParams enc{};
    u32 c = 0, w = 0, W = 1;
    bool unspaced = true, unwindowed = true, canon = false;
    u32 ring_cap = 0;
    // table
    u64 *d_slots = nullptr;
    u64 n_buckets = 0, n_keys = 0, n_displaced = 0, n_overflowed = 0;
    u32 bucket_bits = 0, max_disp = 0, flag_count = 1;
    u32 layout = LAYOUT_HASH, fmt_bits = 0, table_k = 0, disp_bits = DISP_BITS;
    bool no_minimizer = false;        // a LAYOUT_MINIMIZER build of this key set failed (skewed minimizers): use LAYOUT_HASH
    // LAYOUT_MINIMIZER: keys whose chain was full (a few per thousand on real genomes) in a small LAYOUT_HASH table
    u64 *d_stash = nullptr; u32 stash_bits = 0, stash_flags = 1; u64 n_stash = 0;
    u64 *d_fail_k = nullptr; u32 *d_fail_v = nullptr;          // the insert kernel's list of keys that found no room
    std::vector<u32> values;          // sorted distinct DB values (value id -> taxid)
    u32 *d_values = nullptr;
    // taxonomy
    bool tax_loaded = false, tax_ready = false;
    std::vector<u32> tax_child, tax_parent;
    uint4 *d_val_info = nullptr, *d_node_info = nullptr;
    u32 n_nodes = 0, node_of_one = 0;
    // misc device state
    unsigned long long *d_counters = nullptr;   // [0] classified [1] unclassified [2..7] scratch [8..11] build stats
    bool building = false;
    bool timed = false;               // ev0/ev1 have been recorded
    u32 *d_status = nullptr;
    u32 *d_big = nullptr; size_t cap_big = 0;   // global-memory taxon lists of the second classify pass (databases of > 256 values)
    Slot slots[N_SLOTS];
    // host packing (bns_b200_config.host_pack_threads): worker threads, created at the first call large enough to use them
    int pack_threads = 0;             // 0 = the host-buffer calls ship ASCII
    int pack_mode = 0;                // 1 every chunk packed, 2 packed and ASCII chunks side by side (BNS_B200_HOST_PACK_MODE=pack|hybrid),
                                      // 0 (default) decided per call: 2 for pinned / registered host memory, 1 for pageable memory, whose
                                      // "asynchronous" copies are staged by the driver on the calling thread
    u64 pack_min_bases = PACK_MIN_BASES, pack_chunk_reads = PACK_CHUNK_READS;   // BNS_B200_PACK_MIN_BASES / _CHUNK_READS (tests use small ones)
    std::unique_ptr<PackPool> pool;
    std::vector<std::vector<uint64_t>> exc_parts[2];          // exception words per packing task, of the chunk being packed / being queued
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bns_b200_stats stats{};
    std::string err;

    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int cuda_fail(cudaError_t e, const char *what) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? BNS_E_NOMEM : BNS_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
};

#define CK(call)                                                          \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if(e__ != cudaSuccess) return ctx->cuda_fail(e__, #call);         \
    } while(0)

namespace {

thread_local std::string g_open_err;

// Spacer (spacer.h:59-71) + Encoder ctor (encoder.h:133-150) + the for_each dispatch (encoder.h:416-464)
int derive_encoder(bns_b200_ctx *ctx) {
    const bns_b200_config &cf = ctx->cfg;
    if(cf.k < 1 || cf.k > BNS_MAX_K) return ctx->fail(BNS_E_INVAL, "k must be in [1,32], got %u", cf.k);
    if(cf.score > BNS_SCORE_ENTROPY || cf.api > BNS_API_ITER || cf.entropy_cast > BNS_CAST_WRAP)
        return ctx->fail(BNS_E_INVAL, "bad score/api/entropy_cast selector");
    EncParams &P = ctx->enc;
    memset(&P, 0, sizeof P);
    const u32 k = cf.k;
    u32 c = k;
    bool unspaced = true;
    for(u32 i = 0; i + 1 < k; ++i) { c += cf.gaps[i]; if(cf.gaps[i]) unspaced = false; }
    if(c > (u32)CMAX) return ctx->fail(BNS_E_INVAL, "comb size %u exceeds the supported maximum %d", c, CMAX);
    const u32 w = std::max<int>((int)c, (int)cf.w);
    ctx->c = c; ctx->w = w; ctx->unspaced = unspaced; ctx->unwindowed = (k == w);
    ctx->canon = cf.canonicalize && (unspaced || cf.api == BNS_API_ITER);   // encoder.h:148-150 (the iterator calls do not look at it)
    ctx->W = w - c + 1;
    P.k = k; P.c = c; P.W = ctx->W;
    P.cast_wrap = cf.entropy_cast == BNS_CAST_WRAP;
    // contiguous runs of the comb
    {
        u32 pos = 0, run_start = 0, run_len = 1, ns = 0;
        for(u32 i = 0; i + 1 < k; ++i) {
            const u32 step = cf.gaps[i] + 1u;
            if(step == 1) ++run_len;
            else {
                P.seg_off[ns] = (uint16_t)run_start; P.seg_len[ns] = (uint16_t)run_len; ++ns;
                run_start = pos + step; run_len = 1;
            }
            pos += step;
        }
        P.seg_off[ns] = (uint16_t)run_start; P.seg_len[ns] = (uint16_t)run_len; ++ns;
        P.n_seg = ns;
    }
    // n/k * log(n/k) with the host's libm, written exactly as entropy.h:46-47 evaluates it
    {
        const double qi = 1. / k;
        for(u32 n = 1; n <= k; ++n) {
            volatile double p = (double)n * qi;
            volatile double l = std::log(p);
            P.plogp[n] = p * l;
        }
    }
    const bool ent = cf.score == BNS_SCORE_ENTROPY;
    const bool windowed = !ctx->unwindowed;
    P.score_kind = SC_LEX;
    if(cf.api == BNS_API_ITER) {
        // next_canonicalized_minimizer / next_minimizer call by call (encoder.h:616-628): kmer(pos) elements, every window
        // result kept
        P.family = FAM_K; P.canon_elem = ctx->canon ? 1 : 0; P.filter_none = 0; P.score_kind = ent ? SC_ENT_NOTFULL : SC_LEX;
    } else if(cf.api == BNS_API_STRING) {
        if(ctx->canon) {
            if(!windowed) { P.family = FAM_U; P.canon_elem = 1; }
            else if(ent) { P.family = FAM_R; P.score_kind = SC_ENT_ROLL; P.canon_emit = 1; P.tail_flush = 1; }
            else { P.family = FAM_K; P.canon_elem = 1; P.filter_none = 1; }
        } else if(unspaced) {
            if(!windowed) P.family = FAM_U;
            else { P.family = FAM_R; P.tail_flush = 1; P.score_kind = ent ? SC_ENT_ROLL : SC_LEX; }
        } else P.family = FAM_NONE;                                   // encoder.h:437-440
    } else {
        if(ctx->canon) {
            if(!windowed) { P.family = FAM_U; P.canon_elem = 1; }
            else { P.family = FAM_K; P.canon_elem = 1; P.filter_none = 1; P.score_kind = ent ? SC_ENT_NOTFULL : SC_LEX; }
        } else if(unspaced) {
            if(!windowed) P.family = FAM_U;
            else { P.family = FAM_R; P.tail_flush = 1; P.score_kind = ent ? SC_ENT_NOTFULL : SC_LEX; }
        } else { P.family = FAM_K; P.filter_none = 1; P.score_kind = ent ? SC_ENT_NOTFULL : SC_LEX; }
    }
    if(P.family == FAM_R) {
        P.filter_none = 1;                                            // `!= ENCODE_OVERFLOW`, encoder.h:299,337
        P.t_restart = (P.score_kind != SC_ENT_ROLL && k >= 31);       // encoder.h:283 (see stage_tile)
    }
    ctx->ring_cap = (ctx->W > 1 && P.family != FAM_U && P.family != FAM_NONE) ? (ctx->W - 1 + TILE) : 0;
    const size_t smem = stream_smem_bytes(ctx->ring_cap, true);
    if(smem > 200 * 1024) return ctx->fail(BNS_E_INVAL, "window %u needs %zu bytes of shared memory per CTA", w, smem);
    return BNS_OK;
}

int grid_for(bns_b200_ctx *ctx, u64 n_units, int occ) {
    const u64 want = (n_units + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const u64 cap = (u64)ctx->n_sm * (occ > 0 ? occ : 1);
    return (int)std::max<u64>(1, std::min(want, cap));
}

constexpr u64 FAIL_CAP = 1ull << 21;           // keys a minimizer-layout build may send to the stash

void free_table(bns_b200_ctx *ctx) {
    if(ctx->d_slots) cudaFree(ctx->d_slots);
    if(ctx->d_values) cudaFree(ctx->d_values);
    if(ctx->d_stash) cudaFree(ctx->d_stash);
    ctx->d_stash = nullptr; ctx->n_stash = 0; ctx->stash_bits = 0;
    ctx->d_slots = nullptr; ctx->d_values = nullptr;
    ctx->n_buckets = ctx->n_keys = 0; ctx->bucket_bits = 0;
    ctx->values.clear();
    ctx->tax_ready = false;
}

u32 bits_for(u64 n) { u32 b = 0; while((1ull << b) < n) ++b; return b; }

// LAYOUT_MINIMIZER (bns_device.cuh) is taken for tables far beyond L2 (>= 1 GiB: one DRAM line per lookup in the hash
// layout) of unspaced k-mers, 23 <= k <= 31, when the slot has room for the longer remainder (fmt_bits = loc_fmt_bits(k, b)
// in [value bits + displacement bits + 1, 28]). It suits key sets whose minimizers are spread out: a build that finds no
// room, or that had to displace more than a sixth of the keys (heavily repeated 16-mers: real genomes at small scale), is
// redone in the hash layout. BNS_B200_LAYOUT=hash|minimizer overrides the size rule and the displacement check.
// Measured: stress tables of 2^28 / 2^30 keys 283 -> 446 / 278 -> 397 Mreads/s at equal memory.
bool want_minimizer_layout(const bns_b200_ctx *ctx, u32 b) {
    const char *e = getenv("BNS_B200_LAYOUT");
    if(ctx->no_minimizer || (e && !strcmp(e, "hash"))) return false;
    const u32 k = ctx->cfg.k;
    if(k < 23 || k > 31 || b < 8 || b > 32) return false;
    if(b - LOC_GB > LOC_MB) return false;
    const int fmt = (int)loc_fmt_bits(k, b);
    const u32 vb = bits_for(std::max<u32>((u32)ctx->values.size(), 2));
    if(fmt > 28 || fmt < (int)(vb + DISP_BITS_LOC + 1)) return false;
    if(e && !strcmp(e, "minimizer")) return true;
    return ctx->unspaced && (32ull << b) >= (1ull << 30);
}
// a minimizer-layout build that displaced too many keys probes too long: better off in the hash layout
bool minimizer_build_too_crowded(const bns_b200_ctx *ctx) {
    const char *e = getenv("BNS_B200_LAYOUT");
    if(ctx->layout != LAYOUT_MINIMIZER || (e && !strcmp(e, "minimizer"))) return false;
    return ctx->n_displaced * 4 > ctx->n_keys;                    // measured: 10 % displaced is still 1.43x the hash layout at 2^30 keys
}
TableFmt stash_fmt(const bns_b200_ctx *ctx) {
    TableFmt f;
    f.b = ctx->stash_bits; f.fmt_bits = ctx->stash_bits; f.F = ctx->stash_flags; f.layout = LAYOUT_HASH; f.kt = ctx->table_k;
    f.disp_bits = DISP_BITS;
    return f;
}
TableFmt table_fmt(const bns_b200_ctx *ctx) {
    TableFmt f;
    f.b = ctx->bucket_bits; f.fmt_bits = ctx->fmt_bits; f.F = ctx->flag_count; f.layout = ctx->layout; f.kt = ctx->table_k;
    f.disp_bits = ctx->disp_bits;
    return f;
}
TableFmt fmt_for(const bns_b200_ctx *ctx, u32 b, bool minimizer) {
    TableFmt f;
    f.b = b; f.layout = minimizer ? LAYOUT_MINIMIZER : LAYOUT_HASH; f.kt = ctx->cfg.k;
    f.fmt_bits = minimizer ? loc_fmt_bits(ctx->cfg.k, b) : b;
    f.disp_bits = minimizer ? DISP_BITS_LOC : DISP_BITS;
    f.F = flag_count_for(f.fmt_bits, f.disp_bits, (u32)ctx->values.size());
    return f;
}
void adopt_fmt(bns_b200_ctx *ctx, const TableFmt &f) {
    ctx->bucket_bits = f.b; ctx->fmt_bits = f.fmt_bits; ctx->flag_count = f.F; ctx->layout = f.layout; ctx->table_k = f.kt;
    ctx->disp_bits = f.disp_bits;
    ctx->n_buckets = 1ull << f.b;
}

int alloc_table(bns_b200_ctx *ctx, u32 b, int force_layout = -1) {
    if(b > 32) return ctx->fail(BNS_E_NOMEM, "table would need 2^%u buckets", b);
    adopt_fmt(ctx, fmt_for(ctx, b, force_layout >= 0 ? force_layout == (int)LAYOUT_MINIMIZER : want_minimizer_layout(ctx, b)));
    CK(cudaMalloc((void **)&ctx->d_slots, ctx->n_buckets * 32));
    CK(cudaMemsetAsync(ctx->d_slots, 0xff, ctx->n_buckets * 32, ctx->slots[0].st));
    return BNS_OK;
}

u32 choose_bits(u64 n_keys, u32 n_values) {
    double small_load = TARGET_LOAD;
    if(const char *e = getenv("BNS_B200_TARGET_LOAD")) { const double v = atof(e); if(v > 0.05 && v <= 3.0) small_load = v; }   // experiments
    u32 b = bits_for((u64)std::ceil((double)std::max<u64>(n_keys, 1) / small_load));
    // Tables that cannot live in L2 anyway prefer density (more L2 hits, second probes mostly land in the same 128-byte
    // line): up to 1.75 entries per bucket there (measured on the 10.5 M-key DB: 512 Mreads/s at 1.25/bucket and 268 MB
    // against 400 at 0.63/bucket and 537 MB).
    double big_load = TARGET_LOAD_BIG;
    if(const char *e = getenv("BNS_B200_TARGET_LOAD_BIG")) { const double v = atof(e); if(v > 0.05 && v <= 3.5) big_load = v; }   // experiments
    if((32ull << b) > (64ull << 20)) b = bits_for((u64)std::ceil((double)std::max<u64>(n_keys, 1) / big_load));
    b = std::max(b, 12u);
    b = std::max(b, bits_for(std::max<u32>(n_values, 1)) + (u32)DISP_BITS + 1u);
    return b;
}

// tables that will take the minimizer layout: entries per 32-byte bucket from BNS_B200_LOC_LOAD when set (experiments)
u32 layout_bits(const bns_b200_ctx *ctx, u32 b, u64 n_keys) {
    if(const char *f = getenv("BNS_B200_LAYOUT"))                     // forced: small tables grow until the slot format has room
        if(!strcmp(f, "minimizer") && ctx->cfg.k >= 23 && ctx->cfg.k <= 31) {
            const u32 vb = bits_for(std::max<u32>((u32)ctx->values.size(), 2));
            while(b < 32 && (int)loc_fmt_bits(ctx->cfg.k, b) < (int)(vb + DISP_BITS_LOC + 1)) ++b;
        }
    const char *e = getenv("BNS_B200_LOC_LOAD");
    if(!e || !want_minimizer_layout(ctx, b)) return b;
    const double v = atof(e);
    if(!(v > 0.05 && v <= 3.0)) return b;
    return std::max(12u, bits_for((u64)std::ceil((double)std::max<u64>(n_keys, 1) / v)));
}

int upload_values(bns_b200_ctx *ctx) {
    const size_t n = std::max<size_t>(ctx->values.size(), 1);
    CK(cudaMalloc((void **)&ctx->d_values, n * sizeof(u32)));
    if(!ctx->values.empty())
        CK(cudaMemcpyAsync(ctx->d_values, ctx->values.data(), ctx->values.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->slots[0].st));
    return BNS_OK;
}

int refresh_table_stats(bns_b200_ctx *ctx) {
    cudaStream_t st = ctx->slots[0].st;
    CK(cudaMemsetAsync(ctx->d_counters + 2, 0, 3 * sizeof(unsigned long long), st));
    CK(launch_table_stats(st, ctx->d_slots, ctx->n_buckets, table_fmt(ctx), ctx->d_counters + 2));
    ++ctx->stats.kernel_launches;
    unsigned long long h[3];
    CK(cudaMemcpyAsync(h, ctx->d_counters + 2, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->n_keys = h[0] + ctx->n_stash; ctx->n_overflowed = h[1]; ctx->max_disp = (u32)h[2];
    return BNS_OK;
}

int ensure_fail_list(bns_b200_ctx *ctx);
// Insert host pairs through a pinned staging buffer; `next` fills up to cap pairs and returns how many.
template <class Next>
int insert_stream(bns_b200_ctx *ctx, Next next, unsigned long long *h_stats) {
    const size_t CH = 1u << 22;
    u64 *h_keys = nullptr, *d_keys = nullptr;
    u32 *h_vals = nullptr, *d_vals = nullptr;
    cudaStream_t st = ctx->slots[0].st;
    int rc = BNS_OK;
    if(ctx->layout == LAYOUT_MINIMIZER && (rc = ensure_fail_list(ctx)) != BNS_OK) return rc;
    if(cudaMallocHost((void **)&h_keys, 2 * CH * sizeof(u64)) != cudaSuccess ||
       cudaMallocHost((void **)&h_vals, 2 * CH * sizeof(u32)) != cudaSuccess ||
       cudaMalloc((void **)&d_keys, 2 * CH * sizeof(u64)) != cudaSuccess ||
       cudaMalloc((void **)&d_vals, 2 * CH * sizeof(u32)) != cudaSuccess) {
        rc = ctx->cuda_fail(cudaGetLastError(), "staging allocation");
    } else {
        cudaEvent_t done[2];
        cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming);
        cudaMemsetAsync(ctx->d_counters + 5, 0, 3 * sizeof(unsigned long long), st);
        for(int buf = 0;; buf ^= 1) {
            cudaEventSynchronize(done[buf]);
            const size_t n = next(h_keys + buf * CH, h_vals + buf * CH, CH);
            if(!n) break;
            cudaMemcpyAsync(d_keys + buf * CH, h_keys + buf * CH, n * sizeof(u64), cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(d_vals + buf * CH, h_vals + buf * CH, n * sizeof(u32), cudaMemcpyHostToDevice, st);
            cudaError_t e = launch_insert(st, ctx->d_slots, table_fmt(ctx), d_keys + buf * CH, d_vals + buf * CH, n,
                                          ctx->d_values, (u32)ctx->values.size(), ctx->d_counters + 5,
                                          ctx->layout == LAYOUT_MINIMIZER ? ctx->d_fail_k : nullptr, ctx->d_fail_v, FAIL_CAP);
            ++ctx->stats.kernel_launches;
            ctx->stats.h2d_bytes += n * 12;
            if(e != cudaSuccess) { rc = ctx->cuda_fail(e, "insert kernel"); break; }
            cudaEventRecord(done[buf], st);
        }
        cudaMemcpyAsync(h_stats, ctx->d_counters + 5, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if(rc == BNS_OK && e != cudaSuccess) rc = ctx->cuda_fail(e, "table build");
        cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
    }
    if(h_keys) cudaFreeHost(h_keys);
    if(h_vals) cudaFreeHost(h_vals);
    if(d_keys) cudaFree(d_keys);
    if(d_vals) cudaFree(d_vals);
    return rc;
}

// Taxonomy -> Euler-tour arrays. Node 0 is "none". Lenient where the reference is UB (SURVEY B-9): a parent
// that is not itself a node is treated as 0, a DB value that is not a node becomes an isolated root.
int finalize_taxonomy(bns_b200_ctx *ctx) {
    if(ctx->tax_ready) return BNS_OK;
    if(!ctx->tax_loaded) return ctx->fail(BNS_E_STATE, "taxonomy not loaded");
    std::vector<u32> ids(ctx->tax_child);
    ids.push_back(1);
    ids.insert(ids.end(), ctx->values.begin(), ctx->values.end());
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    if(!ids.empty() && ids[0] == 0) ids.erase(ids.begin());          // taxid 0 is "no taxon"
    const u32 n = (u32)ids.size();
    auto node_of = [&](u32 taxid) -> u32 {
        auto it = std::lower_bound(ids.begin(), ids.end(), taxid);
        return (it != ids.end() && *it == taxid) ? (u32)(it - ids.begin()) + 1 : 0;
    };
    std::vector<u32> parent(n + 1, 0);
    for(size_t i = 0; i < ctx->tax_child.size(); ++i) {              // later lines overwrite earlier ones (kh_put + assign)
        const u32 nd = node_of(ctx->tax_child[i]);
        if(nd) parent[nd] = node_of(ctx->tax_parent[i]);
    }
    const u32 one = node_of(1);
    parent[one] = 0;                                                  // util.h:780-781
    // children CSR
    std::vector<u32> deg(n + 2, 0), start(n + 2, 0), kids(n ? n : 1);
    for(u32 v = 1; v <= n; ++v) ++deg[parent[v]];
    for(u32 v = 0; v <= n; ++v) start[v + 1] = start[v] + deg[v];
    {
        std::vector<u32> fill(start.begin(), start.end() - 1);
        for(u32 v = 1; v <= n; ++v) kids[fill[parent[v]]++] = v;
    }
    std::vector<u32> tin(n + 1, 0), tout(n + 1, 0), it(n + 1, 0), stack;
    u32 clock = 0, visited = 0;
    for(u32 ri = start[0]; ri < start[1]; ++ri) {                     // every root (parent 0)
        stack.push_back(kids[ri]);
        tin[kids[ri]] = clock++; ++visited;
        while(!stack.empty()) {
            const u32 v = stack.back();
            if(it[v] < deg[v]) {
                const u32 ch = kids[start[v] + it[v]++];
                tin[ch] = clock++; ++visited;
                stack.push_back(ch);
            } else { tout[v] = clock; stack.pop_back(); }
        }
    }
    if(visited != n) return ctx->fail(BNS_E_TAXONOMY, "taxonomy has a cycle: %u of %u nodes are not reachable from a root", n - visited, n);
    std::vector<uint4> node_info(n + 1), val_info(std::max<size_t>(ctx->values.size(), 1));
    node_info[0] = make_uint4(0xffffffffu, 0, 0, 0);
    for(u32 v = 1; v <= n; ++v) node_info[v] = make_uint4(tin[v], tout[v], parent[v], ids[v - 1]);
    for(size_t i = 0; i < ctx->values.size(); ++i) {
        const u32 nd = node_of(ctx->values[i]);
        // value 0 never scores (resolve_tree's `while(node)`): an empty interval
        val_info[i] = nd ? make_uint4(tin[nd], tout[nd], nd, ctx->values[i]) : make_uint4(0xffffffffu, 0, 0, 0);
    }
    if(ctx->d_node_info) cudaFree(ctx->d_node_info);
    if(ctx->d_val_info) cudaFree(ctx->d_val_info);
    ctx->d_node_info = ctx->d_val_info = nullptr;
    CK(cudaMalloc((void **)&ctx->d_node_info, node_info.size() * sizeof(uint4)));
    CK(cudaMalloc((void **)&ctx->d_val_info, val_info.size() * sizeof(uint4)));
    CK(cudaMemcpy(ctx->d_node_info, node_info.data(), node_info.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_val_info, val_info.data(), val_info.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    ctx->n_nodes = n + 1;
    ctx->node_of_one = one;
    ctx->tax_ready = true;
    return BNS_OK;
}

TableView table_view(const bns_b200_ctx *ctx) {
    TableView T;
    T.slots = ctx->d_slots;
    T.bucket_bits = ctx->bucket_bits;
    T.fmt = table_fmt(ctx);
    T.tag_shift = ctx->fmt_bits - ctx->disp_bits;
    T.flag_shift = T.tag_shift - ctx->flag_count;
    T.flag_mask = ctx->flag_count - 1;
    T.val_mask = (1u << T.flag_shift) - 1;
    T.n_values = (u32)ctx->values.size();
    T.stash = ctx->d_stash;
    T.sfmt = stash_fmt(ctx);
    return T;
}
TaxView tax_view(const bns_b200_ctx *ctx) {
    TaxView X;
    X.val_info = ctx->d_val_info;
    X.node_info = ctx->d_node_info;
    X.n_nodes = ctx->n_nodes;
    X.node_of_one = ctx->node_of_one;
    return X;
}

// the insert kernels listed the keys that found no room (d_fail_k / d_fail_v): a minimizer-layout table keeps them in a small
// LAYOUT_HASH stash when they are few (at most 1/16 of the keys); st[0] becomes 0 when that worked
int ensure_fail_list(bns_b200_ctx *ctx) {
    if(ctx->d_fail_k) return BNS_OK;
    CK(cudaMalloc((void **)&ctx->d_fail_k, FAIL_CAP * sizeof(u64)));
    CK(cudaMalloc((void **)&ctx->d_fail_v, FAIL_CAP * sizeof(u32)));
    return BNS_OK;
}
int settle_failures(bns_b200_ctx *ctx, unsigned long long *st, u64 n_keys, int layout = -1) {
    if(ctx->d_stash) { cudaFree(ctx->d_stash); ctx->d_stash = nullptr; }
    ctx->n_stash = 0; ctx->stash_bits = 0;
    if(layout < 0) layout = (int)ctx->layout;
    if(!st[0] || layout != (int)LAYOUT_MINIMIZER || !ctx->d_fail_k || st[0] > FAIL_CAP || st[0] * 16 > n_keys) return BNS_OK;
    cudaStream_t s0 = ctx->slots[0].st;
    const u64 nf = st[0];
    for(u32 bs = std::max(choose_bits(nf, (u32)ctx->values.size()), 10u); bs <= 30; ++bs) {
        ctx->stash_bits = bs;
        ctx->stash_flags = flag_count_for(bs, DISP_BITS, (u32)ctx->values.size());
        CK(cudaMalloc((void **)&ctx->d_stash, (32ull << bs)));
        CK(cudaMemsetAsync(ctx->d_stash, 0xff, (32ull << bs), s0));
        CK(cudaMemsetAsync(ctx->d_counters + 13, 0, 3 * sizeof(unsigned long long), s0));
        CK(launch_insert(s0, ctx->d_stash, stash_fmt(ctx), ctx->d_fail_k, ctx->d_fail_v, nf, ctx->d_values, (u32)ctx->values.size(), ctx->d_counters + 13));
        ++ctx->stats.kernel_launches;
        unsigned long long h[3];
        CK(cudaMemcpyAsync(h, ctx->d_counters + 13, sizeof h, cudaMemcpyDeviceToHost, s0));
        CK(cudaStreamSynchronize(s0));
        if(h[0] == 0 && h[2] == 0) { ctx->n_stash = nf; st[0] = 0; return BNS_OK; }
        cudaFree(ctx->d_stash); ctx->d_stash = nullptr;
    }
    ctx->stash_bits = 0;
    return BNS_OK;
}

int finish_table(bns_b200_ctx *ctx, const unsigned long long *h_stats) {
    if(h_stats[2]) return ctx->fail(BNS_E_INVAL, "%llu values were not in the value dictionary", h_stats[2]);
    ctx->n_displaced = h_stats[1];
    int rc = refresh_table_stats(ctx);
    ctx->tax_ready = false;
    return rc;
}

// device time of the kernels a slot ran for its last chunk (the slot's stream must be idle)
void collect_kernel_time(bns_b200_ctx *ctx, Slot &s) {
    if(!s.k_timed) return;
    s.k_timed = false;
    float ms = 0.f;
    if(cudaEventElapsedTime(&ms, s.ka, s.kb) == cudaSuccess) { ctx->stats.kernel_ms_total += ms; ctx->stats.last_kernel_ms = ms; }
    else cudaGetLastError();
}

int check_status(bns_b200_ctx *ctx, u32 st) {
    if(st & 2u) return ctx->fail(BNS_E_CAPACITY, "a record hit more than 65536 distinct taxa");
    if(st & 1u) return ctx->fail(BNS_E_CAPACITY, "an output window was too small for the k-mers produced");
    if(st & 4u) return ctx->fail(BNS_E_INVAL, "a taxid passed to resolve is not a database value");
    if(st & 8u) return ctx->fail(BNS_E_CAPACITY, "a record is longer than 2^32-2 bases");
    return BNS_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

const char *bns_b200_version(void) { return "bonsai_b200 0.1 (sm_100a), ABI 1"; }

const char *bns_b200_strerror(int code) {
    switch(code) {
        case BNS_OK: return "ok";
        case BNS_E_INVAL: return "invalid argument";
        case BNS_E_CUDA: return "CUDA error";
        case BNS_E_NOMEM: return "out of memory";
        case BNS_E_STATE: return "table or taxonomy not loaded";
        case BNS_E_TAXONOMY: return "malformed taxonomy";
        case BNS_E_CAPACITY: return "output capacity exceeded";
        case BNS_E_IO: return "I/O error";
        default: return "unknown error";
    }
}
const char *bns_b200_last_error(const bns_b200_t *ctx) { return ctx ? ctx->err.c_str() : g_open_err.c_str(); }

int bns_b200_open(const bns_b200_config *cfg, bns_b200_t **out) {
    if(!cfg || !out) return BNS_E_INVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if(e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_open_err = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return BNS_E_CUDA;                              // no CPU fallback, by design
    }
    bns_b200_ctx *ctx = new bns_b200_ctx();
    ctx->cfg = *cfg;
    int rc = derive_encoder(ctx);
    if(rc != BNS_OK) { g_open_err = ctx->err; delete ctx; return rc; }
    int dev = cfg->device;
    if(dev < 0) cudaGetDevice(&dev);
    if(dev >= ndev) { g_open_err = "device ordinal out of range"; delete ctx; return BNS_E_INVAL; }
    ctx->device = dev;
    auto bail = [&](cudaError_t er, const char *what) {
        g_open_err = std::string(what) + ": " + cudaGetErrorString(er);
        cudaGetLastError();
        bns_b200_close(ctx);
        return BNS_E_CUDA;
    };
    if((e = cudaSetDevice(dev)) != cudaSuccess) return bail(e, "cudaSetDevice");
    // Optional knob: BNS_B200_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity. Measured on B200 it changes nothing:
    // a sector miss fills its whole 128-byte line whatever the limit (profiles/micro/gather_variants_r01.txt).
    if(const char *env = getenv("BNS_B200_L2_FETCH")) {
        const long g = atol(env);
        if(g == 32 || g == 64 || g == 128) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g); cudaGetLastError(); }
    }
    cudaDeviceProp prop;
    if((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    ctx->n_sm = prop.multiProcessorCount;
    for(int i = 0; i < N_SLOTS; ++i)
        if((e = cudaStreamCreateWithFlags(&ctx->slots[i].st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for(int i = 0; i < N_SLOTS; ++i) {
        if((e = cudaMalloc((void **)&ctx->slots[i].d_defer_cnt, 2 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
        cudaMemset(ctx->slots[i].d_defer_cnt, 0, 2 * sizeof(unsigned long long));
        if((e = cudaEventCreate(&ctx->slots[i].ka)) != cudaSuccess || (e = cudaEventCreate(&ctx->slots[i].kb)) != cudaSuccess) return bail(e, "cudaEventCreate");
        if((e = cudaEventCreateWithFlags(&ctx->slots[i].h2d_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    }
    // Host packing: cfg.host_pack_threads = N worker threads, 0 = this process's share of the machine's threads (divided by
    // LOCAL_WORLD_SIZE under torchrun) less the calling one, 0xffffffff = off; BNS_B200_HOST_PACK=N|0 overrides.
    {
        long want = cfg->host_pack_threads == 0xffffffffu ? 0 : (long)cfg->host_pack_threads;
        bool given = cfg->host_pack_threads != 0;
        if(const char *env = getenv("BNS_B200_HOST_PACK")) { want = atol(env); given = true; }
        if(!given) {
            long share = (long)std::thread::hardware_concurrency();
            if(const char *lw = getenv("LOCAL_WORLD_SIZE")) { const long n = atol(lw); if(n > 1) share /= n; }
            want = share - 1;
            if(want < 2) want = 0;                                     // one helper does not beat the copy engine
        }
        ctx->pack_threads = (int)std::max(0l, std::min(want, 64l));
        if(const char *env = getenv("BNS_B200_HOST_PACK_MODE")) ctx->pack_mode = !strcmp(env, "pack") ? 1 : 2;
        if(const char *env = getenv("BNS_B200_PACK_MIN_BASES")) ctx->pack_min_bases = strtoull(env, nullptr, 10);
        if(const char *env = getenv("BNS_B200_PACK_CHUNK_READS")) ctx->pack_chunk_reads = std::max<u64>(64, strtoull(env, nullptr, 10));
    }
    if((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if((e = cudaMalloc((void **)&ctx->d_counters, 16 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    if((e = cudaMalloc((void **)&ctx->d_status, sizeof(u32))) != cudaSuccess) return bail(e, "cudaMalloc");
    cudaMemset(ctx->d_counters, 0, 16 * sizeof(unsigned long long));
    cudaMemset(ctx->d_status, 0, sizeof(u32));
    *out = ctx;
    return BNS_OK;
}

void bns_b200_close(bns_b200_t *ctx) {
    if(!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_table(ctx);
    if(ctx->d_val_info) cudaFree(ctx->d_val_info);
    if(ctx->d_node_info) cudaFree(ctx->d_node_info);
    if(ctx->d_counters) cudaFree(ctx->d_counters);
    if(ctx->d_status) cudaFree(ctx->d_status);
    if(ctx->d_big) cudaFree(ctx->d_big);
    for(auto &s : ctx->slots) {
        if(s.d_bases) cudaFree(s.d_bases);
        if(s.d_offsets) cudaFree(s.d_offsets);
        if(s.d_out) cudaFree(s.d_out);
        if(s.d_taxa) cudaFree(s.d_taxa);
        if(s.d_taxa_offsets) cudaFree(s.d_taxa_offsets);
        if(s.d_kmers) cudaFree(s.d_kmers);
        if(s.d_out_offsets) cudaFree(s.d_out_offsets);
        if(s.d_defer) cudaFree(s.d_defer);
        if(s.d_runs) cudaFree(s.d_runs);
        if(s.d_run_pos) cudaFree(s.d_run_pos);
        if(s.d_nruns) cudaFree(s.d_nruns);
        if(s.d_defer_cnt) cudaFree(s.d_defer_cnt);
        if(s.d_susp) cudaFree(s.d_susp);
        if(s.d_exc) cudaFree(s.d_exc);
        if(s.h_units) cudaFreeHost(s.h_units);
        if(s.h_susp) cudaFreeHost(s.h_susp);
        if(s.h_exc) cudaFreeHost(s.h_exc);
        if(s.h2d_done) cudaEventDestroy(s.h2d_done);
        if(s.ka) cudaEventDestroy(s.ka);
        if(s.kb) cudaEventDestroy(s.kb);
        if(s.st) cudaStreamDestroy(s.st);
    }
    if(ctx->ev0) cudaEventDestroy(ctx->ev0);
    if(ctx->ev1) cudaEventDestroy(ctx->ev1);
    cudaGetLastError();
    delete ctx;
}

int bns_b200_geometry(const bns_b200_t *ctx, uint32_t *c, uint32_t *w, int *unspaced, int *unwindowed, int *canon) {
    if(!ctx) return BNS_E_INVAL;
    if(c) *c = ctx->c;
    if(w) *w = ctx->w;
    if(unspaced) *unspaced = ctx->unspaced;
    if(unwindowed) *unwindowed = ctx->unwindowed;
    if(canon) *canon = ctx->canon;
    return BNS_OK;
}

uint64_t bns_b200_encode_bound(const bns_b200_t *ctx, uint64_t len) {
    if(!ctx || len < ctx->c) return 0;
    return len - ctx->c + 1;         // one element per position at most; the tail flush only fires when nothing else was emitted
}

// ---- table --------------------------------------------------------------------------------------
int bns_b200_load_table(bns_b200_t *ctx, const uint64_t *keys, const uint32_t *vals, const uint32_t *flags, uint64_t n_buckets) {
    if(!ctx || (n_buckets && (!keys || !vals || !flags))) return ctx ? ctx->fail(BNS_E_INVAL, "null khash arrays") : BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    ctx->no_minimizer = false;
    auto occupied = [&](u64 i) { return ((flags[i >> 4] >> ((i & 0xfu) << 1)) & 3u) == 0; };   // !__ac_iseither, khash64.h:171
    uint64_t n_occupied = 0;
    const std::vector<u32> values = distinct_values_khash(vals, flags, n_buckets, &n_occupied);   // the value dictionary (bns_host_util.h)
    const u64 n_keys = n_occupied;
    ctx->values = values;
    for(u32 b = layout_bits(ctx, choose_bits(n_keys, (u32)values.size()), n_keys);; ++b) {
        free_table(ctx);
        ctx->values = values;
        int rc = upload_values(ctx);
        if(rc == BNS_OK) rc = alloc_table(ctx, b);
        if(rc != BNS_OK) return rc;
        u64 cursor = 0;
        unsigned long long st[3] = {0, 0, 0};
        rc = insert_stream(ctx, [&](u64 *hk, u32 *hv, size_t cap) {
            size_t n = 0;
            while(cursor < n_buckets && n < cap) {
                if(occupied(cursor)) { hk[n] = keys[cursor]; hv[n] = vals[cursor]; ++n; }
                ++cursor;
            }
            return n;
        }, st);
        if(rc != BNS_OK) return rc;
        if((rc = settle_failures(ctx, st, n_keys)) != BNS_OK) return rc;
        if(st[0] == 0) {
            rc = finish_table(ctx, st);
            if(rc != BNS_OK || !minimizer_build_too_crowded(ctx)) return rc;
            ctx->no_minimizer = true; --b;                            // same size, hash layout
            continue;
        }
        // some key found no room within MAX_DISP buckets of home: grow and rebuild (a minimizer-layout build that fails --
        // skewed minimizers beyond what the stash takes -- is redone in the hash layout at the same size first)
        if(ctx->layout == LAYOUT_MINIMIZER) { ctx->no_minimizer = true; --b; }
    }
}

int bns_b200_load_pairs(bns_b200_t *ctx, const uint64_t *keys, const uint32_t *vals, uint64_t n) {
    if(!ctx || (n && (!keys || !vals))) return ctx ? ctx->fail(BNS_E_INVAL, "null key/value arrays") : BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    ctx->no_minimizer = false;
    const std::vector<u32> values = distinct_values(vals, n);
    ctx->values = values;
    for(u32 b = layout_bits(ctx, choose_bits(n, (u32)values.size()), n);; ++b) {
        free_table(ctx);
        ctx->values = values;
        int rc = upload_values(ctx);
        if(rc == BNS_OK) rc = alloc_table(ctx, b);
        if(rc != BNS_OK) return rc;
        u64 cursor = 0;
        unsigned long long st[3] = {0, 0, 0};
        rc = insert_stream(ctx, [&](u64 *hk, u32 *hv, size_t cap) {
            const size_t m = (size_t)std::min<u64>(cap, n - cursor);
            memcpy(hk, keys + cursor, m * sizeof(u64));
            memcpy(hv, vals + cursor, m * sizeof(u32));
            cursor += m;
            return m;
        }, st);
        if(rc != BNS_OK) return rc;
        if((rc = settle_failures(ctx, st, n)) != BNS_OK) return rc;
        if(st[0] == 0) {
            rc = finish_table(ctx, st);
            if(rc != BNS_OK || !minimizer_build_too_crowded(ctx)) return rc;
            ctx->no_minimizer = true; --b;                            // same size, hash layout
            continue;
        }
        if(ctx->layout == LAYOUT_MINIMIZER) { ctx->no_minimizer = true; --b; }
    }
}

int bns_b200_load_pairs_device(bns_b200_t *ctx, const uint64_t *d_keys, const uint32_t *d_vals, uint64_t n,
                               const uint32_t *values, uint32_t n_values) {
    if(!ctx || !values || !n_values || (n && (!d_keys || !d_vals))) return ctx ? ctx->fail(BNS_E_INVAL, "bad arguments") : BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    ctx->no_minimizer = false;
    std::vector<u32> vs(values, values + n_values);
    std::sort(vs.begin(), vs.end());
    vs.erase(std::unique(vs.begin(), vs.end()), vs.end());
    cudaStream_t st = ctx->slots[0].st;
    ctx->values = vs;
    for(u32 b = layout_bits(ctx, choose_bits(n, (u32)vs.size()), n);; ++b) {
        free_table(ctx);
        ctx->values = vs;
        int rc = upload_values(ctx);
        if(rc == BNS_OK) rc = alloc_table(ctx, b);
        if(rc != BNS_OK) return rc;
        CK(cudaMemsetAsync(ctx->d_counters + 5, 0, 3 * sizeof(unsigned long long), st));
        if(ctx->layout == LAYOUT_MINIMIZER && (rc = ensure_fail_list(ctx)) != BNS_OK) return rc;
        const u64 CH = 1ull << 28;
        for(u64 off = 0; off < n; off += CH) {
            CK(launch_insert(st, ctx->d_slots, table_fmt(ctx), (const u64 *)d_keys + off, d_vals + off, std::min(CH, n - off),
                             ctx->d_values, (u32)ctx->values.size(), ctx->d_counters + 5,
                             ctx->layout == LAYOUT_MINIMIZER ? ctx->d_fail_k : nullptr, ctx->d_fail_v, FAIL_CAP));
            ++ctx->stats.kernel_launches;
        }
        unsigned long long h[3];
        CK(cudaMemcpyAsync(h, ctx->d_counters + 5, sizeof h, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if(getenv("BNS_B200_VERBOSE"))
            fprintf(stderr, "[bns_b200] table build: 2^%u buckets, layout %u, no room for %llu keys, %llu displaced of %llu\n", b, ctx->layout, h[0], h[1],
                    (unsigned long long)n);
        if((rc = settle_failures(ctx, h, n)) != BNS_OK) return rc;
        if(h[0] == 0) {
            rc = finish_table(ctx, h);
            if(rc != BNS_OK || !minimizer_build_too_crowded(ctx)) return rc;
            ctx->no_minimizer = true; --b;                            // same size, hash layout
            continue;
        }
        if(ctx->layout == LAYOUT_MINIMIZER) { ctx->no_minimizer = true; --b; }
    }
}

int bns_b200_table_info_get(const bns_b200_t *ctx, bns_b200_table_info *info) {
    if(!ctx || !info) return BNS_E_INVAL;
    memset(info, 0, sizeof *info);
    info->n_keys = ctx->n_keys;
    info->n_buckets = ctx->n_buckets;
    info->bytes = ctx->n_buckets * 32;
    info->bucket_bits = ctx->bucket_bits;
    info->val_bits = ctx->bucket_bits ? ctx->fmt_bits - ctx->disp_bits - ctx->flag_count : 0;
    info->layout = ctx->layout;
    info->disp_bits = ctx->disp_bits;
    info->n_values = (u32)ctx->values.size();
    info->max_disp = ctx->max_disp;
    info->n_displaced = ctx->n_displaced;
    info->n_overflowed = ctx->n_overflowed;
    info->n_stash = ctx->n_stash;
    return BNS_OK;
}

int bns_b200_lookup_batch(bns_b200_t *ctx, const uint64_t *keys, uint64_t n, uint32_t *vals_out, uint8_t *found_out) {
    if(!ctx || (n && (!keys || !vals_out || !found_out))) return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    if(!ctx->d_slots) return ctx->fail(BNS_E_STATE, "no table loaded");
    CK(cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[0];
    const u64 CH = 1ull << 24;
    for(u64 off = 0; off < n; off += CH) {
        const u64 m = std::min(CH, n - off);
        int rc = ensure(s.d_kmers, s.cap_kmers, m);
        if(rc == BNS_OK) rc = ensure(s.d_out, s.cap_out, m + (m + 3) / 4);
        if(rc != BNS_OK) return ctx->fail(rc, "device workspace");
        uint8_t *d_found = (uint8_t *)(s.d_out + m);
        CK(cudaMemcpyAsync(s.d_kmers, keys + off, m * sizeof(u64), cudaMemcpyHostToDevice, s.st));
        CK(launch_lookup(s.st, table_view(ctx), ctx->d_values, s.d_kmers, m, s.d_out, d_found));
        ++ctx->stats.kernel_launches;
        CK(cudaMemcpyAsync(vals_out + off, s.d_out, m * sizeof(u32), cudaMemcpyDeviceToHost, s.st));
        CK(cudaMemcpyAsync(found_out + off, d_found, m, cudaMemcpyDeviceToHost, s.st));
        CK(cudaStreamSynchronize(s.st));
    }
    return BNS_OK;
}

int bns_b200_lookup_sectors(bns_b200_t *ctx, const uint64_t *keys, uint64_t n, uint64_t *sectors_out) {
    if(!ctx || !sectors_out || (n && !keys)) return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    if(!ctx->d_slots) return ctx->fail(BNS_E_STATE, "no table loaded");
    CK(cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[0];
    CK(cudaMemsetAsync(ctx->d_counters + 4, 0, sizeof(unsigned long long), s.st));
    const u64 CH = 1ull << 24;
    for(u64 off = 0; off < n; off += CH) {
        const u64 m = std::min(CH, n - off);
        int rc = ensure(s.d_kmers, s.cap_kmers, m);
        if(rc != BNS_OK) return ctx->fail(rc, "device workspace");
        CK(cudaMemcpyAsync(s.d_kmers, keys + off, m * sizeof(u64), cudaMemcpyHostToDevice, s.st));
        CK(launch_sectors(s.st, table_view(ctx), s.d_kmers, m, ctx->d_counters + 4));
        ++ctx->stats.kernel_launches;
        CK(cudaStreamSynchronize(s.st));
    }
    unsigned long long h = 0;
    CK(cudaMemcpy(&h, ctx->d_counters + 4, sizeof h, cudaMemcpyDeviceToHost));
    *sectors_out = h;
    return BNS_OK;
}

// ---- database construction on the device -----------------------------------------------------------------
int bns_b200_reconfigure(bns_b200_t *ctx, const bns_b200_config *cfg) {
    if(!ctx || !cfg) return BNS_E_INVAL;
    const bns_b200_config old = ctx->cfg;
    ctx->cfg = *cfg;
    ctx->cfg.device = old.device;
    const int rc = derive_encoder(ctx);
    if(rc != BNS_OK) { ctx->cfg = old; derive_encoder(ctx); }
    return rc;
}

int bns_b200_build_begin(bns_b200_t *ctx, uint64_t max_kmers, const uint32_t *taxids, uint32_t n_taxids) {
    if(!ctx || !taxids || !n_taxids) return ctx ? ctx->fail(BNS_E_INVAL, "no taxids") : BNS_E_INVAL;
    if(!ctx->tax_loaded) return ctx->fail(BNS_E_STATE, "load the taxonomy before building");
    if(ctx->cfg.api == BNS_API_ITER) return ctx->fail(BNS_E_INVAL, "BNS_API_ITER is an encode-only configuration");
    CK(cudaSetDevice(ctx->device));
    // value dictionary = the taxids and all their ancestors (every lca result lives there), plus taxid 1
    std::vector<u32> order(ctx->tax_child.size());
    for(size_t i = 0; i < order.size(); ++i) order[i] = (u32)i;
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return ctx->tax_child[a] < ctx->tax_child[b]; });
    auto parent_of = [&](u32 t, bool &found) -> u32 {
        if(t == 1) { found = true; return 0; }
        // last line for a taxid wins, as repeated kh_put + assignment does (util.h:773-776)
        auto it = std::upper_bound(order.begin(), order.end(), t, [&](u32 v, u32 idx) { return v < ctx->tax_child[idx]; });
        if(it == order.begin() || ctx->tax_child[*(it - 1)] != t) { found = false; return 0; }
        found = true;
        return ctx->tax_parent[*(it - 1)];
    };
    std::vector<u32> values;
    values.push_back(1);
    for(u32 i = 0; i < n_taxids; ++i) {
        u32 t = taxids[i];
        for(int guard = 0; t && guard < 4096; ++guard) {
            values.push_back(t);
            bool found;
            t = parent_of(t, found);
            if(!found) break;
        }
    }
    std::sort(values.begin(), values.end());
    values.erase(std::unique(values.begin(), values.end()), values.end());
    free_table(ctx);
    ctx->values = values;
    int rc = upload_values(ctx);
    if(rc == BNS_OK) rc = alloc_table(ctx, choose_bits(max_kmers, (u32)values.size()), (int)LAYOUT_HASH);
    if(rc == BNS_OK) rc = finalize_taxonomy(ctx);
    if(rc != BNS_OK) return rc;
    CK(cudaMemsetAsync(ctx->d_counters + 8, 0, 4 * sizeof(unsigned long long), ctx->slots[0].st));
    CK(cudaStreamSynchronize(ctx->slots[0].st));
    ctx->building = true;
    return BNS_OK;
}

int bns_b200_build_add_genome(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_records, uint32_t taxid) {
    if(!ctx || !offsets || (n_records && !bases)) return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    if(!ctx->building) return ctx->fail(BNS_E_STATE, "bns_b200_build_begin first");
    CK(cudaSetDevice(ctx->device));
    const auto it = std::lower_bound(ctx->values.begin(), ctx->values.end(), taxid);
    if(it == ctx->values.end() || *it != taxid) return ctx->fail(BNS_E_INVAL, "taxid %u was not announced to build_begin", taxid);
    const u32 vid = (u32)(it - ctx->values.begin());
    if(!n_records) return BNS_OK;
    // Pieces: long records are cut so that many warps share a genome. Only where the encoder keeps no state across
    // positions other than the window itself (FAM_U, FAM_K): piece i re-reads the W-1 elements before its first window.
    const EncParams &P = ctx->enc;
    const bool can_split = P.family == FAM_U || P.family == FAM_K;
    const u64 SPLIT = 1 << 14, c = P.c, W = P.W;
    std::vector<u64> se;
    for(u64 r = 0; r < n_records; ++r) {
        const u64 b = offsets[r], e = offsets[r + 1], L = e - b;
        if(!can_split || L < c || L - c + 1 <= SPLIT) { se.push_back(b); se.push_back(e); continue; }
        const u64 npos = L - c + 1;
        for(u64 s = 0; s < npos; s += SPLIT) {
            const u64 a = s >= W - 1 ? s - (W - 1) : 0, z = std::min(s + SPLIT, npos);
            se.push_back(b + a); se.push_back(b + z + c - 1);
        }
    }
    const u64 npieces = se.size() / 2, nb = offsets[n_records] - offsets[0];
    Slot &s = ctx->slots[0];
    int rc = ensure(s.d_bases, s.cap_bases, nb + 16);
    if(rc == BNS_OK) rc = ensure(s.d_offsets, s.cap_offsets, se.size() + 1);
    if(rc != BNS_OK) return ctx->fail(rc, "device workspace");
    CK(cudaMemcpyAsync(s.d_bases, bases + offsets[0], nb, cudaMemcpyHostToDevice, s.st));
    CK(cudaMemcpyAsync(s.d_offsets, se.data(), se.size() * 8, cudaMemcpyHostToDevice, s.st));
    const size_t smem = stream_smem_bytes(ctx->ring_cap, false);
    CK(launch_build(ctx->enc, grid_for(ctx, npieces, 4), smem, s.st, s.d_bases - offsets[0], s.d_offsets, npieces, offsets[n_records],
                    ctx->d_slots, table_fmt(ctx), vid, tax_view(ctx), ctx->d_values, (u32)ctx->values.size(), ctx->d_counters + 8,
                    ctx->ring_cap));
    ++ctx->stats.kernel_launches;
    ctx->stats.h2d_bytes += nb + se.size() * 8;
    ctx->stats.bases_processed += nb;
    CK(cudaStreamSynchronize(s.st));          // `se` and the caller's buffers may go away
    return BNS_OK;
}

int bns_b200_build_finish(bns_b200_t *ctx) {
    if(!ctx) return BNS_E_INVAL;
    if(!ctx->building) return ctx->fail(BNS_E_STATE, "no build in progress");
    CK(cudaSetDevice(ctx->device));
    ctx->building = false;
    unsigned long long h[4];
    CK(cudaMemcpy(h, ctx->d_counters + 8, sizeof h, cudaMemcpyDeviceToHost));
    if(h[2]) return ctx->fail(BNS_E_TAXONOMY, "%llu k-mers could not be merged: an lca of two genomes' taxids is not an ancestor of either "
                              "in the loaded taxonomy (malformed nodes file?)", h[2]);
    if(h[0]) return ctx->fail(BNS_E_CAPACITY, "%llu k-mers found no slot: call build_begin with a larger max_kmers", h[0]);
    ctx->n_displaced = 0;
    ctx->tax_ready = true;
    int rc = refresh_table_stats(ctx);
    if(rc != BNS_OK) return rc;
    // The table was sized for the caller's bound and built in the hash layout; minimised sets are far smaller. Re-home the
    // entries (device to device) into a table of the right size -- so that a small database stays L2-resident -- and of the
    // layout that size calls for (a minimizer-layout attempt that finds no room is redone in the hash layout).
    ctx->no_minimizer = false;
    u32 want = std::min(choose_bits(ctx->n_keys, (u32)ctx->values.size()), ctx->bucket_bits);
    const u32 built_bits = ctx->bucket_bits;
    cudaStream_t st = ctx->slots[0].st;
    const u64 n = ctx->n_keys;
    u64 *dk = nullptr;
    u32 *dv = nullptr;
    bool dumped = false;
    auto cleanup = [&](int code) { if(dk) cudaFree(dk); if(dv) cudaFree(dv); return code; };
    while(want <= built_bits) {
        const bool mini = want_minimizer_layout(ctx, want);
        if(!dumped && want == ctx->bucket_bits && !mini && ctx->layout == LAYOUT_HASH) break;   // already there
        if(!dumped) {                                                     // the built table as pairs, once, for every attempt below
            cudaError_t e;
            if((e = cudaMalloc((void **)&dk, std::max<u64>(n, 1) * 8)) != cudaSuccess || (e = cudaMalloc((void **)&dv, std::max<u64>(n, 1) * 4)) != cudaSuccess ||
               (e = cudaMemsetAsync(ctx->d_counters + 12, 0, sizeof(unsigned long long), st)) != cudaSuccess ||
               (e = launch_dump(st, ctx->d_slots, ctx->n_buckets, table_fmt(ctx), ctx->d_values, dk, dv, n, ctx->d_counters + 12)) != cudaSuccess ||
               (e = cudaStreamSynchronize(st)) != cudaSuccess)
                return cleanup(ctx->cuda_fail(e, "dumping the built table"));
            ++ctx->stats.kernel_launches;
            dumped = true;
        }
        const TableFmt nf = fmt_for(ctx, want, mini);
        u64 *new_slots = nullptr;
        unsigned long long h2[3];
        if(mini && (rc = ensure_fail_list(ctx)) != BNS_OK) return cleanup(rc);
        auto rehome = [&]() -> cudaError_t {
            cudaError_t e;
            if((e = cudaMalloc((void **)&new_slots, (1ull << want) * 32)) != cudaSuccess) return e;
            if((e = cudaMemsetAsync(new_slots, 0xff, (1ull << want) * 32, st)) != cudaSuccess) return e;
            if((e = cudaMemsetAsync(ctx->d_counters + 13, 0, 3 * sizeof(unsigned long long), st)) != cudaSuccess) return e;
            if((e = launch_insert(st, new_slots, nf, dk, dv, n, ctx->d_values, (u32)ctx->values.size(), ctx->d_counters + 13,
                                  mini ? ctx->d_fail_k : nullptr, ctx->d_fail_v, FAIL_CAP)) != cudaSuccess) return e;
            ++ctx->stats.kernel_launches;
            if((e = cudaMemcpyAsync(h2, ctx->d_counters + 13, sizeof h2, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
            return cudaStreamSynchronize(st);
        };
        const cudaError_t re = rehome();
        if(re != cudaSuccess) { if(new_slots) cudaFree(new_slots); return cleanup(ctx->cuda_fail(re, "re-homing the built table")); }
        if(!h2[2] && (rc = settle_failures(ctx, h2, n, mini ? (int)LAYOUT_MINIMIZER : (int)LAYOUT_HASH)) != BNS_OK) { cudaFree(new_slots); return cleanup(rc); }
        if(h2[0] || h2[2]) {                                              // no room
            cudaFree(new_slots);
            if(mini) ctx->no_minimizer = true;                            // same size, hash layout
            else ++want;
            continue;
        }
        cudaFree(ctx->d_slots);
        ctx->d_slots = new_slots;
        adopt_fmt(ctx, nf);
        ctx->n_displaced = h2[1];
        rc = refresh_table_stats(ctx);
        if(rc != BNS_OK || !minimizer_build_too_crowded(ctx)) return cleanup(rc);
        ctx->no_minimizer = true;                                         // same size, hash layout
    }
    return cleanup(BNS_OK);
}

int bns_b200_table_dump(bns_b200_t *ctx, uint64_t *keys_out, uint32_t *vals_out, uint64_t cap, uint64_t *n_out) {
    if(!ctx || !n_out || (cap && (!keys_out || !vals_out))) return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    if(!ctx->d_slots) return ctx->fail(BNS_E_STATE, "no table loaded");
    CK(cudaSetDevice(ctx->device));
    *n_out = ctx->n_keys;
    if(!cap) return BNS_OK;
    if(cap < ctx->n_keys) return ctx->fail(BNS_E_CAPACITY, "table holds %llu keys", (unsigned long long)ctx->n_keys);
    u64 *dk = nullptr; u32 *dv = nullptr;
    cudaStream_t st = ctx->slots[0].st;
    auto dump = [&]() -> cudaError_t {
        cudaError_t e;
        if((e = cudaMalloc((void **)&dk, std::max<u64>(ctx->n_keys, 1) * 8)) != cudaSuccess) return e;
        if((e = cudaMalloc((void **)&dv, std::max<u64>(ctx->n_keys, 1) * 4)) != cudaSuccess) return e;
        if((e = cudaMemsetAsync(ctx->d_counters + 12, 0, sizeof(unsigned long long), st)) != cudaSuccess) return e;
        if((e = launch_dump(st, ctx->d_slots, ctx->n_buckets, table_fmt(ctx), ctx->d_values, dk, dv, ctx->n_keys, ctx->d_counters + 12)) != cudaSuccess) return e;
        ++ctx->stats.kernel_launches;
        if(ctx->d_stash) {                                                 // ... and the stash, behind them
            if((e = launch_dump(st, ctx->d_stash, 1ull << ctx->stash_bits, stash_fmt(ctx), ctx->d_values, dk, dv, ctx->n_keys, ctx->d_counters + 12)) != cudaSuccess) return e;
            ++ctx->stats.kernel_launches;
        }
        if((e = cudaMemcpyAsync(keys_out, dk, ctx->n_keys * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if((e = cudaMemcpyAsync(vals_out, dv, ctx->n_keys * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        return cudaStreamSynchronize(st);
    };
    const cudaError_t de = dump();
    if(dk) cudaFree(dk);
    if(dv) cudaFree(dv);
    if(de != cudaSuccess) return ctx->cuda_fail(de, "table dump");
    return BNS_OK;
}

// ---- taxonomy ------------------------------------------------------------------------------------
int bns_b200_load_taxonomy(bns_b200_t *ctx, const uint32_t *child, const uint32_t *parent, uint64_t n) {
    if(!ctx || (n && (!child || !parent))) return ctx ? ctx->fail(BNS_E_INVAL, "null taxonomy arrays") : BNS_E_INVAL;
    ctx->tax_child.assign(child, child + n);
    ctx->tax_parent.assign(parent, parent + n);
    ctx->tax_loaded = true;
    ctx->tax_ready = false;
    return BNS_OK;
}

// build_parent_map, util.h:766-785: key = atoi(line), parent = atoi(strchr(line, '|') + 2); '#', empty lines skipped
int bns_b200_load_taxonomy_file(bns_b200_t *ctx, const char *path) {
    if(!ctx || !path) return BNS_E_INVAL;
    FILE *fp = fopen(path, "r");
    if(!fp) return ctx->fail(BNS_E_IO, "cannot open %s", path);
    std::vector<u32> child, parent;
    char *line = nullptr;
    size_t cap = 0;
    ssize_t len;
    while((len = getline(&line, &cap, fp)) >= 0) {
        if(len && line[len - 1] == '\n') line[len - 1] = 0;
        if(line[0] == '\n' || line[0] == '\0' || line[0] == '#') continue;
        const char *p = strchr(line, '|');
        child.push_back((u32)atoi(line));
        parent.push_back(p ? (u32)atoi(p + 2) : 0xffffffffu);
    }
    free(line);
    fclose(fp);
    if(child.size() < 1) return ctx->fail(BNS_E_IO, "Failed to create taxmap from %s", path);   // util.h:782
    return bns_b200_load_taxonomy(ctx, child.data(), parent.data(), child.size());
}

int bns_b200_resolve_batch(bns_b200_t *ctx, const uint32_t *taxa, const uint16_t *counts, const uint64_t *offsets,
                           uint64_t n_lists, uint32_t *taxon_out) {
    if(!ctx || !offsets || !taxon_out) return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    if(!ctx->d_values) return ctx->fail(BNS_E_STATE, "no table loaded (resolve needs its value dictionary)");
    CK(cudaSetDevice(ctx->device));
    int rc = finalize_taxonomy(ctx);
    if(rc != BNS_OK) return rc;
    if(!n_lists) return BNS_OK;
    const u64 total = offsets[n_lists];
    u32 *d_taxa = nullptr, *d_out = nullptr;
    uint16_t *d_counts = nullptr;
    u64 *d_off = nullptr;
    cudaStream_t st = ctx->slots[0].st;
    CK(cudaMalloc((void **)&d_taxa, std::max<u64>(total, 1) * 4));
    CK(cudaMalloc((void **)&d_counts, std::max<u64>(total, 1) * 2));
    CK(cudaMalloc((void **)&d_off, (n_lists + 1) * 8));
    CK(cudaMalloc((void **)&d_out, n_lists * 4));
    CK(cudaMemcpyAsync(d_taxa, taxa, total * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_counts, counts, total * 2, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_off, offsets, (n_lists + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->d_status, 0, 4, st));
    CK(launch_resolve(grid_for(ctx, n_lists, 4), st, tax_view(ctx), ctx->d_values, (u32)ctx->values.size(), d_taxa, d_counts,
                      d_off, n_lists, d_out, ctx->d_status));
    ++ctx->stats.kernel_launches;
    u32 status = 0;
    CK(cudaMemcpyAsync(taxon_out, d_out, n_lists * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&status, ctx->d_status, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_taxa); cudaFree(d_counts); cudaFree(d_off); cudaFree(d_out);
    return check_status(ctx, status);
}

// ---- replication -----------------------------------------------------------------------------------
int bns_b200_db_export_header(const bns_b200_t *ctx_, bns_b200_db_header *hdr) {
    bns_b200_t *ctx = const_cast<bns_b200_t *>(ctx_);
    if(!ctx || !hdr) return BNS_E_INVAL;
    if(!ctx->d_slots) return ctx->fail(BNS_E_STATE, "no table loaded");
    int rc = finalize_taxonomy(ctx);
    if(rc != BNS_OK) return rc;
    memset(hdr, 0, sizeof *hdr);
    hdr->words[0] = 0x42304e53424e5331ull;      // magic
    hdr->words[1] = ctx->bucket_bits;
    hdr->words[2] = ctx->values.size();
    hdr->words[3] = ctx->n_nodes;
    hdr->words[4] = ctx->node_of_one;
    hdr->words[5] = ctx->n_keys;
    hdr->words[6] = ctx->n_displaced;
    hdr->words[7] = ctx->n_overflowed;
    hdr->words[8] = ctx->max_disp;
    hdr->words[9] = ctx->flag_count;
    hdr->words[10] = ctx->layout;
    hdr->words[11] = ctx->fmt_bits;
    hdr->words[12] = ctx->table_k;
    hdr->words[13] = ctx->disp_bits;
    hdr->words[14] = ctx->d_stash ? ((u64)ctx->stash_bits | ((u64)ctx->stash_flags << 8)) : 0;
    hdr->words[15] = ctx->n_stash;
    return BNS_OK;
}

int bns_b200_db_alloc_from_header(bns_b200_t *ctx, const bns_b200_db_header *hdr) {
    if(!ctx || !hdr || hdr->words[0] != 0x42304e53424e5331ull) return ctx ? ctx->fail(BNS_E_INVAL, "bad database header") : BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    free_table(ctx);
    if(ctx->d_node_info) cudaFree(ctx->d_node_info);
    if(ctx->d_val_info) cudaFree(ctx->d_val_info);
    ctx->d_node_info = ctx->d_val_info = nullptr;
    ctx->values.assign((size_t)hdr->words[2], 0);
    ctx->n_nodes = (u32)hdr->words[3];
    ctx->node_of_one = (u32)hdr->words[4];
    ctx->n_keys = hdr->words[5]; ctx->n_displaced = hdr->words[6]; ctx->n_overflowed = hdr->words[7];
    ctx->max_disp = (u32)hdr->words[8];
    int rc = alloc_table(ctx, (u32)hdr->words[1], (int)hdr->words[10]);
    if(rc != BNS_OK) return rc;
    // the slot format is the sender's (this context may be configured for another k)
    ctx->fmt_bits = (u32)hdr->words[11]; ctx->table_k = (u32)hdr->words[12];
    ctx->disp_bits = hdr->words[13] ? (u32)hdr->words[13] : (u32)DISP_BITS;
    ctx->flag_count = (u32)hdr->words[9];
    if(!ctx->flag_count || (ctx->flag_count & (ctx->flag_count - 1)) || ctx->fmt_bits < ctx->disp_bits + ctx->flag_count)
        return ctx->fail(BNS_E_INVAL, "database header: bad slot format");
    if(hdr->words[14]) {
        ctx->stash_bits = (u32)(hdr->words[14] & 0xff); ctx->stash_flags = (u32)(hdr->words[14] >> 8); ctx->n_stash = hdr->words[15];
        if(ctx->stash_bits > 32 || !ctx->stash_flags) return ctx->fail(BNS_E_INVAL, "database header: bad stash format");
        CK(cudaMalloc((void **)&ctx->d_stash, 32ull << ctx->stash_bits));
    }
    CK(cudaMalloc((void **)&ctx->d_values, std::max<size_t>(ctx->values.size(), 1) * sizeof(u32)));
    CK(cudaMalloc((void **)&ctx->d_val_info, std::max<size_t>(ctx->values.size(), 1) * sizeof(uint4)));
    CK(cudaMalloc((void **)&ctx->d_node_info, std::max<u32>(ctx->n_nodes, 1) * sizeof(uint4)));
    CK(cudaStreamSynchronize(ctx->slots[0].st));
    return BNS_OK;
}

int bns_b200_db_segments(const bns_b200_t *ctx, void **dev_ptrs, uint64_t *bytes, int cap, int *n) {
    if(!ctx || !dev_ptrs || !bytes || !n || cap < 4) return BNS_E_INVAL;
    if(!ctx->d_slots || !ctx->d_val_info) return BNS_E_STATE;
    dev_ptrs[0] = ctx->d_slots;     bytes[0] = ctx->n_buckets * 32;
    dev_ptrs[1] = ctx->d_values;    bytes[1] = ctx->values.size() * sizeof(u32);
    dev_ptrs[2] = ctx->d_val_info;  bytes[2] = ctx->values.size() * sizeof(uint4);
    dev_ptrs[3] = ctx->d_node_info; bytes[3] = (uint64_t)ctx->n_nodes * sizeof(uint4);
    *n = 4;
    if(ctx->d_stash && cap >= 5) { dev_ptrs[4] = ctx->d_stash; bytes[4] = 32ull << ctx->stash_bits; *n = 5; }
    else if(ctx->d_stash) return BNS_E_CAPACITY;
    return BNS_OK;
}

int bns_b200_db_commit(bns_b200_t *ctx) {
    if(!ctx) return BNS_E_INVAL;
    if(!ctx->d_slots || !ctx->d_val_info) return ctx->fail(BNS_E_STATE, "nothing to commit");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    if(!ctx->values.empty())
        CK(cudaMemcpy(ctx->values.data(), ctx->d_values, ctx->values.size() * sizeof(u32), cudaMemcpyDeviceToHost));
    ctx->tax_ready = true;
    ctx->tax_loaded = true;
    return BNS_OK;
}

// ---- several GPUs in one process ------------------------------------------------------------------------
// NCCL is bound at the first replicate call (dlopen of libnccl.so.2, prototypes from <nccl.h>): a process that drives one
// GPU never loads it, and a host process that already carries an NCCL (torch.distributed) shares that copy instead of
// mapping a second one under the same soname.
namespace {
struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};
NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        // NCCL prints its warnings to stdout unless told otherwise: stdout is where `bonsai classify` writes its records
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
        for(const char *name : {"libnccl.so.2", "libnccl.so"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if(a.lib) break;
        }
        if(!a.lib) return a;
        a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.lib, "ncclCommInitAll");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
        a.Broadcast = (decltype(a.Broadcast))dlsym(a.lib, "ncclBroadcast");
        a.GroupStart = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
        a.ok = a.CommInitAll && a.CommDestroy && a.Broadcast && a.GroupStart && a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}
}  // namespace

int bns_b200_open_multi(const bns_b200_config *cfg, int n_gpus, const int *devices, bns_b200_t **out, int *n_out) {
    if(!cfg || !out || n_gpus < 0) return BNS_E_INVAL;
    if(n_out) *n_out = 0;
    int ndev = 0;
    if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_open_err = "no usable CUDA device";
        return BNS_E_CUDA;
    }
    int n = n_gpus ? n_gpus : (cfg->n_gpus ? (int)cfg->n_gpus : ndev);
    if(n > ndev && !devices) { g_open_err = "more GPUs requested than visible"; return BNS_E_INVAL; }
    for(int i = 0; i < n; ++i) out[i] = nullptr;
    for(int i = 0; i < n; ++i) {
        bns_b200_config c = *cfg;
        c.device = devices ? devices[i] : i;
        const int rc = bns_b200_open(&c, &out[i]);
        if(rc != BNS_OK) {
            for(int j = 0; j < i; ++j) { bns_b200_close(out[j]); out[j] = nullptr; }
            return rc;
        }
    }
    if(n_out) *n_out = n;
    return BNS_OK;
}

void bns_b200_close_multi(bns_b200_t **handles, int n) {
    if(!handles) return;
    for(int i = 0; i < n; ++i) { bns_b200_close(handles[i]); handles[i] = nullptr; }
}

int bns_b200_replicate(bns_b200_t *const *handles, int n, int root) {
    if(!handles || n < 1 || root < 0 || root >= n || !handles[root]) return BNS_E_INVAL;
    bns_b200_ctx *ctx = handles[root];                                     // errors are reported on the root context
    for(int i = 0; i < n; ++i) {
        if(!handles[i]) return ctx->fail(BNS_E_INVAL, "null context in the replica list");
        for(int j = 0; j < i; ++j)
            if(handles[j]->device == handles[i]->device) return ctx->fail(BNS_E_INVAL, "two contexts of the replica list share device %d", handles[i]->device);
    }
    if(!ctx->d_slots) return ctx->fail(BNS_E_STATE, "no table loaded on the root context");
    if(n == 1) return BNS_OK;
    bns_b200_db_header hdr;
    int rc = bns_b200_db_export_header(ctx, &hdr);
    if(rc != BNS_OK) return rc;
    for(int i = 0; i < n; ++i)
        if(i != root && (rc = bns_b200_db_alloc_from_header(handles[i], &hdr)) != BNS_OK)
            return ctx->fail(rc, "replica %d: %s", i, handles[i]->err.c_str());
    // BNS_B200_REPLICATE=p2p: peer copies (cudaMemcpyPeerAsync over NVLink) instead of NCCL. ncclCommInitAll costs seconds of
    // topology discovery per process, which a short `bonsai classify --gpus N` run feels; the broadcast itself is milliseconds.
    if(const char *how = getenv("BNS_B200_REPLICATE")) {
        if(!strcmp(how, "p2p")) {
            std::vector<void *> ptrs((size_t)n * 8);
            uint64_t bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int nseg = 0;
            for(int i = 0; i < n; ++i) {
                uint64_t b[8]; int ns = 0;
                rc = bns_b200_db_segments(handles[i], &ptrs[(size_t)i * 8], b, 8, &ns);
                if(rc != BNS_OK || ns < 4) return ctx->fail(BNS_E_STATE, "replica %d has no segments", i);
                if(i == root) { nseg = ns; for(int s = 0; s < ns; ++s) bytes[s] = b[s]; }
            }
            cudaSetDevice(ctx->device);                                   // the copies are queued on the root's stream
            for(int i = 0; i < n; ++i) {
                if(i == root) continue;
                for(int s = 0; s < nseg; ++s)
                    if(bytes[s]) {
                        const cudaError_t e = cudaMemcpyPeerAsync(ptrs[(size_t)i * 8 + s], handles[i]->device, ptrs[(size_t)root * 8 + s], ctx->device,
                                                                  (size_t)bytes[s], ctx->slots[0].st);
                        if(e != cudaSuccess) return ctx->cuda_fail(e, "cudaMemcpyPeerAsync");
                    }
            }
            cudaSetDevice(ctx->device);
            const cudaError_t e = cudaStreamSynchronize(ctx->slots[0].st);
            if(e != cudaSuccess) return ctx->cuda_fail(e, "replication stream");
            for(int i = 0; i < n; ++i)
                if(i != root && (rc = bns_b200_db_commit(handles[i])) != BNS_OK) return ctx->fail(rc, "replica %d: %s", i, handles[i]->err.c_str());
            cudaSetDevice(ctx->device);
            return BNS_OK;
        }
    }
    NcclApi &nc = nccl_api();
    if(!nc.ok) return ctx->fail(BNS_E_CUDA, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
    // NCCL announces itself on stdout ("NCCL version ..."), which is where `bonsai classify` writes its records: while NCCL
    // runs, file descriptor 1 is the process's stderr
    struct StdoutGuard {
        int saved;
        StdoutGuard() { fflush(stdout); saved = dup(1); if(saved >= 0) dup2(2, 1); }
        ~StdoutGuard() { if(saved >= 0) { fflush(stdout); dup2(saved, 1); close(saved); } }
    } stdout_guard;
    std::vector<ncclComm_t> comms((size_t)n, nullptr);
    std::vector<int> devs((size_t)n);
    for(int i = 0; i < n; ++i) devs[(size_t)i] = handles[i]->device;
    ncclResult_t nr = nc.CommInitAll(comms.data(), n, devs.data());
    if(nr != ncclSuccess) return ctx->fail(BNS_E_CUDA, "ncclCommInitAll: %s", nc.GetErrorString(nr));
    auto done = [&](int code) {
        for(int i = 0; i < n; ++i) if(comms[(size_t)i]) nc.CommDestroy(comms[(size_t)i]);
        return code;
    };
    std::vector<void *> ptrs((size_t)n * 8);
    uint64_t bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int nseg = 0;
    for(int i = 0; i < n; ++i) {
        uint64_t b[8]; int ns = 0;
        rc = bns_b200_db_segments(handles[i], &ptrs[(size_t)i * 8], b, 8, &ns);
        if(rc != BNS_OK || ns < 4) return done(ctx->fail(BNS_E_STATE, "replica %d has no segments", i));
        if(i == root) { nseg = ns; for(int s = 0; s < ns; ++s) bytes[s] = b[s]; }
    }
    for(int s = 0; s < nseg && nr == ncclSuccess; ++s) {                    // one broadcast per segment, every rank of it in one group
        if(!bytes[s]) continue;
        nc.GroupStart();
        for(int i = 0; i < n; ++i) {
            cudaSetDevice(handles[i]->device);
            const ncclResult_t r = nc.Broadcast(ptrs[(size_t)root * 8 + s], ptrs[(size_t)i * 8 + s], (size_t)bytes[s], ncclChar, root,
                                                comms[(size_t)i], handles[i]->slots[0].st);
            if(r != ncclSuccess) nr = r;
        }
        const ncclResult_t r = nc.GroupEnd();
        if(r != ncclSuccess) nr = r;
    }
    if(nr != ncclSuccess) return done(ctx->fail(BNS_E_CUDA, "ncclBroadcast: %s", nc.GetErrorString(nr)));
    for(int i = 0; i < n; ++i) {
        cudaSetDevice(handles[i]->device);
        const cudaError_t e = cudaStreamSynchronize(handles[i]->slots[0].st);
        if(e != cudaSuccess) return done(ctx->cuda_fail(e, "replication stream"));
    }
    done(0);
    for(int i = 0; i < n; ++i)
        if(i != root && (rc = bns_b200_db_commit(handles[i])) != BNS_OK) return ctx->fail(rc, "replica %d: %s", i, handles[i]->err.c_str());
    cudaSetDevice(ctx->device);
    return BNS_OK;
}

// ---- encode ----------------------------------------------------------------------------------------
int bns_b200_encode_batch(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_seqs,
                          uint64_t *kmers_out, const uint64_t *out_offsets, uint32_t *counts_out) {
    if(!ctx || !offsets || !out_offsets || !counts_out) return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    if(!n_seqs) return BNS_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t smem = stream_smem_bytes(ctx->ring_cap, false);
    const int occ = encode_occupancy(ctx->enc, smem);
    u32 status_acc = 0;
    int slot_i = 0;
    u64 r0 = 0;
    u32 h_status[N_SLOTS] = {};
    while(r0 < n_seqs) {
        u64 r1 = r0;
        while(r1 < n_seqs && r1 - r0 < CHUNK_READS && (r1 == r0 || offsets[r1 + 1] - offsets[r0] <= CHUNK_BASES)) ++r1;
        const u64 nb = offsets[r1] - offsets[r0], nr = r1 - r0;
        const u64 nk = out_offsets[r1] - out_offsets[r0];
        Slot &s = ctx->slots[slot_i];
        CK(cudaStreamSynchronize(s.st));
        status_acc |= h_status[slot_i];
        int rc = ensure(s.d_bases, s.cap_bases, nb + 16);
        if(rc == BNS_OK) rc = ensure(s.d_offsets, s.cap_offsets, nr + 1);
        if(rc == BNS_OK) rc = ensure(s.d_out_offsets, s.cap_out_offsets, nr + 1);
        if(rc == BNS_OK) rc = ensure(s.d_kmers, s.cap_kmers, nk + 1);
        if(rc == BNS_OK) rc = ensure(s.d_out, s.cap_out, nr + 1);
        if(rc != BNS_OK) return ctx->fail(rc, "device workspace");
        CK(cudaMemcpyAsync(s.d_bases, bases + offsets[r0], nb, cudaMemcpyHostToDevice, s.st));
        CK(cudaMemcpyAsync(s.d_offsets, offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, s.st));
        CK(cudaMemcpyAsync(s.d_out_offsets, out_offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, s.st));
        CK(cudaMemsetAsync(s.d_out + nr, 0, 4, s.st));
        CK(launch_encode(ctx->enc, grid_for(ctx, nr, occ), smem, s.st, s.d_bases - offsets[r0], s.d_offsets, nr, offsets[r1],
                         s.d_kmers - out_offsets[r0], s.d_out_offsets, s.d_out, ctx->ring_cap, s.d_out + nr));
        ++ctx->stats.kernel_launches;
        if(nk) CK(cudaMemcpyAsync(kmers_out + out_offsets[r0], s.d_kmers, nk * 8, cudaMemcpyDeviceToHost, s.st));
        CK(cudaMemcpyAsync(counts_out + r0, s.d_out, nr * 4, cudaMemcpyDeviceToHost, s.st));
        CK(cudaMemcpyAsync(&h_status[slot_i], s.d_out + nr, 4, cudaMemcpyDeviceToHost, s.st));
        ctx->stats.h2d_bytes += nb + 16 * (nr + 1);
        ctx->stats.d2h_bytes += nk * 8 + nr * 4;
        ctx->stats.reads_processed += nr;
        ctx->stats.bases_processed += nb;
        r0 = r1;
        slot_i = (slot_i + 1) % N_RING;
    }
    for(int i = 0; i < N_SLOTS; ++i) { CK(cudaStreamSynchronize(ctx->slots[i].st)); status_acc |= h_status[i]; }
    return check_status(ctx, status_acc & 1u);
}

// ---- classify --------------------------------------------------------------------------------------
static int classify_ready(bns_b200_t *ctx) {
    if(ctx->cfg.api == BNS_API_ITER) return ctx->fail(BNS_E_INVAL, "BNS_API_ITER is an encode-only configuration");
    if(!ctx->d_slots) return ctx->fail(BNS_E_STATE, "no table loaded");
    return finalize_taxonomy(ctx);
}

int bns_b200_classify_device(bns_b200_t *ctx, const char *d_bases, const uint64_t *d_offsets, uint64_t n_reads, int paired,
                             uint32_t *d_taxon, uint32_t *d_n_hit, uint32_t *d_n_missing,
                             uint32_t *d_taxa, const uint64_t *d_taxa_offsets, void *stream) {
    if(!ctx || !d_offsets || !d_taxon) return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = classify_ready(ctx);
    if(rc != BNS_OK) return rc;
    const u32 mates = paired ? 2 : 1;
    const u64 n_rec = n_reads / mates;
    if(!n_rec) return BNS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // Fully asynchronous: nothing is read back. The total base count (offsets[n_reads]) lives on the device; the generic
    // kernel, which bounds its staging loads with it, reads it there (sentinel ~0), the lean kernel does not need it.
    const u64 total_bases = ~0ull;
    const ClassifyPlan pl = plan_classify(ctx->enc, table_view(ctx), ctx->ring_cap, ctx->n_sm, n_rec, mates, d_taxa != nullptr, false, d_n_hit || d_n_missing);
    Slot &s0 = ctx->slots[0];
    if(pl.second_pass) {
        rc = ensure(s0.d_defer, s0.cap_defer, n_rec);
        if(rc == BNS_OK) rc = ensure(ctx->d_big, ctx->cap_big, pass2_scratch_words(pl));
        if(rc != BNS_OK) return ctx->fail(rc, "device workspace");
        CK(cudaMemsetAsync(s0.d_defer_cnt, 0, sizeof(unsigned long long), st));
    }
    int nl = 1;
    CK(cudaEventRecord(ctx->ev0, st));
    CK(launch_classify(ctx->enc, pl, st, d_bases, (const u64 *)d_offsets, n_rec, mates, total_bases,
                       table_view(ctx), tax_view(ctx), d_taxon, d_n_hit, d_n_missing, d_taxa,
                       (const u64 *)d_taxa_offsets, nullptr, ctx->ring_cap, ctx->d_counters, ctx->d_status,
                       s0.d_defer, s0.d_defer_cnt, &nl, nullptr, ctx->d_big));
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->stats.kernel_launches += nl;
    ctx->stats.reads_processed += n_reads;          // bases_processed is not known on the host for device-resident calls
    ctx->timed = true;
    return BNS_OK;
}

int bns_b200_classify_device_runs(bns_b200_t *ctx, const char *d_bases, const uint64_t *d_offsets, uint64_t n_reads, int paired,
                                  uint32_t *d_taxon, uint32_t *d_n_hit, uint32_t *d_n_missing, uint64_t *d_runs, uint64_t runs_cap,
                                  uint64_t *d_run_pos, uint32_t *d_n_runs, uint64_t *d_runs_total, void *stream) {
    if(!ctx || !d_offsets || !d_taxon || !d_n_hit || !d_runs || !d_run_pos || !d_n_runs || !d_runs_total)
        return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = classify_ready(ctx);
    if(rc != BNS_OK) return rc;
    const u32 mates = paired ? 2 : 1;
    const u64 n_rec = n_reads / mates;
    if(!n_rec) return BNS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const ClassifyPlan pl = plan_classify(ctx->enc, table_view(ctx), ctx->ring_cap, ctx->n_sm, n_rec, mates, true, false, true, true);
    if(!pl.runs) return ctx->fail(BNS_E_INVAL, "this encoder's run lists come from bns_b200_classify_batch_runs (host buffers)");
    Slot &s0 = ctx->slots[0];
    if(pl.second_pass) {
        rc = ensure(s0.d_defer, s0.cap_defer, n_rec);
        if(rc == BNS_OK) rc = ensure(ctx->d_big, ctx->cap_big, pass2_scratch_words(pl));
        if(rc != BNS_OK) return ctx->fail(rc, "device workspace");
        CK(cudaMemsetAsync(s0.d_defer_cnt, 0, sizeof(unsigned long long), st));
    }
    CK(cudaMemsetAsync(d_runs_total, 0, sizeof(unsigned long long), st));
    int nl = 1;
    const RunsOut ro{(u64 *)d_runs, runs_cap, (unsigned long long *)d_runs_total, (u64 *)d_run_pos, d_n_runs};
    CK(cudaEventRecord(ctx->ev0, st));
    CK(launch_classify(ctx->enc, pl, st, d_bases, (const u64 *)d_offsets, n_rec, mates, ~0ull, table_view(ctx), tax_view(ctx), d_taxon, d_n_hit,
                       d_n_missing, nullptr, nullptr, nullptr, ctx->ring_cap, ctx->d_counters, ctx->d_status, s0.d_defer, s0.d_defer_cnt, &nl, &ro,
                       ctx->d_big));
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->stats.kernel_launches += nl;
    ctx->stats.reads_processed += n_reads;
    ctx->timed = true;
    return BNS_OK;
}

// classify_batch_ex for large batches when the context has packing threads: the batch is cut into chunks of PACK_CHUNK_READS
// reads that two lanes draw from. The PACK lane has the worker threads turn a chunk's ASCII bases into 2-bit units in pinned
// staging memory (bns_pack.h), then queues the copy of those (a quarter of the bytes) and the packed-input variant of the lean
// kernel; the RAW lane keeps the copy engine busy meanwhile with chunks that cross as they are (at most two queued). Host cores
// and PCIe link work side by side; every chunk's results land at its records' places in the caller's arrays.
static int classify_batch_packing(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, u64 n_rec_total, u32 mates,
                                  uint32_t *taxon_out, uint32_t *n_hit_out, uint32_t *n_missing_out, uint32_t *mate1_kmers_out, bool bases_pinned) {
    if(!ctx->pool) ctx->pool.reset(new PackPool((unsigned)ctx->pack_threads));
    PackPool &pool = *ctx->pool;
    const int pack_mode = ctx->pack_mode ? ctx->pack_mode : bases_pinned ? 2 : 1;
    for(int i = 0; i < N_SLOTS; ++i) { CK(cudaStreamSynchronize(ctx->slots[i].st)); ctx->slots[i].busy = ctx->slots[i].reserved = false; }
    const bool counts = n_hit_out || n_missing_out;
    // The chunk list, and for every chunk whether all its records have one length (then the kernel generates the offsets itself
    // and they do not cross PCIe): the scan of offsets[] is the one per-read pass of the call, the worker threads share it.
    struct Chunk { u64 q0, q1; u32 fixed_len; };
    std::vector<Chunk> chunks;
    // whole waves of the kernel per chunk (2^18 reads would be 2.3 waves on 148 SMs: a third of the warps idle in the last one)
    const u64 chunk_reads = whole_waves(ctx->pack_chunk_reads, ctx->n_sm);
    for(u64 q0 = 0; q0 < n_rec_total;) {
        u64 q1 = std::min(n_rec_total, q0 + std::max<u64>(1, chunk_reads / mates));
        while(q1 > q0 + 1 && offsets[q1 * mates] - offsets[q0 * mates] > CHUNK_BASES) q1 = q0 + (q1 - q0) / 2;
        chunks.push_back(Chunk{q0, q1, 0u});
        q0 = q1;
    }
    {
        Chunk *cp = chunks.data();
        pool.start((unsigned)chunks.size(), [=](unsigned i) {
            const u64 r0 = cp[i].q0 * mates, r1 = cp[i].q1 * mates, nr = r1 - r0, nb = offsets[r1] - offsets[r0];
            if(!nr || nb % nr || nb / nr == 0 || nb / nr >= 0xffffffffull) return;
            const u64 flen = nb / nr;
            u64 diff = 0;
            for(u64 r = r0; r < r1; ++r) diff |= (offsets[r + 1] - offsets[r]) ^ flen;
            if(!diff) cp[i].fixed_len = (u32)flen;
        });
        pool.wait();
    }
    // Every slot gets its buffers for the largest chunk now: which slot carries a packed and which an ASCII chunk is decided as the
    // call goes, and an allocation (pinned or device) in the middle of it would stall both lanes.
    {
        u64 max_nb = 0, max_nr = 0;
        for(const Chunk &c : chunks) {
            max_nb = std::max<u64>(max_nb, offsets[c.q1 * mates] - offsets[c.q0 * mates]);
            max_nr = std::max<u64>(max_nr, (c.q1 - c.q0) * mates);
        }
        const u64 n_units = (max_nb + 7) / 8, n_sw = (max_nb + 255) / 256;
        const int n_use = (int)std::min<size_t>(N_SLOTS, chunks.size());
        for(int i = 0; i < n_use; ++i) {
            Slot &s = ctx->slots[i];
            int rc = ensure(s.d_bases, s.cap_bases, std::max<u64>(max_nb + 16, 2 * n_units + 64));
            if(rc == BNS_OK) rc = ensure(s.d_offsets, s.cap_offsets, max_nr + 1);
            if(rc == BNS_OK) rc = ensure(s.d_out, s.cap_out, 4 * max_nr);
            if(rc == BNS_OK) rc = ensure(s.d_susp, s.cap_d_susp, n_sw + 2);
            if(rc == BNS_OK) rc = ensure(s.d_exc, s.cap_d_exc, 1024);
            if(rc == BNS_OK) rc = ensure_pinned(s.h_units, s.cap_h_units, n_units + 32);
            if(rc == BNS_OK) rc = ensure_pinned(s.h_susp, s.cap_h_susp, n_sw + 2);
            if(rc == BNS_OK) rc = ensure_pinned(s.h_exc, s.cap_h_exc, 1024);
            if(rc != BNS_OK) return ctx->fail(rc, "chunk buffers");
        }
    }
    // launch geometry per (records in the chunk, packed or not): all chunks but the last have one size
    struct PlanKey { u64 nq; bool pk; ClassifyPlan pl; };
    std::vector<PlanKey> plans;
    auto plan_for = [&](u64 nq, bool pk) -> ClassifyPlan {
        for(const PlanKey &k : plans) if(k.nq == nq && k.pk == pk) return k.pl;
        plans.push_back(PlanKey{nq, pk, plan_classify(ctx->enc, table_view(ctx), ctx->ring_cap, ctx->n_sm, nq, mates, false, mate1_kmers_out != nullptr, counts, false, pk)});
        return plans.back().pl;
    };
    auto free_slot = [&]() -> int {
        for(int i = 0; i < (int)std::min<size_t>(N_SLOTS, chunks.size()); ++i) {
            Slot &s = ctx->slots[i];
            if(s.reserved) continue;
            if(s.busy && cudaStreamQuery(s.st) == cudaSuccess) { s.busy = false; collect_kernel_time(ctx, s); }
            if(!s.busy) return i;
        }
        cudaGetLastError();                                            // cudaErrorNotReady is not an error
        return -1;
    };
    auto raw_queued = [&]() -> int {
        int n = 0;
        for(int i = 0; i < N_SLOTS; ++i) if(ctx->slots[i].busy && ctx->slots[i].raw && cudaEventQuery(ctx->slots[i].h2d_done) != cudaSuccess) ++n;
        cudaGetLastError();
        return n;
    };
    // copies + kernel + result copies of one chunk on its slot's stream; `pk` = the chunk was packed into the slot's staging
    auto enqueue = [&](int si, Chunk ch, bool pk, int set) -> int {
        Slot &s = ctx->slots[si];
        auto &exc_parts = ctx->exc_parts[set];
        const u64 r0 = ch.q0 * mates, r1 = ch.q1 * mates, nr = r1 - r0, nq = ch.q1 - ch.q0, nb = offsets[r1] - offsets[r0];
        ClassifyPlan pl = plan_for(nq, pk);
        if(pk && !pl.packed) return ctx->fail(BNS_E_STATE, "packed chunk without a packed-input kernel");
        const u64 n_units = (nb + 7) / 8, n_sw = (nb + 255) / 256;
        u64 n_exc = 0;
        int rc = ensure(s.d_out, s.cap_out, 4 * nq);
        if(rc == BNS_OK) rc = ensure(s.d_offsets, s.cap_offsets, nr + 1);
        if(rc == BNS_OK) rc = ensure(s.d_bases, s.cap_bases, pk ? 2 * n_units + 64 : nb + 16);
        if(rc == BNS_OK && pk) {
            for(auto &v : exc_parts) n_exc += v.size();
            rc = ensure(s.d_susp, s.cap_d_susp, n_sw + 2);
            if(rc == BNS_OK) rc = ensure(s.d_exc, s.cap_d_exc, n_exc + 1);
            if(rc == BNS_OK) rc = ensure_pinned(s.h_exc, s.cap_h_exc, n_exc + 1);
            if(n_exc >= 0xffffffffull) rc = BNS_E_CAPACITY;
        }
        if(rc != BNS_OK) return ctx->fail(rc, "device workspace");
        if(ch.fixed_len) { pl.fixed_len = ch.fixed_len; pl.fixed_base = offsets[r0]; }
        PackedIn pki{nullptr, nullptr, 0u, 0ull};
        if(pk) {
            u64 at = 0;
            for(auto &v : exc_parts) { if(!v.empty()) memcpy(s.h_exc + at, v.data(), v.size() * 8); at += v.size(); }
            CK(cudaMemcpyAsync(s.d_bases, s.h_units, 2 * n_units + 32, cudaMemcpyHostToDevice, s.st));   // + the zeroed slack the last tile may touch
            CK(cudaMemcpyAsync(s.d_susp, s.h_susp, (n_sw + 2) * 4, cudaMemcpyHostToDevice, s.st));
            if(n_exc) CK(cudaMemcpyAsync(s.d_exc, s.h_exc, n_exc * 8, cudaMemcpyHostToDevice, s.st));
            pki = PackedIn{s.d_susp, (const unsigned long long *)s.d_exc, (u32)n_exc, offsets[r0]};
            ctx->stats.h2d_bytes += 2 * n_units + 32 + (n_sw + 2) * 4 + n_exc * 8;
        } else {
            CK(cudaMemcpyAsync(s.d_bases, bases + offsets[r0], nb, cudaMemcpyHostToDevice, s.st));
            ctx->stats.h2d_bytes += nb;
        }
        if(!pl.fixed_len) { CK(cudaMemcpyAsync(s.d_offsets, offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, s.st)); ctx->stats.h2d_bytes += 8 * (nr + 1); }
        CK(cudaEventRecord(s.h2d_done, s.st));
        int nl = 1;
        CK(cudaEventRecord(s.ka, s.st));
        CK(launch_classify(ctx->enc, pl, s.st, pk ? s.d_bases : s.d_bases - offsets[r0], s.d_offsets, nq, mates, offsets[r1],
                           table_view(ctx), tax_view(ctx), s.d_out, n_hit_out ? s.d_out + nq : nullptr,
                           n_missing_out ? s.d_out + 2 * nq : nullptr, nullptr, nullptr, mate1_kmers_out ? s.d_out + 3 * nq : nullptr, ctx->ring_cap,
                           ctx->d_counters, ctx->d_status, s.d_defer, s.d_defer_cnt, &nl, nullptr, ctx->d_big, pk ? &pki : nullptr));
        CK(cudaEventRecord(s.kb, s.st));
        s.k_timed = true;
        ctx->stats.kernel_launches += nl;
        CK(cudaMemcpyAsync(taxon_out + ch.q0, s.d_out, nq * 4, cudaMemcpyDeviceToHost, s.st));
        if(n_hit_out) CK(cudaMemcpyAsync(n_hit_out + ch.q0, s.d_out + nq, nq * 4, cudaMemcpyDeviceToHost, s.st));
        if(n_missing_out) CK(cudaMemcpyAsync(n_missing_out + ch.q0, s.d_out + 2 * nq, nq * 4, cudaMemcpyDeviceToHost, s.st));
        if(mate1_kmers_out) CK(cudaMemcpyAsync(mate1_kmers_out + ch.q0, s.d_out + 3 * nq, nq * 4, cudaMemcpyDeviceToHost, s.st));
        ctx->stats.d2h_bytes += nq * 4 * (1 + (n_hit_out != nullptr) + (n_missing_out != nullptr) + (mate1_kmers_out != nullptr));
        ctx->stats.reads_processed += nr;
        ctx->stats.bases_processed += nb;
        s.busy = true; s.raw = !pk; s.reserved = false;
        return BNS_OK;
    };
    // the worker threads pack the chunk into the slot's staging buffers (pieces of whole suspicious-bit words)
    auto start_pack = [&](int si, Chunk ch, int set) -> int {
        Slot &s = ctx->slots[si];
        s.busy = s.reserved = true; s.raw = false;
        const u64 b0 = offsets[ch.q0 * mates], nb = offsets[ch.q1 * mates] - b0, n_units = (nb + 7) / 8, n_sw = (nb + 255) / 256;
        int rc = ensure_pinned(s.h_units, s.cap_h_units, n_units + 32);
        if(rc == BNS_OK) rc = ensure_pinned(s.h_susp, s.cap_h_susp, n_sw + 2);
        if(rc != BNS_OK) return ctx->fail(rc, "pinned staging memory");
        memset(s.h_units + n_units, 0, 64);
        s.h_susp[n_sw] = s.h_susp[n_sw + 1] = 0;
        const unsigned tasks = std::max(1u, std::min<unsigned>(pool.size() * 2, (unsigned)((nb + 65535) / 65536)));
        const u64 per = ((nb + tasks - 1) / tasks + 255) / 256 * 256;
        ctx->exc_parts[set].resize(tasks);
        for(auto &v : ctx->exc_parts[set]) v.clear();
        const char *src = bases + b0;
        uint16_t *hu = s.h_units; u32 *hs = s.h_susp;
        auto *parts = &ctx->exc_parts[set];
        pool.start(tasks, [=](unsigned t) {
            const u64 b = std::min(nb, per * t), e = std::min(nb, per * (t + 1));
            if(e > b) pack_range(src + b, e - b, hu + b / 8, hs + b / 256, (*parts)[t], b / 8);
        });
        return BNS_OK;
    };

    size_t c_next = 0;
    int rc = BNS_OK;
    bool packing = false;                                              // the workers are on pack_chunk, bound for slot pack_slot
    int pack_slot = -1, pack_set = 0;
    Chunk pack_chunk{0, 0, 0u};
    const bool verbose = getenv("BNS_B200_VERBOSE") != nullptr;
    u64 n_packed = 0, n_raw = 0;
    double t_idle = 0, t_enq = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_free = now();                                             // since when the workers have had nothing to do
    const double t_call = t_free;
    while(rc == BNS_OK && (c_next < chunks.size() || packing)) {
        bool progressed = false;
        int done_slot = -1, done_set = 0;
        Chunk done_chunk{0, 0, 0u};
        if(packing && pool.done()) {                                   // (a) a packed chunk is ready ...
            done_slot = pack_slot; done_chunk = pack_chunk; done_set = pack_set;
            packing = false;
            t_free = now();
            progressed = true;
        }
        if(!packing && c_next < chunks.size()) {                       // (b) ... the workers go on with the next one first ...
            const int si = free_slot();
            if(si >= 0) {
                pack_chunk = chunks[c_next++];
                pack_slot = si;
                pack_set ^= 1;
                rc = start_pack(si, pack_chunk, pack_set);
                if(rc != BNS_OK) break;
                t_idle += now() - t_free;
                packing = true;
                progressed = true;
            }
        }
        if(done_slot >= 0) {                                           // ... then the ready chunk's copies and kernel are queued
            const double t0 = now();
            rc = enqueue(done_slot, done_chunk, true, done_set);
            t_enq += now() - t0;
            ++n_packed;
            if(rc != BNS_OK) break;
        }
        if(pack_mode == 2 && c_next < chunks.size() && raw_queued() < 2) {   // (c) keep the copy engine fed with ASCII chunks
            const int si = free_slot();
            if(si >= 0) {
                rc = enqueue(si, chunks[c_next++], false, 0);
                ++n_raw;
                progressed = true;
            }
        }
        if(!progressed) std::this_thread::yield();
    }
    pool.wait();
    for(int i = 0; i < N_SLOTS; ++i) {
        const cudaError_t e = cudaStreamSynchronize(ctx->slots[i].st);
        if(e != cudaSuccess && rc == BNS_OK) rc = ctx->cuda_fail(e, "cudaStreamSynchronize");
        ctx->slots[i].busy = ctx->slots[i].reserved = false;
        collect_kernel_time(ctx, ctx->slots[i]);
    }
    if(verbose)
        fprintf(stderr, "[classify_batch] %llu chunks packed on %u threads (%s), %llu as ASCII; %.2f ms in all, workers without a chunk %.2f ms, queueing packed chunks %.2f ms\n",
                (unsigned long long)n_packed, pool.size(), pack_isa(), (unsigned long long)n_raw, (now() - t_call) * 1e3, t_idle * 1e3, t_enq * 1e3);
    return rc;
}

int bns_b200_classify_batch(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_reads, int paired,
                            uint32_t *taxon_out, uint32_t *n_hit_out, uint32_t *n_missing_out,
                            uint32_t *taxa_out, const uint64_t *taxa_offsets) {
    return bns_b200_classify_batch_ex(ctx, bases, offsets, n_reads, paired, taxon_out, n_hit_out, n_missing_out, taxa_out,
                                      taxa_offsets, nullptr);
}

int bns_b200_classify_batch_ex(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_reads, int paired,
                               uint32_t *taxon_out, uint32_t *n_hit_out, uint32_t *n_missing_out,
                               uint32_t *taxa_out, const uint64_t *taxa_offsets, uint32_t *mate1_kmers_out) {
    if(!ctx || !offsets || !taxon_out || (taxa_out && !taxa_offsets)) return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = classify_ready(ctx);
    if(rc != BNS_OK) return rc;
    const u32 mates = paired ? 2 : 1;
    const u64 n_rec_total = n_reads / mates;
    if(!n_rec_total) return BNS_OK;
    CK(cudaMemsetAsync(ctx->d_status, 0, 4, ctx->slots[0].st));
    CK(cudaStreamSynchronize(ctx->slots[0].st));
    // large batches of what `bonsai classify` runs: the host's cores pack chunks to 2 bits next to the chunks crossing as ASCII
    // (memory the CPU cannot read -- a device pointer handed to the host-buffer call copies device to device below -- is not packed)
    cudaPointerAttributes at;
    const bool at_ok = ctx->pack_threads > 0 && cudaPointerGetAttributes(&at, bases) == cudaSuccess;
    cudaGetLastError();
    if(at_ok && at.type != cudaMemoryTypeDevice && !taxa_out && offsets[n_rec_total * mates] - offsets[0] >= ctx->pack_min_bases &&
       plan_classify(ctx->enc, table_view(ctx), ctx->ring_cap, ctx->n_sm, n_rec_total, mates, false, mate1_kmers_out != nullptr,
                     n_hit_out || n_missing_out, false, true).packed) {
        rc = classify_batch_packing(ctx, bases, offsets, n_rec_total, mates, taxon_out, n_hit_out, n_missing_out, mate1_kmers_out,
                                    at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
        if(rc != BNS_OK) return rc;
        u32 status = 0;
        CK(cudaMemcpy(&status, ctx->d_status, 4, cudaMemcpyDeviceToHost));
        return check_status(ctx, status & 11u);
    }
    int slot_i = 0;
    u64 q0 = 0;                                                    // record cursor
    const u64 chunk_reads = whole_waves(CHUNK_READS, ctx->n_sm);
    while(q0 < n_rec_total) {
        u64 q1 = q0;
        while(q1 < n_rec_total && (q1 - q0) * mates < chunk_reads &&
              (q1 == q0 || offsets[(q1 + 1) * mates] - offsets[q0 * mates] <= CHUNK_BASES)) ++q1;
        const u64 r0 = q0 * mates, r1 = q1 * mates, nr = r1 - r0, nq = q1 - q0;
        const u64 nb = offsets[r1] - offsets[r0];
        Slot &s = ctx->slots[slot_i];
        CK(cudaStreamSynchronize(s.st));
        collect_kernel_time(ctx, s);
        rc = ensure(s.d_bases, s.cap_bases, nb + 16);
        if(rc == BNS_OK) rc = ensure(s.d_offsets, s.cap_offsets, nr + 1);
        if(rc == BNS_OK) rc = ensure(s.d_out, s.cap_out, 4 * nq);
        u64 nt = 0;
        if(rc == BNS_OK && taxa_out) {
            nt = taxa_offsets[q1] - taxa_offsets[q0];
            rc = ensure(s.d_taxa, s.cap_taxa, nt + 1);
            if(rc == BNS_OK) rc = ensure(s.d_taxa_offsets, s.cap_taxa_offsets, nq + 1);
        }
        ClassifyPlan pl = plan_classify(ctx->enc, table_view(ctx), ctx->ring_cap, ctx->n_sm, nq, mates, taxa_out != nullptr, mate1_kmers_out != nullptr,
                                         n_hit_out || n_missing_out);
        const bool windowed_lean = pl.lean && (pl.lean_mode == LEAN_K || pl.lean_mode == LEAN_R);
        if(rc == BNS_OK && pl.second_pass) rc = ensure(s.d_defer, s.cap_defer, nq);
        if(rc == BNS_OK && pl.second_pass) {
            // the scratch is shared by the slots: the pass that uses it is stream-ordered behind the previous chunk's only when
            // it does not move, so it is sized once for the largest plan before any chunk is in flight
            if(pass2_scratch_words(pl) > ctx->cap_big) { for(int i = 0; i < N_SLOTS; ++i) CK(cudaStreamSynchronize(ctx->slots[i].st)); }
            rc = ensure(ctx->d_big, ctx->cap_big, pass2_scratch_words(pl));
        }
        if(rc != BNS_OK) return ctx->fail(rc, "device workspace");
        if(pl.second_pass) CK(cudaMemsetAsync(s.d_defer_cnt, 0, sizeof(unsigned long long), s.st));
        // Fixed-length batches (every read of a sequencing run has the same length): the offsets are an arithmetic
        // progression the kernel generates itself; they do not cross PCIe (5 % of the bytes of a 150 bp read).
        if(pl.lean && !windowed_lean && nb % nr == 0 && nb / nr < 0xffffffffull) {
            const u64 flen = nb / nr;
            bool fixed = flen > 0;
            for(u64 r = r0; fixed && r < r1; ++r) fixed = offsets[r + 1] - offsets[r] == flen;
            if(fixed) { pl.fixed_len = (u32)flen; pl.fixed_base = offsets[r0]; }
        }
        CK(cudaMemcpyAsync(s.d_bases, bases + offsets[r0], nb, cudaMemcpyHostToDevice, s.st));
        if(!pl.fixed_len) CK(cudaMemcpyAsync(s.d_offsets, offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, s.st));
        if(taxa_out) CK(cudaMemcpyAsync(s.d_taxa_offsets, taxa_offsets + q0, (nq + 1) * 8, cudaMemcpyHostToDevice, s.st));
        int nl = 1;
        CK(cudaEventRecord(s.ka, s.st));
        CK(launch_classify(ctx->enc, pl, s.st, s.d_bases - offsets[r0], s.d_offsets, nq, mates, offsets[r1],
                           table_view(ctx), tax_view(ctx), s.d_out, n_hit_out ? s.d_out + nq : nullptr,
                           n_missing_out ? s.d_out + 2 * nq : nullptr, taxa_out ? s.d_taxa - taxa_offsets[q0] : nullptr,
                           taxa_out ? s.d_taxa_offsets : nullptr, mate1_kmers_out ? s.d_out + 3 * nq : nullptr, ctx->ring_cap,
                           ctx->d_counters, ctx->d_status, s.d_defer, s.d_defer_cnt, &nl, nullptr, ctx->d_big));
        CK(cudaEventRecord(s.kb, s.st));
        s.k_timed = true;
        ctx->stats.kernel_launches += nl;
        CK(cudaMemcpyAsync(taxon_out + q0, s.d_out, nq * 4, cudaMemcpyDeviceToHost, s.st));
        if(n_hit_out) CK(cudaMemcpyAsync(n_hit_out + q0, s.d_out + nq, nq * 4, cudaMemcpyDeviceToHost, s.st));
        if(n_missing_out) CK(cudaMemcpyAsync(n_missing_out + q0, s.d_out + 2 * nq, nq * 4, cudaMemcpyDeviceToHost, s.st));
        if(mate1_kmers_out) CK(cudaMemcpyAsync(mate1_kmers_out + q0, s.d_out + 3 * nq, nq * 4, cudaMemcpyDeviceToHost, s.st));
        if(taxa_out && nt) CK(cudaMemcpyAsync(taxa_out + taxa_offsets[q0], s.d_taxa, nt * 4, cudaMemcpyDeviceToHost, s.st));
        ctx->stats.h2d_bytes += nb + (pl.fixed_len ? 0 : 8 * (nr + 1)) + (taxa_out ? 8 * (nq + 1) : 0);
        ctx->stats.d2h_bytes += nq * 4 * (1 + (n_hit_out != nullptr) + (n_missing_out != nullptr)) + nt * 4;
        ctx->stats.reads_processed += nr;
        ctx->stats.bases_processed += nb;
        q0 = q1;
        slot_i = (slot_i + 1) % N_RING;
    }
    for(int i = 0; i < N_SLOTS; ++i) { CK(cudaStreamSynchronize(ctx->slots[i].st)); collect_kernel_time(ctx, ctx->slots[i]); }
    u32 status = 0;
    CK(cudaMemcpy(&status, ctx->d_status, 4, cudaMemcpyDeviceToHost));
    return check_status(ctx, status & 11u);
}

// classify_seqs with the hit lists run-length encoded on the device. Chunks are pipelined over the stream slots like
// classify_batch_ex; a chunk's runs are fetched (their count is only known once its kernel is done) while the next chunk's
// copies and kernels are already queued. What `bonsai classify` runs (every k-mer, no window) produces the runs in the lean
// kernel itself; the other encoders go through the ordered hit list of the generic kernel and bns_rle_kernel.
int bns_b200_classify_batch_runs(bns_b200_t *ctx, const char *bases, const uint64_t *offsets, uint64_t n_reads, int paired,
                                 uint32_t *taxon_out, uint32_t *n_hit_out, uint32_t *n_missing_out, uint32_t *mate1_kmers_out,
                                 uint64_t *runs_out, uint64_t runs_cap, uint64_t *run_pos_out, uint32_t *n_runs_out,
                                 uint64_t *n_runs_total_out) {
    if(!ctx || !offsets || !taxon_out || !n_hit_out || !run_pos_out || !n_runs_out || !n_runs_total_out || (runs_cap && !runs_out))
        return ctx ? ctx->fail(BNS_E_INVAL, "null buffers") : BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = classify_ready(ctx);
    if(rc != BNS_OK) return rc;
    *n_runs_total_out = 0;
    const u32 mates = paired ? 2 : 1;
    const u64 n_rec_total = n_reads / mates;
    if(!n_rec_total) return BNS_OK;
    CK(cudaMemsetAsync(ctx->d_status, 0, 4, ctx->slots[0].st));
    CK(cudaStreamSynchronize(ctx->slots[0].st));
    // a call that fails for want of room in runs_out must leave classified_[] as it found them (the caller calls again)
    unsigned long long counters0[2];
    CK(cudaMemcpy(counters0, ctx->d_counters, sizeof counters0, cudaMemcpyDeviceToHost));
    struct Pending { bool live = false; u64 q0 = 0, nq = 0; } pend[N_SLOTS];
    u64 used = 0;
    std::vector<u64> h_toffs;
    auto no_room = [&](u64 need) -> int {
        for(int i = 0; i < N_SLOTS; ++i) cudaStreamSynchronize(ctx->slots[i].st);
        cudaMemcpy(ctx->d_counters, counters0, sizeof counters0, cudaMemcpyHostToDevice);
        *n_runs_total_out = need;                                          // a lower bound on what the call needs
        return ctx->fail(BNS_E_CAPACITY, "the run buffer holds %llu entries, at least %llu are needed", (unsigned long long)runs_cap, (unsigned long long)need);
    };
    // fetch the runs of the chunk a slot holds: count first, then exactly that many entries
    auto finalize = [&](int si) -> int {
        Pending &pd = pend[si];
        if(!pd.live) return BNS_OK;
        pd.live = false;
        Slot &s = ctx->slots[si];
        unsigned long long total = 0;
        CK(cudaMemcpyAsync(&total, s.d_defer_cnt + 1, sizeof total, cudaMemcpyDeviceToHost, s.st));
        CK(cudaStreamSynchronize(s.st));
        collect_kernel_time(ctx, s);
        if(used + total > runs_cap) return no_room(used + total);
        if(total) CK(cudaMemcpyAsync(runs_out + used, s.d_runs, total * 8, cudaMemcpyDeviceToHost, s.st));
        CK(cudaMemcpyAsync(run_pos_out + pd.q0, s.d_run_pos, pd.nq * 8, cudaMemcpyDeviceToHost, s.st));
        CK(cudaMemcpyAsync(n_runs_out + pd.q0, s.d_nruns, pd.nq * 4, cudaMemcpyDeviceToHost, s.st));
        CK(cudaStreamSynchronize(s.st));
        for(u64 r = 0; r < pd.nq; ++r) run_pos_out[pd.q0 + r] += used;      // chunk-relative -> position in runs_out
        ctx->stats.d2h_bytes += total * 8 + pd.nq * 12;
        used += total;
        return BNS_OK;
    };
    int slot_i = 0;
    u64 q0 = 0;
    const u64 chunk_reads = whole_waves(CHUNK_READS, ctx->n_sm);
    while(q0 < n_rec_total) {
        u64 q1 = q0;
        while(q1 < n_rec_total && (q1 - q0) * mates < chunk_reads &&
              (q1 == q0 || offsets[(q1 + 1) * mates] - offsets[q0 * mates] <= CHUNK_BASES)) ++q1;
        const u64 r0 = q0 * mates, r1 = q1 * mates, nr = r1 - r0, nq = q1 - q0;
        const u64 nb = offsets[r1] - offsets[r0];
        rc = finalize(slot_i);                                             // the chunk this slot carried three chunks ago
        if(rc != BNS_OK) return rc;
        Slot &s = ctx->slots[slot_i];
        CK(cudaStreamSynchronize(s.st));
        ClassifyPlan pl = plan_classify(ctx->enc, table_view(ctx), ctx->ring_cap, ctx->n_sm, nq, mates, true, mate1_kmers_out != nullptr, true, true);
        // one run-buffer entry per k-mer position plus two per record bounds the hits, as the reference's vector does
        const u64 nt = nb + 2 * nq;
        rc = ensure(s.d_bases, s.cap_bases, nb + 16);
        if(rc == BNS_OK) rc = ensure(s.d_offsets, s.cap_offsets, nr + 1);
        if(rc == BNS_OK) rc = ensure(s.d_out, s.cap_out, 4 * nq);
        if(rc == BNS_OK) rc = ensure(s.d_runs, s.cap_runs, nt + 1 + runs_slack(pl));
        if(rc == BNS_OK) rc = ensure(s.d_run_pos, s.cap_run_pos, nq);
        if(rc == BNS_OK) rc = ensure(s.d_nruns, s.cap_nruns, nq);
        if(rc == BNS_OK && !pl.runs) {
            rc = ensure(s.d_taxa, s.cap_taxa, nt + 1);
            if(rc == BNS_OK) rc = ensure(s.d_taxa_offsets, s.cap_taxa_offsets, nq + 1);
        }
        if(rc == BNS_OK && pl.second_pass) rc = ensure(s.d_defer, s.cap_defer, nq);
        if(rc == BNS_OK && pl.second_pass) {
            if(pass2_scratch_words(pl) > ctx->cap_big) { for(int i = 0; i < N_SLOTS; ++i) CK(cudaStreamSynchronize(ctx->slots[i].st)); }
            rc = ensure(ctx->d_big, ctx->cap_big, pass2_scratch_words(pl));
        }
        if(rc != BNS_OK) return ctx->fail(rc, "device workspace");
        if(pl.runs && nb % nr == 0 && nb / nr < 0xffffffffull) {              // fixed-length batch: offsets stay on the host
            const u64 flen = nb / nr;
            bool fixed = flen > 0;
            for(u64 r = r0; fixed && r < r1; ++r) fixed = offsets[r + 1] - offsets[r] == flen;
            if(fixed) { pl.fixed_len = (u32)flen; pl.fixed_base = offsets[r0]; }
        }
        CK(cudaMemsetAsync(s.d_defer_cnt, 0, 2 * sizeof(unsigned long long), s.st)); // second-pass list and run counter of this chunk
        CK(cudaMemcpyAsync(s.d_bases, bases + offsets[r0], nb, cudaMemcpyHostToDevice, s.st));
        if(!pl.fixed_len) CK(cudaMemcpyAsync(s.d_offsets, offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, s.st));
        if(!pl.runs) {
            // hit-list windows of the chunk on the device: one slot per k-mer position plus two
            h_toffs.resize(nq + 1);
            h_toffs[0] = 0;
            for(u64 q = 0; q < nq; ++q) h_toffs[q + 1] = h_toffs[q] + (offsets[(q0 + q + 1) * mates] - offsets[(q0 + q) * mates]) + 2;
            CK(cudaMemcpyAsync(s.d_taxa_offsets, h_toffs.data(), (nq + 1) * 8, cudaMemcpyHostToDevice, s.st));
            CK(cudaStreamSynchronize(s.st));                               // h_toffs is reused by the next chunk
        }
        int nl = 1;
        const RunsOut ro{s.d_runs, s.cap_runs, s.d_defer_cnt + 1, s.d_run_pos, s.d_nruns};
        CK(cudaEventRecord(s.ka, s.st));
        CK(launch_classify(ctx->enc, pl, s.st, s.d_bases - offsets[r0], s.d_offsets, nq, mates, offsets[r1],
                           table_view(ctx), tax_view(ctx), s.d_out, s.d_out + nq, s.d_out + 2 * nq, pl.runs ? nullptr : s.d_taxa,
                           pl.runs ? nullptr : s.d_taxa_offsets, mate1_kmers_out ? s.d_out + 3 * nq : nullptr, ctx->ring_cap, ctx->d_counters,
                           ctx->d_status, s.d_defer, s.d_defer_cnt, &nl, &ro, ctx->d_big));
        if(!pl.runs) { CK(launch_rle(s.st, s.d_taxa, s.d_taxa_offsets, s.d_out + nq, nq, s.d_runs, s.d_defer_cnt + 1, s.d_run_pos, s.d_nruns)); ++nl; }
        CK(cudaEventRecord(s.kb, s.st));
        s.k_timed = true;
        ctx->stats.kernel_launches += nl;
        CK(cudaMemcpyAsync(taxon_out + q0, s.d_out, nq * 4, cudaMemcpyDeviceToHost, s.st));
        CK(cudaMemcpyAsync(n_hit_out + q0, s.d_out + nq, nq * 4, cudaMemcpyDeviceToHost, s.st));
        if(n_missing_out) CK(cudaMemcpyAsync(n_missing_out + q0, s.d_out + 2 * nq, nq * 4, cudaMemcpyDeviceToHost, s.st));
        if(mate1_kmers_out) CK(cudaMemcpyAsync(mate1_kmers_out + q0, s.d_out + 3 * nq, nq * 4, cudaMemcpyDeviceToHost, s.st));
        ctx->stats.h2d_bytes += nb + (pl.fixed_len ? 0 : 8 * (nr + 1)) + (pl.runs ? 0 : 8 * (nq + 1));
        ctx->stats.d2h_bytes += nq * 4 * (2 + (n_missing_out != nullptr) + (mate1_kmers_out != nullptr));
        ctx->stats.reads_processed += nr;
        ctx->stats.bases_processed += nb;
        pend[slot_i].live = true; pend[slot_i].q0 = q0; pend[slot_i].nq = nq;
        q0 = q1;
        slot_i = (slot_i + 1) % N_RING;
    }
    for(int i = 0; i < N_RING; ++i) {                                       // in chunk order: positions in runs_out follow the records
        rc = finalize((slot_i + i) % N_RING);
        if(rc != BNS_OK) return rc;
    }
    *n_runs_total_out = used;
    u32 status = 0;
    CK(cudaMemcpy(&status, ctx->d_status, 4, cudaMemcpyDeviceToHost));
    return check_status(ctx, status & 11u);
}

int bns_b200_set_host_pack_threads(bns_b200_t *ctx, uint32_t n_threads) {
    if(!ctx) return BNS_E_INVAL;
    const int n = (int)std::min<uint32_t>(n_threads, 64u);
    if(n != ctx->pack_threads) { ctx->pool.reset(); ctx->pack_threads = n; }
    return BNS_OK;
}
int bns_b200_host_pack_threads(const bns_b200_t *ctx) { return ctx ? ctx->pack_threads : BNS_E_INVAL; }

int bns_b200_device_status(bns_b200_t *ctx) {
    if(!ctx) return BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    u32 status = 0;
    CK(cudaMemcpy(&status, ctx->d_status, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemset(ctx->d_status, 0, 4));
    return check_status(ctx, status & 11u);
}

int bns_b200_sync(bns_b200_t *ctx) {
    if(!ctx) return BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    return BNS_OK;
}

int bns_b200_stats_get(const bns_b200_t *ctx_, bns_b200_stats *out) {
    bns_b200_t *ctx = const_cast<bns_b200_t *>(ctx_);
    if(!ctx || !out) return BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    unsigned long long h[2] = {0, 0};
    CK(cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    ctx->stats.n_classified = h[0];
    ctx->stats.n_unclassified = h[1];
    float ms = 0.f;
    if(ctx->timed) {
        if(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->stats.last_kernel_ms = ms;
        else cudaGetLastError();
    }
    *out = ctx->stats;
    return BNS_OK;
}

int bns_b200_stats_reset(bns_b200_t *ctx) {
    if(!ctx) return BNS_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(ctx->d_counters, 0, 2 * sizeof(unsigned long long)));
    ctx->stats = bns_b200_stats{};
    return BNS_OK;
}

int bns_b200_host_alloc(void **ptr, size_t bytes) {
    if(!ptr) return BNS_E_INVAL;
    if(cudaMallocHost(ptr, bytes) != cudaSuccess) { cudaGetLastError(); *ptr = nullptr; return BNS_E_NOMEM; }
    return BNS_OK;
}
int bns_b200_host_free(void *ptr) {
    if(ptr && cudaFreeHost(ptr) != cudaSuccess) { cudaGetLastError(); return BNS_E_CUDA; }
    return BNS_OK;
}

int bns_b200_bench_gather(bns_b200_t *ctx, uint64_t n_loads, uint64_t seed, double *ms_out) {
    if(!ctx || !ms_out) return BNS_E_INVAL;
    if(!ctx->d_slots) return ctx->fail(BNS_E_STATE, "no table loaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].st;
    CK(cudaEventRecord(ctx->ev0, st));
    CK(launch_gather(ctx->n_sm * 8, st, ctx->d_slots, ctx->bucket_bits, n_loads, seed, ctx->d_counters + 7));
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    ++ctx->stats.kernel_launches;
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *ms_out = ms;
    return BNS_OK;
}

}  // extern "C"
