// bns_kernels.cu -- hand-written sm_100a kernels of the classify hot path.
//
//   bns_classify_kernel<FAM,TAXA>  warp per record: stage the read tile in shared memory (2-bit packed), generate the
//                             k-mer / minimizer stream of Encoder::for_each (include/bonsai/encoder.h:416-442 and the mode
//                             bodies :211-353), probe the device table (kh_get, khash64.h:250-263), keep the per-record
//                             distinct-taxon counts (linear::counter, linear.h:229) and finish with resolve_tree
//                             (util.h:831-869) -- all in the same warp, nothing round-trips through HBM.
//   bns_encode_kernel<FAM>    the same front end writing the stream out (Encoder API parity surface)
//   bns_insert_kernel         builds the bucketised open-addressed table with 64-bit CAS
//   bns_lookup_kernel         kh_get + kh_val over a key batch;  bns_sectors_kernel counts sectors per probe
//   bns_resolve_kernel        resolve_tree over explicit (taxid,count) lists
//   bns_gather_kernel         random 32-byte-sector gather (the measured random-access ceiling)
//
// FAM is a template parameter so that the default classify mode (FAM_U: unspaced, unwindowed, what `bonsai classify`
// always runs, bin/bonsai.cpp:152) compiles to a lean kernel without the window / spaced-seed machinery.
#include "bns_device.cuh"
#include "bns_kernels.h"
#include <algorithm>

namespace bns {

// ---------------------------------------------------------------------------------------------
// per-warp shared memory carve-up (dynamic)
// ---------------------------------------------------------------------------------------------
constexpr int NWORDS = (TILE + CMAX + 16) / 16 + 3;     // 2-bit code words / invalid-bit words per tile
#ifndef BNS_CLASSIFY_MIN_CTAS
#define BNS_CLASSIFY_MIN_CTAS 3
#endif
constexpr int VI_CAP = 512;                              // val_info records staged in shared memory by TMA

struct WarpSmem {
    u32 *codes;      // [NWORDS] 16 bases per word, first base in the top bits
    u32 *bad;        // [NWORDS] 16 invalid-bits per word (low half), first base in bit 15
    ulonglong2 *raw; // [ring_cap] window ring: (element, score) pairs, the last W-1 carried across tiles
    ulonglong2 *work;// [ring_cap] sliding minima over 2^j elements (sparse-table levels, rebuilt per tile)
    u32 *ids;        // [AGG_CAP] distinct value ids      (ClassifySink)
    u32 *cnt;        // [AGG_CAP] their counts
    u32 *tin;        // [AGG_CAP] scratch for resolve
    u32 *tout;       // [AGG_CAP]
};

__host__ __device__ inline size_t warp_smem_bytes(u32 ring_cap, bool classify) {
    size_t b = 2 * NWORDS * sizeof(u32);
    b = (b + 15) & ~size_t(15);
    b += 2 * (size_t)ring_cap * sizeof(ulonglong2);
    if(classify) b += 4 * AGG_CAP * sizeof(u32);
    return (b + 15) & ~size_t(15);
}
__device__ inline WarpSmem carve(unsigned char *base, u32 ring_cap) {
    WarpSmem s;
    s.codes = (u32 *)base;
    s.bad = s.codes + NWORDS;
    size_t off = 2 * NWORDS * sizeof(u32);
    off = (off + 15) & ~size_t(15);
    s.raw = (ulonglong2 *)(base + off);
    s.work = s.raw + ring_cap;
    off += 2 * (size_t)ring_cap * sizeof(ulonglong2);
    s.ids = (u32 *)(base + off);
    s.cnt = s.ids + AGG_CAP;
    s.tin = s.cnt + AGG_CAP;
    s.tout = s.tin + AGG_CAP;
    return s;
}

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }

// exclusive prefix sum of a small per-lane count + warp total
__device__ __forceinline__ u32 warp_excl_scan(u32 v, u32 lane, u32 &total) {
    u32 x = v;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
        const u32 y = __shfl_up_sync(FULL, x, d);
        if(lane >= (u32)d) x += y;
    }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}

// ---------------------------------------------------------------------------------------------
// tile staging: bases [p0, p0+span) of the sequence -> shared memory. Returns the coordinate shift (the
// tile's first base sits at coordinate `shift` because loads are 16-byte aligned) and whether any staged
// base is invalid. One LDG.128 per lane covers the whole tile (<= 398 bytes).
// ---------------------------------------------------------------------------------------------
// The 16-byte block lane `lane` stages for a tile starting at `a0` with `nbases` bases in range (zero outside).
__device__ __forceinline__ uint4 load_tile_block(const char *a0, u32 nbases, const char *buf_end, u32 lane) {
    const u32 shift = (u32)((uintptr_t)a0 & 15u);
    const u32 nblk = (shift + nbases + 15) >> 4;
    const char *p = a0 - shift + 16 * lane;
    uint4 v = make_uint4(0, 0, 0, 0);
    if(lane < nblk && p < buf_end) v = __ldg(reinterpret_cast<const uint4 *>(p));
    return v;
}
__device__ __forceinline__ u32 stage_tile(const EncParams &cP, const WarpSmem &S, const char *seq, u64 p0, u64 L, u32 span,
                                          const char *buf_end, u32 lane, bool &any_invalid, u32 &t_carry,
                                          const uint4 *pre = nullptr) {
    const char *a0 = seq + p0;
    const u32 shift = (u32)((uintptr_t)a0 & 15u);
    const char *aligned = a0 - shift;
    const u64 remaining = L - p0;
    const u32 nbases = (u32)(remaining < span ? remaining : span);      // bases of this sequence in the tile
    const u32 hi = shift + nbases;                                      // valid coordinates: [shift, hi)
    const u32 nblk = (hi + 15) >> 4;
    u32 codes = 0, bad = 0, tmask = 0;
    uint4 v = make_uint4(0, 0, 0, 0);
    u32 suspicious = 0;
    if(lane < nblk) {
        if(pre) v = *pre;                                   // this tile was requested while the previous record ran
        else { const char *p = aligned + 16 * lane; if(p < buf_end) v = __ldg(reinterpret_cast<const uint4 *>(p)); }
        suspicious = pack16_fast(v, codes);
    }
    // common case: every staged byte is one of ACGTacgt -> no invalid-mask work at all
    any_invalid = false;
    if(__any_sync(FULL, suspicious != 0) || cP.t_restart) {
        if(lane < nblk) {
            pack16(v, codes, bad, tmask);
            // bases outside [shift, hi) belong to neighbours (or to nobody): never read by a k-mer of this sequence
            const int before = (int)shift - (int)(16 * lane);
            const int after = (int)hi - (int)(16 * lane);
            u32 keep = 0xffffu;
            if(before > 0) keep &= (before >= 16) ? 0u : (0xffffu >> before);
            if(after < 16) keep &= (after <= 0) ? 0u : ~(0xffffu >> after);
            bad &= keep;
            tmask &= keep;
        }
        if(cP.t_restart) {
            // for_each_uncanon_unspaced_windowed tests the OR-ed 64-bit word against ~0 BEFORE masking (encoder.h:283):
            // with k >= 31 that word holds 32 bases, so the 32nd, 64th, ... consecutive T of a run restarts the rolling
            // state exactly like an invalid base. Mark those bases invalid. e = T-run length entering each word.
            u32 e = t_carry, e_at_cx = 0;
            const u32 cx = shift + TILE - 1;                     // the next tile starts right after this coordinate
            for(u32 wd = 0; wd < nblk; ++wd) {
                const u32 tm = __shfl_sync(FULL, tmask, wd);
                const u32 first = wd == 0 ? shift : 0u;
                if(wd == (cx >> 4)) e_at_cx = e;
                const u32 m = tm << (16 + first);                // first in-range base of the word at bit 31
                u32 lead = __clz(~m);
                lead = lead < 16 - first ? lead : 16 - first;
                if(lead) {
                    const u32 i = 31u - (e & 31u);
                    if(i < lead && lane == wd) bad |= 0x8000u >> (first + i);
                }
                if(lead == 16 - first) e += lead;
                else e = __ffs(~tm) - 1;                          // trailing T of the word start a new run
            }
            if((cx >> 4) < nblk) {                                // run length ending at coordinate cx
                const u32 wd = cx >> 4, first = wd == 0 ? shift : 0u;
                const u32 tm = __shfl_sync(FULL, tmask, wd);
                const u32 upto = (tm >> (15 - (cx & 15u)));       // bit 0 = coordinate cx, bit j = cx - j
                const u32 nbits = (cx & 15u) + 1 - first;         // in-range bases of the word up to cx
                u32 run = __ffs(~upto) - 1;                       // consecutive T ending at cx
                if(run >= nbits) run = nbits + e_at_cx;
                t_carry = run;
            } else t_carry = 0;
        }
        any_invalid = __any_sync(FULL, bad != 0);
        if(lane < NWORDS) S.bad[lane] = bad;
    }
    if(lane < NWORDS) S.codes[lane] = codes;
    __syncwarp();
    return shift;
}

// the comb's k bases starting at coordinate q (Encoder::kmer, encoder.h:547-592); inval = some base is not ACGT
__device__ __forceinline__ u64 gather_kmer(const EncParams &cP, const WarpSmem &S, u32 q, bool check_bad, bool &inval) {
    u64 x = 0;
    inval = false;
    const u32 ns = cP.n_seg;
    for(u32 s = 0; s < ns; ++s) {
        const u32 off = cP.seg_off[s], len = cP.seg_len[s];
        const u64 bits = extract_bases(S.codes, q + off, len);
        x = (len == 32) ? bits : ((x << (2 * len)) | bits);
        if(check_bad) inval |= any_bad(S.bad, q + off, len) != 0;
    }
    return x;
}

// score of one window element (scorer_(kmer, data), encoder.h:616-628 ; :337 for the rolling entropy)
__device__ __forceinline__ u64 score_of(const EncParams &cP, u64 x) {
    const u32 sk = cP.score_kind;
    if(sk == SC_LEX) return lex_score(x);
    if(sk == SC_ENT_NOTFULL)                       // ent_score with CircusEnt::NOT_FULL, encoder.h:55-58, entropy.h:45
        return cast_u64(__ull2double_rn(x) / (-1. + 1e-4), cP.cast_wrap);
    // rolling entropy over the k bases of x: sum in ska::flat_hash_map iteration order A, G, C, T
    // (fibonacci-hashed slots 0,1,4,6 of its 8-slot table, hll/include/flat_hash_map/flat_hash_map.hpp:1268)
    const u32 k = cP.k;
    const u64 m = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    const u64 lo = x & 0x5555555555555555ull, hi = (x >> 1) & 0x5555555555555555ull;
    const u32 nT = __popcll(hi & lo), nG = __popcll(hi & ~lo), nC = __popcll(~hi & lo & m);
    const u32 nA = k - nT - nG - nC;
    double h = 0.;
    if(nA) h = h + cP.plogp[nA];
    if(nG) h = h + cP.plogp[nG];
    if(nC) h = h + cP.plogp[nC];
    if(nT) h = h + cP.plogp[nT];
    return cast_u64(__ull2double_rn(x) / (h + .001), cP.cast_wrap);
}

// ---------------------------------------------------------------------------------------------
// table probe
// ---------------------------------------------------------------------------------------------
// exact test of the (64-b)-bit remainder and the 3-bit displacement of four slots against `tag`; returns the value id
__device__ __forceinline__ u32 match4(const TableView &T, u64 tag, u64 s0, u64 s1, u64 s2, u64 s3) {
    const u32 th = (u32)(tag >> 32), tl = (u32)tag, hm = ~0u << T.tag_shift;      // tag_shift <= 28
    const u32 d0 = ((u32)(s0 >> 32) ^ th) | (((u32)s0 ^ tl) & hm);
    const u32 d1 = ((u32)(s1 >> 32) ^ th) | (((u32)s1 ^ tl) & hm);
    const u32 d2 = ((u32)(s2 >> 32) ^ th) | (((u32)s2 ^ tl) & hm);
    const u32 d3 = ((u32)(s3 >> 32) ^ th) | (((u32)s3 ^ tl) & hm);
    u32 v = VAL_MISS;
    if(d0 == 0) v = (u32)s0;
    if(d1 == 0) v = (u32)s1;
    if(d2 == 0) v = (u32)s2;
    if(d3 == 0) v = (u32)s3;
    return (d0 && d1 && d2 && d3) ? VAL_MISS : (v & T.val_mask);
}
// One probe of a key covers T.fmt.unit_slots() slots: one 32-byte bucket (LAYOUT_HASH) or a 64-byte unit of two
// (LAYOUT_MINIMIZER). Exact test of every slot of the probe at `bucket`. flagw: the low word of the probe's first slot (the
// overflow flags); last_free: the probe's last slot is empty.
__device__ __forceinline__ u32 match_probe(const TableView &T, u64 bucket, u64 tag, u32 &flagw, bool &last_free) {
    u64 a, b, c, d;
    ld_bucket(T.slots + (bucket << 2), a, b, c, d);
    flagw = (u32)a;
    last_free = d == ~0ull;
    u32 v = match4(T, tag, a, b, c, d);
    if(T.fmt.layout == LAYOUT_MINIMIZER) {
        ld_bucket(T.slots + (bucket << 2) + 4, a, b, c, d);
        last_free = d == ~0ull;
        const u32 v2 = match4(T, tag, a, b, c, d);
        if(v == VAL_MISS) v = v2;
    }
    return v;
}
// probes after the home probe (rare: the home probe was full when the table was built). exhausted: the chain was full to its
// end -- a LAYOUT_MINIMIZER key may then sit in the stash.
__device__ __noinline__ u32 probe_displaced(const TableView T, u64 home, u64 tag, bool *exhausted = nullptr) {
    if(exhausted) *exhausted = false;
    for(u32 d = 1; d <= T.fmt.max_disp(); ++d) {
        u32 fw; bool last_free;
        const u32 v = match_probe(T, probe_bucket(T.fmt.layout, home, d, T.fmt.b), tag | ((u64)d << T.tag_shift), fw, last_free);
        if(v != VAL_MISS || last_free) return v;                   // a probe with a free slot ends the run
    }
    if(exhausted) *exhausted = true;
    return VAL_MISS;
}
// Home-bucket test relying on the build-time invariant that no two entries of a bucket share their upper 32 bits
// (bns_insert_kernel rejects such a pair and the table is rebuilt one size up): pick the first slot whose high word
// equals the tag's, then verify the remaining remainder + displacement bits of that one candidate exactly.
__device__ __forceinline__ u32 match4_home(const TableView &T, u64 tag, u64 s0, u64 s1, u64 s2, u64 s3) {
    const u32 th = (u32)(tag >> 32), tl = (u32)tag, hm = ~0u << T.tag_shift;
    const bool m0 = (u32)(s0 >> 32) == th, m1 = (u32)(s1 >> 32) == th, m2 = (u32)(s2 >> 32) == th, m3 = (u32)(s3 >> 32) == th;
    const u32 cand = m0 ? (u32)s0 : m1 ? (u32)s1 : m2 ? (u32)s2 : (u32)s3;       // slots fill in order: first match wins
    const bool ok = (m0 | m1 | m2 | m3) && (((cand ^ tl) & hm) == 0);            // an empty slot (disp 15) never passes
    return ok ? (cand & T.val_mask) : VAL_MISS;
}
// the stash of a LAYOUT_MINIMIZER table: an ordinary LAYOUT_HASH probe
__device__ __noinline__ u32 probe_stash(const TableView T, u64 key) {
    const TableView S = stash_view(T);
    bool possible;
    const TableHash h = table_hash(S.fmt, key, possible);
    u64 a, b, c, d;
    ld_bucket(S.slots + (h.home << 2), a, b, c, d);
    u32 v = match4_home(S, h.tag, a, b, c, d);
    if(v == VAL_MISS && !(((u32)a >> (S.flag_shift + (h.fsel & S.flag_mask))) & 1u)) v = probe_displaced(S, h.home, h.tag);
    return v;
}
// kh_get + kh_val for one key (generic kernels: lookup, the stream kernels of LAYOUT_MINIMIZER tables)
__device__ __forceinline__ u32 probe_key(const TableView &T, u64 key) {
    bool possible;
    const TableHash h = table_hash(T.fmt, key, possible);
    if(!possible) return VAL_MISS;
    u32 fw; bool last_free;
    const u32 v = match_probe(T, probe_bucket(T.fmt.layout, h.home, 0, T.fmt.b), h.tag, fw, last_free);
    // The overflow mark is a CLEARED bit in slot 0 of a full home probe (an empty slot is all ones, so an empty or
    // part-filled one reads "no overflow"): a miss costs one probe unless a key homed there was displaced.
    if(v == VAL_MISS && !((fw >> (T.flag_shift + (h.fsel & T.flag_mask))) & 1u)) {
        bool exhausted;
        const u32 v2 = probe_displaced(T, h.home, h.tag, &exhausted);
        return (v2 == VAL_MISS && exhausted && T.stash) ? probe_stash(T, key) : v2;
    }
    return v;
}
// kh_get + kh_val for PPL keys per lane: all home-bucket sectors are requested before any is inspected. Lanes / slots
// without a live k-mer probe anyway (their result is masked by the caller): no predication, no register init.
__device__ __forceinline__ void probe4(const TableView &T, const u64 (&x)[PPL], u32 (&val)[PPL]) {
    if(T.fmt.layout != LAYOUT_HASH) {                             // 64-byte probes: key by key (the lean kernel has the fast path)
#pragma unroll
        for(int i = 0; i < PPL; ++i) val[i] = probe_key(T, x[i]);
        return;
    }
    TableHash h[PPL];
    u64 s[PPL][4];
#pragma unroll
    for(int i = 0; i < PPL; ++i) {
        bool possible;
        h[i] = table_hash(T.fmt, x[i], possible);
        ld_bucket(T.slots + (h[i].home << 2), s[i][0], s[i][1], s[i][2], s[i][3]);
    }
    u32 more = 0;
#pragma unroll
    for(int i = 0; i < PPL; ++i) {
        val[i] = match4_home(T, h[i].tag, s[i][0], s[i][1], s[i][2], s[i][3]);
        if(val[i] == VAL_MISS && !(((u32)s[i][0] >> (T.flag_shift + (h[i].fsel & T.flag_mask))) & 1u)) more |= 1u << i;
    }
    if(more) {                                                    // rare
#pragma unroll
        for(int i = 0; i < PPL; ++i) if(more >> i & 1u) val[i] = probe_displaced(T, h[i].home, h[i].tag);
    }
}

// ---------------------------------------------------------------------------------------------
// sinks
// ---------------------------------------------------------------------------------------------
struct StoreSink {
    u64 *out;         // this record's output window
    u64 cap, n;       // warp-uniform
    __device__ __forceinline__ void begin(u64 *o, u64 c) { out = o; cap = c; n = 0; }
    __device__ __forceinline__ void consume(const WarpSmem &, const u64 (&x)[PPL], u32 mask, u32 lane) {
        u32 total;
        const u32 ex = warp_excl_scan(__popc(mask), lane, total);
        u64 idx = n + ex;
#pragma unroll
        for(int i = 0; i < PPL; ++i)
            if(mask >> i & 1u) { if(idx < cap) out[idx] = x[i]; ++idx; }
        n += total;
    }
};

template <bool TAXA>
struct ClassifySink {
    TableView T;
    const uint4 *vi;  // val_info: shared-memory copy (TMA-staged) when it fits, else global
    u32 *taxa_out;    // ordered hit list of this record (raw taxids) when TAXA
    u32 n_distinct, n_hit, n_miss, overflow;   // warp-uniform
    u32 cap = AGG_CAP;                         // entries the distinct-taxon lists hold (AGG_CAP in shared memory; more in the overflow pass)
    u64 taxa_cap = 0;                          // entries of this record's hit-list window (the caller sized it)
    u32 taxa_short = 0;                        // ... and whether it was too small

    __device__ __forceinline__ void begin(u32 *taxa, u64 tcap = ~0ull) {
        taxa_out = taxa; taxa_cap = tcap; n_distinct = n_hit = n_miss = overflow = taxa_short = 0;
    }

    // linear::counter::add of `total` hits of value id v
    __device__ __forceinline__ void add(const WarpSmem &S, u32 v, u32 total, u32 lane) {
        int found = -1;
        for(u32 base = 0; base < n_distinct; base += 32) {
            const u32 idx = base + lane;
            const u32 bm = __ballot_sync(FULL, idx < n_distinct && S.ids[idx] == v);
            if(bm) { found = (int)(base + __ffs(bm) - 1); break; }
        }
        if(found < 0) {
            if(n_distinct < cap) {
                if(lane == 0) { S.ids[n_distinct] = v; S.cnt[n_distinct] = total; }
                ++n_distinct;
            } else overflow = 1;
        } else if(lane == 0) S.cnt[found] += total;
        __syncwarp();
    }

    __device__ __forceinline__ void consume(const WarpSmem &S, const u64 (&x)[PPL], u32 mask, u32 lane) {
        if(!__any_sync(FULL, mask != 0)) return;
        u32 val[PPL];
        probe4(T, x, val);
        u32 todo = 0;
#pragma unroll
        for(int i = 0; i < PPL; ++i) if(val[i] != VAL_MISS) todo |= 1u << i;
        todo &= mask;
        const u32 both = __reduce_add_sync(FULL, __popc(mask) | (__popc(todo) << 16));   // emitted | hits << 16
        const u32 hits_total = both >> 16;
        if(TAXA) {
            u32 dummy;
            u32 idx = n_hit + warp_excl_scan(__popc(todo), lane, dummy);
#pragma unroll
            for(int i = 0; i < PPL; ++i) if(todo >> i & 1u) { if(idx < taxa_cap) taxa_out[idx] = vi[val[i]].w; ++idx; }
            if((u64)n_hit + hits_total > taxa_cap) taxa_short = 1;
        }
        n_hit += hits_total;
        n_miss += (both & 0xffffu) - hits_total;
        // fold the hits into the per-record distinct list, one distinct value per iteration
        for(;;) {
            const u32 bal = __ballot_sync(FULL, todo != 0);
            if(!bal) break;
            const u32 leader = __ffs(bal) - 1;
            u32 fv = 0;
#pragma unroll
            for(int i = PPL - 1; i >= 0; --i) if(todo >> i & 1u) fv = val[i];
            const u32 v = __shfl_sync(FULL, fv, leader);
            u32 c = 0;
#pragma unroll
            for(int i = 0; i < PPL; ++i) if((todo >> i & 1u) && val[i] == v) { ++c; todo &= ~(1u << i); }
            add(S, v, __reduce_add_sync(FULL, c), lane);
        }
    }

    // resolve_tree (util.h:831-869): score(t) = sum of u16 counts over t's root path; unique max wins, ties -> lca of
    // all tied taxa. Root paths are tested with Euler-tour intervals instead of walking the parent map.
    __device__ __forceinline__ u32 resolve(const WarpSmem &S, const TaxView &X, u32 lane) const {
        const u32 n = n_distinct;
        if(n == 0) return 0;
        if(n == 1) return vi[S.ids[0]].w;
        for(u32 i = lane; i < n; i += 32) {
            const uint4 inf = vi[S.ids[i]];
            S.tin[i] = inf.x; S.tout[i] = inf.y;
        }
        __syncwarp();
        u32 best = 0, sc0 = 0;                                // sc0: this lane's score in the first (usually only) round of 32
        for(u32 base = 0; base < n; base += 32) {
            const u32 i = base + lane;
            u32 sc = 0;
            if(i < n) {
                const u32 ti = S.tin[i];
                for(u32 j = 0; j < n; ++j)
                    if(S.tin[j] <= ti && ti < S.tout[j]) sc += S.cnt[j] & 0xffffu;
            }
            if(base == 0) sc0 = sc;
            best = max(best, __reduce_max_sync(FULL, sc));
        }
        // ties (score == best), folded in list order
        u32 node = 0, ntied = 0, first_id = 0;
        for(u32 base = 0; base < n; base += 32) {
            const u32 i = base + lane;
            u32 sc = sc0;
            if(base != 0) {                                    // lists of more than 32 taxa: the later rounds' scores again
                sc = 0;
                if(i < n) {
                    const u32 ti = S.tin[i];
                    for(u32 j = 0; j < n; ++j)
                        if(S.tin[j] <= ti && ti < S.tout[j]) sc += S.cnt[j] & 0xffffu;
                }
            }
            u32 tied = __ballot_sync(FULL, i < n && sc == best);
            while(tied) {
                const u32 l = __ffs(tied) - 1;
                tied &= tied - 1;
                const u32 id = S.ids[base + l];
                const uint4 inf = vi[id];                     // uniform address
                if(ntied == 0) { node = inf.z; first_id = id; }
                else {
                    // lca(node, b): climb from `node` until its interval covers b (util.h:634-663)
                    const u32 tb = inf.x;
                    u32 a = node;
                    while(a) {
                        const uint4 na = X.node_info[a];
                        if(na.x <= tb && tb < na.y) break;
                        a = na.z;
                    }
                    node = a ? a : X.node_of_one;
                }
                ++ntied;
            }
        }
        if(ntied == 1) return vi[first_id].w;
        return X.node_info[node].w;
    }
};

// value -> dense id by binary search in the sorted distinct-value list
__device__ __forceinline__ u32 value_id(const u32 *__restrict__ values, u32 n, u32 v) {
    u32 lo = 0, hi = n;
    while(lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if(values[mid] < v) lo = mid + 1; else hi = mid;
    }
    return (lo < n && values[lo] == v) ? lo : VAL_MISS;
}

// Database construction on the device: every k-mer the encoder emits for a genome is inserted with the genome's value
// id, or -- if the key is already present with another value -- merged with lca(tax, new, old), which is what
// fill_set_genome + update_lca_map do (feature_min.h:68-83,205-228). lca is idempotent, commutative and associative on
// a well-formed taxonomy, so concurrent CAS merges converge to the reference's sequential result.
struct BuildSink {
    u64 *slots;
    u32 b, tag_shift, val_mask, flag_shift, flag_mask;
    TableFmt fmt;
    u32 vid;                           // value id of the genome being added
    const uint4 *val_info, *node_info; // Euler intervals: {tin, tout, node, taxid} / {tin, tout, parent node, taxid}
    const u32 *values;
    u32 n_values, node_of_one;
    unsigned long long *stats;         // [0] no room  [1] displaced  [2] bad taxonomy / value (lca outside the dictionary, upper-word clash)  [3] new keys
    u32 n_new, n_fail, n_bad;          // per lane

    __device__ __forceinline__ u32 lca_id(u32 a_id, u32 b_id) const {     // lca(), util.h:634-663, on value ids
        if(a_id == b_id) return a_id;
        u32 a = val_info[a_id].z;
        const u32 tb = val_info[b_id].x;
        while(a) {
            const uint4 na = node_info[a];
            if(na.x <= tb && tb < na.y) break;
            a = na.z;
        }
        if(!a) a = node_of_one;
        return value_id(values, n_values, node_info[a].w);
    }
    __device__ __forceinline__ void insert(u64 key) {
        bool possible;
        const TableHash th = table_hash(fmt, key, possible);
        if(!possible || fmt.layout != LAYOUT_HASH) { ++n_fail; return; }   // databases are built in LAYOUT_HASH and re-homed (build_finish)
        const u64 home = th.home, tag = th.tag;
        for(u32 d = 0; d <= fmt.max_disp(); ++d) {
            u64 *bk = slots + (probe_bucket(fmt.layout, home, d, fmt.b) << 2);
            const u64 entry = tag | ((u64)d << tag_shift) | (((1ull << tag_shift) - 1) & ~(u64)val_mask) | vid;
            for(int s = 0; s < 4; ++s) {
                u64 cur = bk[s];
                if(cur == ~0ull) {
                    cur = atomicCAS((unsigned long long *)&bk[s], ~0ull, (unsigned long long)entry);
                    if(cur == ~0ull) { ++n_new; return; }
                }
                if(((cur ^ entry) >> tag_shift) == 0) {                  // key present: update_lca_map's merge branch
                    for(;;) {
                        const u32 old = (u32)cur & val_mask;
                        if(old == vid) return;
                        const u32 m = lca_id(vid, old);
                        if(m == old) return;
                        if(m == VAL_MISS) { ++n_bad; return; }
                        const u64 want = (cur & ~(u64)val_mask) | m;
                        const u64 prev = atomicCAS((unsigned long long *)&bk[s], (unsigned long long)cur, (unsigned long long)want);
                        if(prev == cur) return;
                        cur = prev;                                      // value or overflow mark changed under us: retry
                    }
                }
                if(fmt.layout == LAYOUT_HASH && (u32)(cur >> 32) == (u32)(entry >> 32)) { ++n_fail; return; }   // unique-upper-word invariant: a larger table re-draws the words
            }
            if(d == 0) atomicAnd((unsigned long long *)&bk[0], ~(1ull << (flag_shift + (th.fsel & flag_mask))));
        }
        ++n_fail;
    }
    __device__ __forceinline__ void consume(const WarpSmem &, const u64 (&x)[PPL], u32 mask, u32) {
#pragma unroll
        for(int i = 0; i < PPL; ++i) if(mask >> i & 1u) insert(x[i]);
    }
};

// ---------------------------------------------------------------------------------------------
// window ring (QueueMap, qmap.h:79-96) over the elements a tile produced. `m` new elements were written to
// S.raw[hist .. hist+m) by the caller. Produces the window minima for the new elements this lane owns
// (indices j = PPL*lane + i) and rolls the ring forward. total_before = elements pushed before this tile.
// ---------------------------------------------------------------------------------------------
// (score, element) order of ElScore::operator< (qmap.h:23)
__device__ __forceinline__ ulonglong2 pair_min(ulonglong2 a, ulonglong2 b) {
    return (b.y < a.y || (b.y == a.y && b.x < a.x)) ? b : a;
}
// `m` new (element, score) pairs sit in S.raw[hist .. hist+m). Sliding minima over W elements in O(log W) passes:
// level j holds min over the last 2^j elements; the window of W is the min of two overlapping windows of
// t = 2^floor(log2 W). Only entries whose full window lies inside the buffer are used (g >= W-1 globally).
__device__ __forceinline__ u32 window_outputs(const WarpSmem &S, u32 hist, u32 m, u64 total_before, u32 W, u32 lane,
                                              u64 (&o)[PPL]) {
    const u32 have = hist + m;
    if(m == 0) {                                       // warp-uniform: nothing new, nothing to emit
#pragma unroll
        for(int i = 0; i < PPL; ++i) o[i] = KMER_NONE;
        return 0;
    }
    // level 1: raw -> work
    for(u32 g = lane; g < have; g += 32) {
        ulonglong2 v = S.raw[g];
        if(g >= 1) v = pair_min(v, S.raw[g - 1]);
        S.work[g] = v;
    }
    __syncwarp();
    u32 t = 2;
    for(; t * 2 <= W; t *= 2) {                       // work[g] = min(work[g], work[g - t]) : window 2t, in place
        // Rounds of 8 x 32 entries, from the top of the buffer downwards: a round reads only indices <= its own, so the
        // rounds below it are still at the old level; inside a round every read precedes every write.
        for(int base = (int)((have - 1) / 256) * 256; base >= 0; base -= 256) {
            ulonglong2 v[8];
#pragma unroll
            for(int r = 0; r < 8; ++r) {
                const u32 g = (u32)base + lane + 32 * r;
                if(g < have) { v[r] = S.work[g]; if(g >= t) v[r] = pair_min(v[r], S.work[g - t]); }
            }
            __syncwarp();
#pragma unroll
            for(int r = 0; r < 8; ++r) {
                const u32 g = (u32)base + lane + 32 * r;
                if(g < have) S.work[g] = v[r];
            }
            __syncwarp();
        }
    }
    u32 mask = 0;
    const u32 back = W - t;                            // second window starts `back` elements earlier (0 if W == t)
#pragma unroll
    for(int i = 0; i < PPL; ++i) {
        const u32 j = PPL * lane + i;
        o[i] = KMER_NONE;
        if(j < m && total_before + j + 1 >= W) {
            const u32 g = hist + j;
            ulonglong2 v = S.work[g];
            if(back) v = pair_min(v, S.work[g - back]);
            o[i] = v.x;
            mask |= 1u << i;
        }
    }
    __syncwarp();
    return mask;
}
__device__ __forceinline__ u32 roll_ring(const WarpSmem &S, u32 hist, u32 m, u32 W, u32 lane) {
    const u32 have = hist + m;
    const u32 keep = have < W - 1 ? have : W - 1;
    const u32 src = have - keep;
    if(src) {
        for(u32 base = 0; base < keep; base += 32) {
            const u32 i = base + lane;
            ulonglong2 e = make_ulonglong2(0, 0);
            if(i < keep) e = S.raw[src + i];
            __syncwarp();
            if(i < keep) S.raw[i] = e;
            __syncwarp();
        }
    }
    return keep;
}

// ---------------------------------------------------------------------------------------------
// FAM_U fast path: unspaced, unwindowed (for_each_uncanon_unspaced_unwindowed, encoder.h:240-272, + canonical wrapper
// :218-232). The lane's four consecutive k-mers and their reverse complements come out of ONE 96-bit window of the
// staged tile: forward k-mer i = window bits [2i, 2i+2k), reverse complement i = low 2k bits of (rc(window) >> 2i).
// ---------------------------------------------------------------------------------------------
template <class Sink>
__device__ __forceinline__ void encode_sequence_u(const EncParams &cP, const WarpSmem &S, const char *seq, u64 L,
                                                  const char *buf_end, Sink &sink, u32 lane, const uint4 *pre = nullptr) {
    const u32 k = cP.k;
    if(L < k) return;                                            // has_next_kmer(), encoder.h:418,594
    const u64 npos = L - k + 1;
    const u32 span = TILE + k - 1;
    const u32 down = 64 - 2 * k;
    const bool canon = cP.canon_elem != 0;
    u32 t_carry = 0;
    for(u64 p0 = 0; p0 < npos; p0 += TILE) {
        bool any_invalid;
        const u32 shift = stage_tile(cP, S, seq, p0, L, span, buf_end, lane, any_invalid, t_carry, p0 == 0 ? pre : nullptr);
        const u32 q0 = shift + PPL * lane;
        const u64 left = npos - p0;
        const u32 nlive = left > (u64)PPL * lane ? (u32)min((u64)PPL, left - (u64)PPL * lane) : 0u;
        u32 mask = (1u << nlive) - 1;
        u64 x[PPL];
        {
            const u32 wi = q0 >> 4, s = (q0 & 15u) * 2u;
            const u32 w0 = S.codes[wi], w1 = S.codes[wi + 1], w2 = S.codes[wi + 2], w3 = S.codes[wi + 3];
            const u32 A = __funnelshift_l(w1, w0, s), B = __funnelshift_l(w2, w1, s), C = __funnelshift_l(w3, w2, s);
            u32 R0 = 0, R1 = 0, R2 = 0;
            if(canon) {
                // rc of the 48-base window: reverse the bit string, swap the two bits of each pair back, complement
                R0 = __brev(C); R1 = __brev(B); R2 = __brev(A);
                R0 = ~(((R0 >> 1) & 0x55555555u) | ((R0 & 0x55555555u) << 1));
                R1 = ~(((R1 >> 1) & 0x55555555u) | ((R1 & 0x55555555u) << 1));
                R2 = ~(((R2 >> 1) & 0x55555555u) | ((R2 & 0x55555555u) << 1));
            }
            const u64 kmask = ~0ull >> down;
#pragma unroll
            for(int i = 0; i < PPL; ++i) {
                const u32 hi = __funnelshift_l(B, A, 2 * i), lo = __funnelshift_l(C, B, 2 * i);
                u64 f = (((u64)hi << 32) | lo) >> down;
                if(canon) {
                    const u32 rl = __funnelshift_r(R2, R1, 2 * i), rh = __funnelshift_r(R1, R0, 2 * i);
                    const u64 r = (((u64)rh << 32) | rl) & kmask;
                    f = f < r ? f : r;
                }
                x[i] = f;
            }
        }
        if(any_invalid) {                                        // rare: drop the windows that cover an invalid base
#pragma unroll
            for(int i = 0; i < PPL; ++i) if((mask >> i & 1u) && any_bad(S.bad, q0 + i, k)) mask &= ~(1u << i);
        }
        __syncwarp();
        sink.consume(S, x, mask, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// generic path: FAM_K (kmer(pos)-based, windowed/spaced) and FAM_R (rolling + window + tail flush)
// ---------------------------------------------------------------------------------------------
template <int FAM, class Sink>
__device__ __forceinline__ void encode_sequence_g(const EncParams &cP, const WarpSmem &S, const char *seq, u64 L,
                                                  const char *buf_end, Sink &sink, u32 lane) {
    const u32 k = cP.k, c = cP.c, W = cP.W;
    if(L < c) return;                                           // has_next_kmer(), encoder.h:418,594
    const u64 npos = L - c + 1;
    const u32 span = TILE + c - 1;
    u32 hist = 0;                                               // ring entries carried from earlier tiles
    u64 total = 0;                                              // elements pushed so far (QueueMap list_)
    u32 t_carry = 0;                                            // T-run length entering the tile (t_restart only)
    for(u64 p0 = 0; p0 < npos; p0 += TILE) {
        bool any_invalid;
        const u32 shift = stage_tile(cP, S, seq, p0, L, span, buf_end, lane, any_invalid, t_carry);
        u64 x[PPL];
        u32 live = 0;                                           // positions of this lane inside the sequence
        u32 okm = 0;                                            // ... whose comb covers only valid bases
#pragma unroll
        for(int i = 0; i < PPL; ++i) {
            const u64 p = p0 + PPL * lane + i;
            x[i] = KMER_NONE;
            if(p < npos) {
                bool inval;
                live |= 1u << i;
                x[i] = gather_kmer(cP, S, shift + PPL * lane + i, any_invalid, inval);
                if(!inval) okm |= 1u << i;
            }
        }
        // element set of this tile
        u32 emask;
        if(FAM == FAM_K) {
            emask = live;                                       // every position pushes (invalid -> ~0, or 0 if canon)
#pragma unroll
            for(int i = 0; i < PPL; ++i) if(!(okm >> i & 1u)) x[i] = KMER_NONE;
            if(cP.canon_elem) {
#pragma unroll
                for(int i = 0; i < PPL; ++i) if(live >> i & 1u) x[i] = canonical(x[i], k);
            }
        } else emask = okm;                                     // FAM_R: only valid k-mers push
        if(W == 1) {                                            // window of one: the element itself
            u32 mask = emask;
            if(cP.filter_none) {
#pragma unroll
                for(int i = 0; i < PPL; ++i) if(x[i] == KMER_NONE) mask &= ~(1u << i);
            }
            if(cP.canon_emit) {
#pragma unroll
                for(int i = 0; i < PPL; ++i) if(mask >> i & 1u) x[i] = canonical(x[i], k);
            }
            __syncwarp();
            sink.consume(S, x, mask, lane);
            total += __reduce_add_sync(FULL, __popc(emask));
            continue;
        }
        u32 m;
        const u32 ex = warp_excl_scan(__popc(emask), lane, m);
        {
            u32 idx = hist + ex;
#pragma unroll
            for(int i = 0; i < PPL; ++i)
                if(emask >> i & 1u) { S.raw[idx] = make_ulonglong2(x[i], score_of(cP, x[i])); ++idx; }
        }
        __syncwarp();
        u64 o[PPL];
        u32 mask = window_outputs(S, hist, m, total, W, lane, o);
#pragma unroll
        for(int i = 0; i < PPL; ++i) {
            if(cP.filter_none && o[i] == KMER_NONE) mask &= ~(1u << i);
            if(cP.canon_emit && (mask >> i & 1u)) o[i] = canonical(o[i], k);
        }
        hist = roll_ring(S, hist, m, W, lane);
        total += m;
        sink.consume(S, o, mask, lane);
    }
    // tail flush: a queue that never filled emits its minimum once (encoder.h:304-305,343-344)
    if(cP.tail_flush && W > 1 && total > 0 && total < W) {
        ulonglong2 best = S.raw[0];
        for(u32 t = 1; t < (u32)total; ++t) best = pair_min(best, S.raw[t]);
        const u64 be = best.x;
        u64 o[PPL] = {cP.canon_emit ? canonical(be, k) : be, KMER_NONE, KMER_NONE, KMER_NONE};
        __syncwarp();
        sink.consume(S, o, lane == 0 ? 1u : 0u, lane);
    }
    __syncwarp();
}

template <int FAM, class Sink>
__device__ __forceinline__ void encode_sequence(const EncParams &cP, const WarpSmem &S, const char *seq, u64 L,
                                                const char *buf_end, Sink &sink, u32 lane, const uint4 *pre = nullptr) {
    if(FAM == FAM_U) encode_sequence_u(cP, S, seq, L, buf_end, sink, lane, pre);
    else if(FAM == FAM_K || FAM == FAM_R) encode_sequence_g<FAM>(cP, S, seq, L, buf_end, sink, lane);
    // FAM_NONE: the string overload with a spaced seed emits nothing (encoder.h:437-440)
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char g_smem[];

template <int FAM>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
bns_encode_kernel(const __grid_constant__ EncParams P, const char *__restrict__ bases, const u64 *__restrict__ offsets,
                  u64 n_seqs, u64 total_bases, u64 *__restrict__ kmers_out, const u64 *__restrict__ out_offsets,
                  u32 *__restrict__ counts_out, u32 ring_cap, u32 *__restrict__ status) {
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const WarpSmem S = carve(g_smem + wid * warp_smem_bytes(ring_cap, false), ring_cap);
    const u64 nwarps = (u64)gridDim.x * WARPS_PER_CTA;
    StoreSink sink;
    for(u64 r = (u64)blockIdx.x * WARPS_PER_CTA + wid; r < n_seqs; r += nwarps) {
        const u64 b = offsets[r], e = offsets[r + 1];
        const u64 ob = out_offsets[r], oe = out_offsets[r + 1];
        sink.begin(kmers_out + ob, oe - ob);
        encode_sequence<FAM>(P, S, bases + b, e - b, bases + total_bases, sink, lane);
        if(lane == 0) {
            counts_out[r] = (u32)sink.n;
            if(sink.n > sink.cap) atomicOr(status, 1u);
        }
    }
}

// The per-value taxonomy records (val_info: Euler interval, node, taxid) are the only taxonomy data a record with
// <= 1 tied maximum touches. When they fit (n_values <= VI_CAP) each CTA brings them into shared memory with one
// TMA bulk copy (cp.async.bulk -> UBLKCP) completing on an mbarrier, while the warps already stage their first reads.
__device__ __forceinline__ void tma_stage_val_info(uint4 *dst, const uint4 *src, u32 bytes, unsigned long long *mbar) {
    const u32 mb = (u32)__cvta_generic_to_shared(mbar), d = (u32)__cvta_generic_to_shared(dst);
    if(threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if(threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(src), "r"(bytes), "r"(mb) : "memory");
    }
}
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, u32 parity) {
    const u32 mb = (u32)__cvta_generic_to_shared(mbar);
    u32 done = 0;
    while(!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mb), "r"(parity) : "memory");
    }
}

template <int FAM, bool TAXA>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, BNS_CLASSIFY_MIN_CTAS)
bns_classify_kernel(const __grid_constant__ EncParams P, const char *__restrict__ bases, const u64 *__restrict__ offsets,
                    u64 n_records, u32 mates, u64 total_bases, TableView T, TaxView X,
                    u32 *__restrict__ taxon_out, u32 *__restrict__ nhit_out, u32 *__restrict__ nmiss_out,
                    u32 *__restrict__ taxa_out, const u64 *__restrict__ taxa_offsets, u32 *__restrict__ mate1_out, u32 ring_cap,
                    unsigned long long *__restrict__ counters, u32 *__restrict__ status,
                    const u32 *__restrict__ rec_list, const unsigned long long *__restrict__ rec_count,
                    u32 *__restrict__ ovf_idx, unsigned long long *__restrict__ ovf_cnt, u32 *__restrict__ big_scratch, u32 big_cap) {
    // rec_list != nullptr: only the *rec_count records it names (what the first pass left for this one: windowed records of
    // more than one tile, 32-T restarts, records with more than AGG_CAP distinct taxa). ovf_idx != nullptr (first pass): a
    // record whose distinct-taxon list overflows shared memory is appended there instead of failing the call; the pass over
    // rec_list keeps its lists in big_scratch (global memory, big_cap entries per list).
    __shared__ __align__(16) uint4 s_vi[VI_CAP];
    __shared__ __align__(8) unsigned long long s_mbar;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const bool staged = T.n_values > 0 && T.n_values <= (u32)VI_CAP;
    if(staged) tma_stage_val_info(s_vi, X.val_info, T.n_values * (u32)sizeof(uint4), &s_mbar);
    WarpSmem S = carve(g_smem + wid * warp_smem_bytes(ring_cap, true), ring_cap);
    const u64 nwarps = (u64)gridDim.x * WARPS_PER_CTA;
    ClassifySink<TAXA> sink;
    sink.T = T;
    sink.vi = staged ? s_vi : X.val_info;
    if(big_scratch) {
        S.ids = big_scratch + ((size_t)blockIdx.x * WARPS_PER_CTA + wid) * 4 * big_cap;
        S.cnt = S.ids + big_cap; S.tin = S.cnt + big_cap; S.tout = S.tin + big_cap;
        sink.cap = big_cap;
    }
    if(staged) mbar_wait(&s_mbar, 0);
    u32 n_cls = 0, n_uncls = 0;
    // total_bases == ~0: the caller's offsets live on the device only (bns_b200_classify_device); the bound of the
    // 16-byte staging loads is the last offset
    const char *buf_end = bases + (total_bases == ~0ull ? offsets[n_records * mates] : total_bases);
    const u64 r_first = (u64)blockIdx.x * WARPS_PER_CTA + wid;
    if(FAM == FAM_U && mates == 1 && !rec_list) {
        // Software pipeline over this warp's records: the offsets of record r+2*nwarps and the first tile of record
        // r+nwarps are requested before record r is processed, so neither the offset fetch nor the read bytes (both
        // stream from HBM) stall the warp when their turn comes.
        const u32 span = TILE + P.k - 1;
        u64 b0 = 0, e0 = 0, b1 = 0, e1 = 0;
        uint4 v0 = make_uint4(0, 0, 0, 0);
        if(r_first < n_records) {
            b0 = offsets[r_first]; e0 = offsets[r_first + 1];
            v0 = load_tile_block(bases + b0, (u32)min((u64)span, e0 - b0), buf_end, lane);
        }
        if(r_first + nwarps < n_records) { b1 = offsets[r_first + nwarps]; e1 = offsets[r_first + nwarps + 1]; }
        for(u64 r = r_first; r < n_records; r += nwarps) {
            u64 b2 = 0, e2 = 0;
            if(r + 2 * nwarps < n_records) { b2 = offsets[r + 2 * nwarps]; e2 = offsets[r + 2 * nwarps + 1]; }
            uint4 v1 = make_uint4(0, 0, 0, 0);
            if(r + nwarps < n_records) v1 = load_tile_block(bases + b1, (u32)min((u64)span, e1 - b1), buf_end, lane);
            sink.begin(TAXA ? taxa_out + taxa_offsets[r] : nullptr, TAXA ? taxa_offsets[r + 1] - taxa_offsets[r] : 0ull);
            encode_sequence<FAM>(P, S, bases + b0, e0 - b0, buf_end, sink, lane, &v0);
            const u32 taxon = sink.resolve(S, X, lane);
            if(lane == 0) {
                if(mate1_out) mate1_out[r] = sink.n_hit + sink.n_miss;
                taxon_out[r] = taxon;
                if(nhit_out) nhit_out[r] = sink.n_hit;
                if(nmiss_out) nmiss_out[r] = sink.n_miss;
                if(sink.overflow) { if(ovf_idx) ovf_idx[atomicAdd(ovf_cnt, 1ull)] = (u32)r; else atomicOr(status, 2u); }
                if(TAXA && sink.taxa_short) atomicOr(status, 1u);
            }
            if(sink.overflow && ovf_idx) {} else if(taxon) ++n_cls; else ++n_uncls;
            __syncwarp();
            b0 = b1; e0 = e1; v0 = v1; b1 = b2; e1 = e2;
        }
    } else {
    const u64 n_loop = rec_list ? (u64)*rec_count : n_records;
    for(u64 it = r_first; it < n_loop; it += nwarps) {
        const u64 r = rec_list ? (u64)rec_list[it] : it;
        sink.begin(TAXA ? taxa_out + taxa_offsets[r] : nullptr, TAXA ? taxa_offsets[r + 1] - taxa_offsets[r] : 0ull);
        for(u32 mt = 0; mt < mates; ++mt) {
            const u64 b = offsets[r * mates + mt], e = offsets[r * mates + mt + 1];
            encode_sequence<FAM>(P, S, bases + b, e - b, buf_end, sink, lane);
            // k-mers the first mate produced (classify_seq's first ambig_count term, classifier.h:232)
            if(mt == 0 && mate1_out && lane == 0) mate1_out[r] = sink.n_hit + sink.n_miss;
        }
        const u32 taxon = sink.resolve(S, X, lane);
        if(lane == 0) {
            taxon_out[r] = taxon;
            if(nhit_out) nhit_out[r] = sink.n_hit;
            if(nmiss_out) nmiss_out[r] = sink.n_miss;
            if(sink.overflow) { if(ovf_idx) ovf_idx[atomicAdd(ovf_cnt, 1ull)] = (u32)r; else atomicOr(status, 2u); }
            if(TAXA && sink.taxa_short) atomicOr(status, 1u);
        }
        if(sink.overflow && ovf_idx) {} else if(taxon) ++n_cls; else ++n_uncls;
        __syncwarp();
    }
    }
    if(lane == 0 && (n_cls | n_uncls)) {
        atomicAdd(&counters[0], (unsigned long long)n_cls);
        atomicAdd(&counters[1], (unsigned long long)n_uncls);
    }
}

// Bucketised open addressing, 4 x u64 slots per 32-byte bucket:
//   slot = [ low (64-b) bits of mix64(key) | disp:4 | overflow flags:F | value id:(b-4-F) ],  empty = ~0.
// A key lives in its home bucket (disp 0) or, if that was full, in the first later bucket with room (disp <= 14);
// the home bucket's slot 0 then gets the key's overflow flag CLEARED (flag_count_for, bns_device.cuh) so that misses
// stop after one sector otherwise.
__global__ void bns_insert_kernel(u64 *__restrict__ slots, TableFmt fmt, const u64 *__restrict__ keys,
                                  const u32 *__restrict__ vals, u64 n, const u32 *__restrict__ values, u32 n_values,
                                  unsigned long long *__restrict__ stats /* [0] failed, [1] displaced, [2] bad value */,
                                  u64 *__restrict__ fail_keys, u32 *__restrict__ fail_vals, u64 fail_cap) {
    // fail_keys != nullptr: the (key, value) pairs that found no room are listed there (the first fail_cap of them): the
    // stash of a LAYOUT_MINIMIZER table is built from that list
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const u64 key = keys[i];
    const u32 vid = value_id(values, n_values, vals[i]);
    if(vid == VAL_MISS) { atomicAdd(&stats[2], 1ull); return; }
    bool possible;
    const TableHash th = table_hash(fmt, key, possible);
    if(!possible) { atomicAdd(&stats[0], 1ull); return; }
    const u32 F = fmt.F;
    const u64 home = th.home;
    const u32 tag_shift = fmt.tag_shift(), flag_shift = fmt.flag_shift();
    const u64 tag = th.tag;
    for(u32 d = 0; d <= fmt.max_disp(); ++d) {
        u64 *bk = slots + (probe_bucket(fmt.layout, home, d, fmt.b) << 2);
        const u64 entry = tag | ((u64)d << tag_shift) | (((1ull << F) - 1) << flag_shift) | vid;
        const int ns = (int)fmt.unit_slots();
        for(int s = 0; s < ns; ++s) {
            u64 cur = bk[s];
            if(cur == ~0ull) {
                cur = atomicCAS((unsigned long long *)&bk[s], ~0ull, (unsigned long long)entry);
                if(cur == ~0ull) { if(d) atomicAdd(&stats[1], 1ull); return; }
            }
            if(((cur ^ entry) >> tag_shift) == 0) return;          // same key already present: first value stays
            // invariant for match4_home: upper words are unique within a bucket. A clash (2^-32 per pair) fails the
            // build; the host rebuilds with one more bucket bit, which re-draws every upper word.
            if(fmt.layout == LAYOUT_HASH && (u32)(cur >> 32) == (u32)(entry >> 32)) { atomicAdd(&stats[0], 1ull); return; }
        }
        if(d == 0) atomicAnd((unsigned long long *)&bk[0], ~(1ull << (flag_shift + (th.fsel & (F - 1)))));
    }
    const unsigned long long at = atomicAdd(&stats[0], 1ull);
    if(fail_keys && at < fail_cap) { fail_keys[at] = key; fail_vals[at] = vals[i]; }
}

__global__ void bns_table_stats_kernel(const u64 *__restrict__ slots, u64 n_buckets, TableFmt fmt,
                                       unsigned long long *__restrict__ out /* [0] entries [1] overflowed probes [2] max disp */) {
    const u64 ns = fmt.unit_slots();
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;        // one thread per probe unit
    if(i >= n_buckets * 4 / ns) return;
    const u32 tag_shift = fmt.tag_shift(), F = fmt.F;
    u32 cnt = 0, md = 0;
    bool flagged = false;
    const u64 flags = ((1ull << F) - 1) << (tag_shift - F);
    for(u64 s = 0; s < ns; ++s) {
        const u64 v = slots[ns * i + s];
        if(v != ~0ull) { ++cnt; md = max(md, (u32)((v >> tag_shift) & ((1u << fmt.disp_bits) - 1))); }
    }
    flagged = cnt == ns && (slots[ns * i] & flags) != flags;
    if(cnt) atomicAdd(&out[0], (unsigned long long)cnt);
    if(flagged) atomicAdd(&out[1], 1ull);
    if(md) atomicMax(&out[2], (unsigned long long)md);
}

__global__ void bns_lookup_kernel(TableView T, const u32 *__restrict__ dict, const u64 *__restrict__ keys, u64 n,
                                  u32 *__restrict__ vals_out, uint8_t *__restrict__ found_out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    u32 v;
    if(T.fmt.layout == LAYOUT_HASH) {                              // the home-bucket test the classify kernels use
        bool possible;
        const TableHash h = table_hash(T.fmt, keys[i], possible);
        u64 a, b, c, d;
        ld_bucket(T.slots + (h.home << 2), a, b, c, d);
        v = match4_home(T, h.tag, a, b, c, d);
        if(v == VAL_MISS && !(((u32)a >> (T.flag_shift + (h.fsel & T.flag_mask))) & 1u)) v = probe_displaced(T, h.home, h.tag);
    } else v = probe_key(T, keys[i]);
    found_out[i] = v != VAL_MISS;
    vals_out[i] = v != VAL_MISS ? dict[v] : 0u;
}

// 32-byte sectors the probes of `keys` touch (a LAYOUT_MINIMIZER probe is two adjacent ones)
__global__ void bns_sectors_kernel(TableView T, const u64 *__restrict__ keys, u64 n, unsigned long long *__restrict__ total) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 touched = 0;
    if(i < n) {
        bool possible;
        const TableHash h = table_hash(T.fmt, keys[i], possible);
        const u32 per = T.fmt.unit_slots() / 4;
        for(u32 d = 0; d <= T.fmt.max_disp(); ++d) {
            u32 fw; bool last_free;
            const u32 v = match_probe(T, probe_bucket(T.fmt.layout, h.home, d, T.fmt.b), h.tag | ((u64)d << T.tag_shift), fw, last_free);
            touched += per;
            if(v != VAL_MISS) break;
            if(d == 0 ? (((fw >> (T.flag_shift + (h.fsel & T.flag_mask))) & 1u) != 0) : last_free) break;
        }
    }
    touched = __reduce_add_sync(FULL, touched);
    if((threadIdx.x & 31) == 0 && touched) atomicAdd(total, (unsigned long long)touched);
}

// resolve_tree over explicit lists; one warp per list. Values must be DB values (they are looked up in `values`).
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
bns_resolve_kernel(TaxView X, const u32 *__restrict__ values, u32 n_values, const u32 *__restrict__ taxa,
                   const uint16_t *__restrict__ counts, const u64 *__restrict__ offsets, u64 n_lists,
                   u32 *__restrict__ taxon_out, u32 *__restrict__ status) {
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const WarpSmem S = carve(g_smem + wid * warp_smem_bytes(0, true), 0);
    const u64 nwarps = (u64)gridDim.x * WARPS_PER_CTA;
    ClassifySink<false> sink;
    sink.vi = X.val_info;
    for(u64 r = (u64)blockIdx.x * WARPS_PER_CTA + wid; r < n_lists; r += nwarps) {
        sink.begin(nullptr);
        const u64 b = offsets[r], e = offsets[r + 1];
        // linear::counter::add semantics: repeated keys accumulate (u16 wrap applied at resolve time)
        for(u64 i = b; i < e; ++i) {
            const u32 id = value_id(values, n_values, taxa[i]);
            if(id == VAL_MISS) { if(lane == 0) atomicOr(status, 4u); continue; }
            sink.add(S, id, counts[i], lane);
        }
        if(sink.overflow && lane == 0) atomicOr(status, 2u);
        const u32 t = sink.resolve(S, X, lane);
        if(lane == 0) taxon_out[r] = t;
        __syncwarp();
    }
}

// bonsai build on the device: one warp per genome record (contig), every emitted k-mer goes to BuildSink::insert
template <int FAM>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
bns_build_kernel(const __grid_constant__ EncParams P, const char *__restrict__ bases, const u64 *__restrict__ offsets,
                 u64 n_seqs, u64 total_bases, BuildSink proto, u32 ring_cap) {
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const WarpSmem S = carve(g_smem + wid * warp_smem_bytes(ring_cap, false), ring_cap);
    const u64 nwarps = (u64)gridDim.x * WARPS_PER_CTA;
    BuildSink sink = proto;
    sink.n_new = sink.n_fail = sink.n_bad = 0;
    for(u64 r = (u64)blockIdx.x * WARPS_PER_CTA + wid; r < n_seqs; r += nwarps) {
        const u64 b = offsets[2 * r], e = offsets[2 * r + 1];           // (start, end) pairs: pieces may overlap
        encode_sequence<FAM>(P, S, bases + b, e - b, bases + total_bases, sink, lane);
    }
    const u32 nn = __reduce_add_sync(FULL, sink.n_new), nf = __reduce_add_sync(FULL, sink.n_fail), nbad = __reduce_add_sync(FULL, sink.n_bad);
    if(lane == 0) {
        if(nn) atomicAdd(&sink.stats[3], (unsigned long long)nn);
        if(nf) atomicAdd(&sink.stats[0], (unsigned long long)nf);
        if(nbad) atomicAdd(&sink.stats[2], (unsigned long long)nbad);
    }
}

// table -> (key, value) pairs: the home bucket is bucket - disp, the key is unmix64(home : remainder)
__global__ void bns_dump_kernel(const u64 *__restrict__ slots, u64 n_buckets, TableFmt fmt, const u32 *__restrict__ dict,
                                u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, u64 cap,
                                unsigned long long *__restrict__ counter) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_buckets * 4) return;
    const u64 v = slots[i];
    if(v == ~0ull) return;
    const u32 tag_shift = fmt.tag_shift(), b = fmt.b;
    const u64 bucket = i >> 2, disp = (v >> tag_shift) & ((1u << fmt.disp_bits) - 1);
    const u64 home = probe_home(fmt.layout, bucket, (u32)disp, fmt.b);
    const u64 rem_mask = ~0ull << fmt.fmt_bits;                    // the remainder, left-aligned
    const u64 key = fmt.layout == LAYOUT_MINIMIZER ? loc_decode(home, v & rem_mask, fmt.kt, b)
                                                   : unmix64((home << (64 - b)) | (v >> b));
    const u64 at = atomicAdd(counter, 1ull);
    if(at < cap) { keys_out[at] = key; vals_out[at] = dict[(u32)v & ((1u << fmt.flag_shift()) - 1)]; }
}

// Run-length encoding of the ordered per-k-mer hit list of each record (what append_taxa_runs prints, classifier.h:46-61):
// one thread per record counts its runs, reserves that many entries of the chunk's run buffer with one atomic per warp
// and writes (taxid << 32 | run length) words. The verbose output then costs 8 bytes per RUN over PCIe instead of 4 bytes
// per k-mer window slot.
__global__ void bns_rle_kernel(const u32 *__restrict__ taxa, const u64 *__restrict__ taxa_offsets, const u32 *__restrict__ nhit,
                               u64 n_records, u64 *__restrict__ runs, unsigned long long *__restrict__ total,
                               u64 *__restrict__ run_pos, u32 *__restrict__ n_runs) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 lane = threadIdx.x & 31u;
    u32 n = 0, nr = 0;
    const u32 *t = nullptr;
    if(r < n_records) {
        n = nhit[r];
        t = taxa + taxa_offsets[r];
        for(u32 i = 0; i < n; ++i) nr += (i == 0 || t[i] != t[i - 1]);
    }
    // one reservation per warp
    u32 incl = nr;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, d); if(lane >= (u32)d) incl += y; }
    const u32 wtot = __shfl_sync(FULL, incl, 31);
    unsigned long long base = 0;
    if(lane == 31 && wtot) base = atomicAdd(total, (unsigned long long)wtot);
    base = __shfl_sync(FULL, base, 31);
    if(r >= n_records) return;
    u64 at = base + (incl - nr);
    run_pos[r] = at;
    n_runs[r] = nr;
    u32 last = 0, run = 0;
    for(u32 i = 0; i < n; ++i) {
        const u32 v = t[i];
        if(i && v != last) { runs[at++] = ((u64)last << 32) | run; run = 0; }
        last = v; ++run;
    }
    if(n) runs[at] = ((u64)last << 32) | run;
}

// independent 32-byte loads at uniformly random buckets: the random-access ceiling the lookup is measured against
__global__ void bns_gather_kernel(const u64 *__restrict__ slots, u32 b, u64 n_loads, u64 seed,
                                  unsigned long long *__restrict__ sink_out) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (u64)gridDim.x * blockDim.x;
    u64 acc = 0;
    for(u64 i = tid * 4; i < n_loads; i += nthreads * 4) {
        u64 s[4][4];
#pragma unroll
        for(int j = 0; j < 4; ++j) {
            const u64 h = mix64(seed + i + j);
            ld_bucket(slots + ((h >> (64 - b)) << 2), s[j][0], s[j][1], s[j][2], s[j][3]);
        }
#pragma unroll
        for(int j = 0; j < 4; ++j) acc ^= s[j][0] ^ s[j][1] ^ s[j][2] ^ s[j][3];
    }
    if(acc == 0x1234567u) *sink_out = acc;       // keep the loads alive
}

}  // namespace bns
#include "bns_classify_u.cuh"
namespace bns {

// ---------------------------------------------------------------------------------------------
// host-side launchers (called from bns_api.cu)
// ---------------------------------------------------------------------------------------------
size_t stream_smem_bytes(u32 ring_cap, bool classify) { return WARPS_PER_CTA * warp_smem_bytes(ring_cap, classify); }

typedef void (*encode_fn)(const EncParams, const char *, const u64 *, u64, u64, u64 *, const u64 *, u32 *, u32, u32 *);
typedef void (*classify_fn)(const EncParams, const char *, const u64 *, u64, u32, u64, TableView, TaxView, u32 *, u32 *, u32 *,
                            u32 *, const u64 *, u32 *, u32, unsigned long long *, u32 *, const u32 *, const unsigned long long *,
                            u32 *, unsigned long long *, u32 *, u32);

static encode_fn pick_encode(u32 fam) {
    switch(fam) {
        case FAM_U: return bns_encode_kernel<FAM_U>;
        case FAM_K: return bns_encode_kernel<FAM_K>;
        case FAM_R: return bns_encode_kernel<FAM_R>;
        default: return bns_encode_kernel<FAM_NONE>;
    }
}
static classify_fn pick_classify(u32 fam, bool taxa) {
    switch(fam) {
        case FAM_U: return taxa ? bns_classify_kernel<FAM_U, true> : bns_classify_kernel<FAM_U, false>;
        case FAM_K: return taxa ? bns_classify_kernel<FAM_K, true> : bns_classify_kernel<FAM_K, false>;
        case FAM_R: return taxa ? bns_classify_kernel<FAM_R, true> : bns_classify_kernel<FAM_R, false>;
        default: return taxa ? bns_classify_kernel<FAM_NONE, true> : bns_classify_kernel<FAM_NONE, false>;
    }
}

cudaError_t launch_encode(const EncParams &P, int grid, size_t smem, cudaStream_t st, const char *bases, const u64 *offsets,
                          u64 n_seqs, u64 total_bases, u64 *kmers_out, const u64 *out_offsets, u32 *counts_out, u32 ring_cap,
                          u32 *status) {
    encode_fn f = pick_encode(P.family);
    cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    f<<<grid, WARPS_PER_CTA * 32, smem, st>>>(P, bases, offsets, n_seqs, total_bases, kmers_out, out_offsets, counts_out,
                                             ring_cap, status);
    return cudaGetLastError();
}
// Which kernel a classify call runs and with what geometry. The lean kernel covers single-end records without an ordered
// hit list for every unspaced encoder: what `bonsai classify` runs (FAM_U) and the windowed minimizer modes (FAM_K
// canonical, FAM_R); everything else (spaced seeds, paired records, hit lists) goes to the generic stream kernel.
static int lean_mode(const EncParams &P, u32 mates, bool taxa, bool mate1) {
#ifdef BNS_NO_LEAN
    return -1;
#endif
    (void)mate1;
    if(mates > 2 || taxa) return -1;                                  // pairs share the batch machinery: two sequences per record
    if(P.family == FAM_U) return LEAN_U;
    const bool unspaced = P.c == P.k;
    if(!unspaced) {                                                   // spaced seed, window of one, comb inside the lane's 48-base window
        if(P.family != FAM_K || P.canon_elem || P.W != 1 || P.c > 45) return -1;
        for(u32 s = 0; s < P.n_seg; ++s) if(P.seg_len[s] > 29) return -1;
        return LEAN_S;
    }
    if(P.W > (u32)TILE) return -1;
    if(P.family == FAM_K && P.canon_elem) return LEAN_K;
    if(P.family == FAM_R) return LEAN_R;
    return -1;
}
typedef void (*classify_u_fn)(const EncParams, const char *, const u64 *, u64, TableView, TaxView, u32 *, u32 *, u32 *,
                              unsigned long long *, u32 *, u32 *, unsigned long long *, u32, u64, u32, u32 *,
                              u64 *, u64, unsigned long long *, u64 *, u32 *, const PackedIn);
static size_t lean_smem(bool runs, int warps) {
    return (size_t)warps * (4 * AGG_CAP * sizeof(u32) + LEAN_STAGE_BYTES + (runs ? RUNBUF * sizeof(u64) : 0));
}
template <int MODE, bool CANON, bool COUNTS, int KEY, bool RUNS = false>
static classify_u_fn pick_lean_k(u32 k, bool loc, bool pk = false) {
    if(pk) {                                                          // host-packed bases: LEAN_U / LEAN_S without run lists (plan_classify)
        constexpr bool OK = (MODE == LEAN_U || MODE == LEAN_S) && !RUNS;
        if(!OK) return nullptr;
        constexpr int M = OK ? MODE : LEAN_U;                          // keeps the other modes from instantiating packed variants
        constexpr bool C2 = OK ? CANON : true, N2 = OK ? COUNTS : true;
        if(loc) return k == 31 ? bns_classify_u_kernel<M, C2, 31, N2, 0, true, false, true> : bns_classify_u_kernel<M, C2, 0, N2, 0, true, false, true>;
        return k == 31 ? bns_classify_u_kernel<M, C2, 31, N2, 0, false, false, true> : bns_classify_u_kernel<M, C2, 0, N2, 0, false, false, true>;
    }
    if(loc) return k == 31 ? bns_classify_u_kernel<MODE, CANON, 31, COUNTS, KEY, true, RUNS> : bns_classify_u_kernel<MODE, CANON, 0, COUNTS, KEY, true, RUNS>;
    return k == 31 ? bns_classify_u_kernel<MODE, CANON, 31, COUNTS, KEY, false, RUNS> : bns_classify_u_kernel<MODE, CANON, 0, COUNTS, KEY, false, RUNS>;
}
template <int MODE, bool CANON>
static classify_u_fn pick_lean_key(u32 k, int key, bool loc) {
    if(key == LEAN_KEY_LEX) return pick_lean_k<MODE, CANON, true, LEAN_KEY_LEX>(k, loc);
    if(key == LEAN_KEY_ELEM) return pick_lean_k<MODE, CANON, true, LEAN_KEY_ELEM>(k, loc);
    return pick_lean_k<MODE, CANON, true, LEAN_KEY_PAIR>(k, loc);
}
// How the (score, k-mer) pairs of a window can be ordered by one 64-bit word (bit-identical minima):
//  * Lex: the score is a bijection of the k-mer (ties in score are ties in k-mer).
//  * entropy scores under the saturating cast: every k-mer but the all-A one scores ~0 once kmer / 0.001 >= 2^64 for the
//    other homopolymers, i.e. (4^k-1)/3 * 1000 >= 2^64, k >= 28 (rolling entropy), and for every k when the entropy is
//    the NOT_FULL constant (kmer / -0.9999 <= -1 for kmer >= 1); the all-A k-mer is 0 and scores 0: the pair order is
//    the k-mer order.
static int lean_key(const EncParams &P) {
    if(P.score_kind == SC_LEX) return LEAN_KEY_LEX;
    if(!P.cast_wrap && (P.score_kind == SC_ENT_NOTFULL || (P.score_kind == SC_ENT_ROLL && P.k >= 28))) return LEAN_KEY_ELEM;
    return LEAN_KEY_PAIR;
}
static classify_u_fn pick_lean(const EncParams &P, int mode, bool counts, bool loc, bool runs = false, bool pk = false, bool sv = false) {
    if(sv) {                                                          // small value dictionary: counts in lane registers (plan_classify)
        if(runs) return loc ? bns_classify_u_kernel<LEAN_U, true, 31, true, 0, true, true, false, true> : bns_classify_u_kernel<LEAN_U, true, 31, true, 0, false, true, false, true>;
        if(pk) {
            if(loc) return counts ? bns_classify_u_kernel<LEAN_U, true, 31, true, 0, true, false, true, true> : bns_classify_u_kernel<LEAN_U, true, 31, false, 0, true, false, true, true>;
            return counts ? bns_classify_u_kernel<LEAN_U, true, 31, true, 0, false, false, true, true> : bns_classify_u_kernel<LEAN_U, true, 31, false, 0, false, false, true, true>;
        }
        if(loc) return counts ? bns_classify_u_kernel<LEAN_U, true, 31, true, 0, true, false, false, true> : bns_classify_u_kernel<LEAN_U, true, 31, false, 0, true, false, false, true>;
        return counts ? bns_classify_u_kernel<LEAN_U, true, 31, true, 0, false, false, false, true> : bns_classify_u_kernel<LEAN_U, true, 31, false, 0, false, false, false, true>;
    }
    if(pk) {
        if(mode == LEAN_S) return pick_lean_k<LEAN_S, false, true, 0>(P.k, loc, true);
        if(P.canon_elem) return counts ? pick_lean_k<LEAN_U, true, true, 0>(P.k, loc, true) : pick_lean_k<LEAN_U, true, false, 0>(P.k, loc, true);
        return counts ? pick_lean_k<LEAN_U, false, true, 0>(P.k, loc, true) : pick_lean_k<LEAN_U, false, false, 0>(P.k, loc, true);
    }
    if(mode == LEAN_S) return pick_lean_k<LEAN_S, false, true, 0>(P.k, loc);
    if(mode == LEAN_K) return pick_lean_key<LEAN_K, true>(P.k, lean_key(P), loc);
    if(mode == LEAN_R) return P.canon_emit ? pick_lean_key<LEAN_R, true>(P.k, lean_key(P), loc) : pick_lean_key<LEAN_R, false>(P.k, lean_key(P), loc);
    if(runs) return P.canon_elem ? pick_lean_k<LEAN_U, true, true, 0, true>(P.k, loc) : pick_lean_k<LEAN_U, false, true, 0, true>(P.k, loc);
    if(P.canon_elem) return counts ? pick_lean_k<LEAN_U, true, true, 0>(P.k, loc) : pick_lean_k<LEAN_U, true, false, 0>(P.k, loc);
    return counts ? pick_lean_k<LEAN_U, false, true, 0>(P.k, loc) : pick_lean_k<LEAN_U, false, false, 0>(P.k, loc);
}

ClassifyPlan plan_classify(const EncParams &P, const TableView &T, u32 ring_cap, int n_sm, u64 n_records, u32 mates, bool taxa, bool mate1, bool counts, bool runs,
                           bool packed) {
    ClassifyPlan pl;
    // run lists come out of the lean kernel for what `bonsai classify` runs (every k-mer, no window); the other encoders keep the
    // ordered hit list of the generic kernel, run-length encoded by bns_rle_kernel
    pl.lean_mode = lean_mode(P, mates, taxa && !runs, mate1);
    if(runs && pl.lean_mode != LEAN_U) pl.lean_mode = lean_mode(P, mates, true, mate1);
    // the lean kernel spells LAYOUT_MINIMIZER keys as k-mers of ITS k: a table of another k goes through the generic probe
    pl.loc = T.fmt.layout == LAYOUT_MINIMIZER;
    if(pl.loc && T.fmt.kt != P.k) pl.lean_mode = -1;
    pl.lean = pl.lean_mode >= 0;
    pl.runs = runs && pl.lean_mode == LEAN_U;
    pl.counts = counts || mates == 2 || mate1 || runs;                // the pair bookkeeping lives in the COUNTS variants
    // host-packed bases (bns_pack.h) are read by the variants of what `bonsai classify` runs; a database of more than AGG_CAP
    // values may send records to the generic kernel's second pass, which reads ASCII
    pl.packed = packed && (pl.lean_mode == LEAN_U || pl.lean_mode == LEAN_S) && !pl.runs && T.n_values <= (u32)AGG_CAP;
    // value dictionaries of at most 32 entries (one lane per value): what `bonsai classify` runs at k = 31 keeps its counts in registers
    static const bool no_sv = [] { const char *e = getenv("BNS_B200_NO_SV"); return e && e[0] == '1'; }();
    pl.sv = !no_sv && pl.lean_mode == LEAN_U && P.canon_elem && P.k == 31 && T.n_values >= 1 && T.n_values <= 32;
    int nb = 0;
    if(pl.lean) {
        classify_u_fn f = pick_lean(P, pl.lean_mode, pl.counts, pl.loc, pl.runs, pl.packed, pl.sv);
        // one CTA of LEAN_WARPS warps per SM when the batches fill every SM that way, else CTAs of LEAN_WARPS_SMALL (bns_classify_u.cuh)
        const u64 n_batches = (n_records * mates + RB - 1) / RB;
        pl.lean_warps = n_batches >= (u64)n_sm * LEAN_WARPS ? LEAN_WARPS : LEAN_WARPS_SMALL;
        pl.smem = lean_smem(pl.runs, pl.lean_warps);
        cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_smem(pl.runs, LEAN_WARPS));
        cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f, pl.lean_warps * 32, pl.smem);
        const u64 want = (n_batches + pl.lean_warps - 1) / pl.lean_warps;                     // one batch per warp at least
        pl.grid = (int)std::max<u64>(1, std::min<u64>(want, (u64)n_sm * (nb > 0 ? nb : 1)));
    }
    // A second pass of the generic kernel takes what the first leaves: windowed records of more than one tile and 32-T restarts
    // (lean windowed modes), and -- only possible when the database holds more than AGG_CAP distinct values -- records that hit
    // more distinct taxa than the shared-memory lists hold; it keeps its lists in global memory.
    pl.big_cap = T.n_values > (u32)AGG_CAP ? std::min<u32>((T.n_values + 31u) & ~31u, 1u << 16) : 0u;
    pl.second_pass = (pl.lean && (pl.lean_mode == LEAN_K || pl.lean_mode == LEAN_R)) || pl.big_cap != 0;
    if(!pl.lean || pl.second_pass) {                                   // the generic kernel: everything, or the deferred records
        int ng = 0;
        classify_fn f = pick_classify(P.family, taxa);
        pl.gen_smem = stream_smem_bytes(ring_cap, true);
        cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.gen_smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ng, f, WARPS_PER_CTA * 32, pl.gen_smem);
        const u64 want = (n_records + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        pl.gen_grid = (int)std::max<u64>(1, std::min<u64>(want, (u64)n_sm * (ng > 0 ? ng : 1)));
        if(!pl.lean) { nb = ng; pl.grid = pl.gen_grid; pl.smem = pl.gen_smem; }
        pl.pass2_grid = pl.big_cap ? std::min(pl.gen_grid, 32) : pl.gen_grid;
    }
    pl.occupancy = nb;
    return pl;
}

// defer_idx / defer_cnt: device scratch of n_records u32 and one zeroed counter (windowed lean modes only)
cudaError_t launch_classify(const EncParams &P, const ClassifyPlan &pl, cudaStream_t st, const char *bases, const u64 *offsets,
                            u64 n_records, u32 mates, u64 total_bases, const TableView &T, const TaxView &X,
                            u32 *taxon_out, u32 *nhit_out, u32 *nmiss_out, u32 *taxa_out, const u64 *taxa_offsets,
                            u32 *mate1_out, u32 ring_cap, unsigned long long *counters, u32 *status,
                            u32 *defer_idx, unsigned long long *defer_cnt, int *n_launched, const RunsOut *ro, u32 *big_scratch, const PackedIn *pk) {
    if(n_launched) *n_launched = 1;
    if(pl.lean) {
        classify_u_fn f = pick_lean(P, pl.lean_mode, pl.counts, pl.loc, pl.runs, pl.packed, pl.sv);
        if(pl.packed && !pk) return cudaErrorInvalidValue;
        const PackedIn pki = pl.packed ? *pk : PackedIn{nullptr, nullptr, 0u, 0ull};
        f<<<pl.grid, pl.lean_warps * 32, pl.smem, st>>>(P, bases, offsets, n_records * mates, T, X, taxon_out, nhit_out, nmiss_out,
                                                    counters, status, defer_idx, defer_cnt, pl.fixed_len, pl.fixed_base, mates, mate1_out,
                                                    pl.runs ? ro->runs : nullptr, pl.runs ? ro->cap : 0, pl.runs ? ro->total : nullptr,
                                                    pl.runs ? ro->run_pos : nullptr, pl.runs ? ro->n_runs : nullptr, pki);
        cudaError_t e = cudaGetLastError();
        if(e != cudaSuccess || !pl.second_pass) return e;
    } else {
        classify_fn f = pick_classify(P.family, taxa_out != nullptr);
        f<<<pl.grid, WARPS_PER_CTA * 32, pl.smem, st>>>(P, bases, offsets, n_records, mates, total_bases, T, X, taxon_out, nhit_out,
                                                        nmiss_out, taxa_out, taxa_offsets, mate1_out, ring_cap, counters, status,
                                                        nullptr, nullptr, pl.second_pass ? defer_idx : nullptr, defer_cnt, nullptr, 0u);
        cudaError_t e = cudaGetLastError();
        if(e != cudaSuccess || !pl.second_pass) return e;
    }
    // the records the first pass left: usually none, the kernel reads the count on the device and returns at once
    classify_fn g = pick_classify(P.family, !pl.lean && taxa_out != nullptr);
    g<<<pl.pass2_grid, WARPS_PER_CTA * 32, pl.gen_smem, st>>>(P, bases, offsets, n_records, mates, total_bases, T, X, taxon_out,
                                                             nhit_out, nmiss_out, pl.lean ? nullptr : taxa_out, pl.lean ? nullptr : taxa_offsets,
                                                             mate1_out, ring_cap, counters, status, defer_idx, defer_cnt, nullptr, nullptr,
                                                             pl.big_cap ? big_scratch : nullptr, pl.big_cap);
    if(n_launched) *n_launched = 2;
    return cudaGetLastError();
}
u64 lean_wave_reads(int n_sm) { return (u64)n_sm * LEAN_WARPS * RB; }
size_t pass2_scratch_words(const ClassifyPlan &pl) { return pl.big_cap ? (size_t)pl.pass2_grid * WARPS_PER_CTA * 4 * pl.big_cap : 0; }
u64 runs_slack(const ClassifyPlan &pl) { return pl.runs ? (u64)pl.grid * pl.lean_warps * RUN_BLOCK : 0; }
int encode_occupancy(const EncParams &P, size_t smem) {
    int nb = 0;
    encode_fn f = pick_encode(P.family);
    cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f, WARPS_PER_CTA * 32, smem);
    return nb;
}
cudaError_t launch_build(const EncParams &P, int grid, size_t smem, cudaStream_t st, const char *bases, const u64 *offsets,
                         u64 n_seqs, u64 total_bases, u64 *slots, const TableFmt &fmt, u32 vid, const TaxView &X, const u32 *values,
                         u32 n_values, unsigned long long *stats, u32 ring_cap) {
    BuildSink sk;
    sk.fmt = fmt;
    sk.slots = slots; sk.b = fmt.b; sk.tag_shift = fmt.tag_shift(); sk.flag_shift = fmt.flag_shift(); sk.flag_mask = fmt.F - 1;
    sk.val_mask = (1u << sk.flag_shift) - 1; sk.vid = vid;
    sk.val_info = X.val_info; sk.node_info = X.node_info; sk.values = values; sk.n_values = n_values;
    sk.node_of_one = X.node_of_one; sk.stats = stats; sk.n_new = sk.n_fail = sk.n_bad = 0;
    void (*f)(const EncParams, const char *, const u64 *, u64, u64, BuildSink, u32) =
        P.family == FAM_U ? bns_build_kernel<FAM_U> : P.family == FAM_K ? bns_build_kernel<FAM_K>
        : P.family == FAM_R ? bns_build_kernel<FAM_R> : bns_build_kernel<FAM_NONE>;
    cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    f<<<grid, WARPS_PER_CTA * 32, smem, st>>>(P, bases, offsets, n_seqs, total_bases, sk, ring_cap);
    return cudaGetLastError();
}
cudaError_t launch_dump(cudaStream_t st, const u64 *slots, u64 n_buckets, const TableFmt &fmt, const u32 *dict, u64 *keys_out, u32 *vals_out,
                        u64 cap, unsigned long long *counter) {
    const u64 n = n_buckets * 4;
    bns_dump_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(slots, n_buckets, fmt, dict, keys_out, vals_out, cap, counter);
    return cudaGetLastError();
}
cudaError_t launch_insert(cudaStream_t st, u64 *slots, const TableFmt &fmt, const u64 *keys, const u32 *vals, u64 n, const u32 *values,
                          u32 n_values, unsigned long long *stats, u64 *fail_keys, u32 *fail_vals, u64 fail_cap) {
    if(!n) return cudaSuccess;
    bns_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(slots, fmt, keys, vals, n, values, n_values, stats, fail_keys, fail_vals, fail_cap);
    return cudaGetLastError();
}
cudaError_t launch_table_stats(cudaStream_t st, const u64 *slots, u64 n_buckets, const TableFmt &fmt, unsigned long long *out) {
    const u64 n_units = n_buckets * 4 / fmt.unit_slots();
    bns_table_stats_kernel<<<(unsigned)((n_units + 255) / 256), 256, 0, st>>>(slots, n_buckets, fmt, out);
    return cudaGetLastError();
}
cudaError_t launch_lookup(cudaStream_t st, const TableView &T, const u32 *dict, const u64 *keys, u64 n, u32 *vals_out,
                          uint8_t *found_out) {
    if(!n) return cudaSuccess;
    bns_lookup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(T, dict, keys, n, vals_out, found_out);
    return cudaGetLastError();
}
cudaError_t launch_sectors(cudaStream_t st, const TableView &T, const u64 *keys, u64 n, unsigned long long *total) {
    if(!n) return cudaSuccess;
    bns_sectors_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(T, keys, n, total);
    return cudaGetLastError();
}
cudaError_t launch_resolve(int grid, cudaStream_t st, const TaxView &X, const u32 *values, u32 n_values, const u32 *taxa,
                           const uint16_t *counts, const u64 *offsets, u64 n_lists, u32 *taxon_out, u32 *status) {
    const size_t smem = WARPS_PER_CTA * warp_smem_bytes(0, true);
    bns_resolve_kernel<<<grid, WARPS_PER_CTA * 32, smem, st>>>(X, values, n_values, taxa, counts, offsets, n_lists, taxon_out, status);
    return cudaGetLastError();
}
cudaError_t launch_rle(cudaStream_t st, const u32 *taxa, const u64 *taxa_offsets, const u32 *nhit, u64 n_records, u64 *runs,
                       unsigned long long *total, u64 *run_pos, u32 *n_runs) {
    if(!n_records) return cudaSuccess;
    bns_rle_kernel<<<(unsigned)((n_records + 255) / 256), 256, 0, st>>>(taxa, taxa_offsets, nhit, n_records, runs, total, run_pos, n_runs);
    return cudaGetLastError();
}
cudaError_t launch_gather(int grid, cudaStream_t st, const u64 *slots, u32 b, u64 n_loads, u64 seed, unsigned long long *sink) {
    bns_gather_kernel<<<grid, 256, 0, st>>>(slots, b, n_loads, seed, sink);
    return cudaGetLastError();
}

}  // namespace bns
