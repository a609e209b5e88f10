// bns_classify_u.cuh -- the lean kernel: what `bonsai classify` always runs (bin/bonsai.cpp:152: k == w, no spaced seed: every
// (canonical) k-mer of the record looked up, taxon by resolve_tree) plus the windowed-minimizer and spaced-seed encoders,
// for single-end records and mate pairs, without the ordered hit list.
// Included by bns_kernels.cu after the shared building blocks (ClassifySink, probe_displaced, the TMA staging helpers).
//
// Same algorithm and table as the generic stream kernel bns_classify_kernel; what differs is the bookkeeping around the
// k-mers, which was two thirds of the instructions there (ncu, profiles/ncu_r01d_classify.txt vs ncu_r01g_classify.txt):
//   * a warp takes a BATCH of 32 consecutive sequences: their offsets arrive in shared memory by cp.async one batch ahead
//     (or are generated, for fixed-length batches), the first 8-byte blocks of the next sequence one sequence ahead, and
//     the batch's results leave with one coalesced store per output array and one atomic pair;
//   * bases are packed 8 per lane; the 2-bit words go from the lanes that packed them to the lanes that need them by SHFL,
//     not through shared memory; the lane's four k-mers and their reverse complements come out of one 96-bit window;
//   * everything is 32-bit: the bucket arrives as eight u32 (LDG.256), the bucket address is one IMAD.WIDE, tags are
//     funnel shifts;
//   * the first distinct taxon of a record and its count live in (warp-uniform) registers; shared memory is touched
//     only by records with >= 2 distinct taxa (resolve_tree's real work) -- 78 % of the classified reads have one; with a
//     value dictionary of at most 32 entries (SV variants) the counts live one value id per lane and no list exists at all;
//   * displaced-key probes (home bucket overflowed at build time: the key's overflow flag is cleared there) run in one
//     warp-level loop behind a one-vote conservative test;
//   * window minima (LEAN_K / LEAN_R) are computed with shuffles only, on one 64-bit word per element where that is exact;
//   * LAYOUT_MINIMIZER tables: the minimizers of all four k-mers of a lane come from 19 mixed 16-mers (8 computed, 11 by
//     SHFL.DOWN) and 32-bit min chains.
#pragma once

namespace bns {

constexpr int RB = 32;                       // records per warp batch
constexpr int LEAN_STAGE_BYTES = 2 * (RB + 2) * 8 + 2 * 32 * 8;      // per warp: offsets ring + first-tile ring
constexpr int RUNBUF = 128;                  // run-list variants: runs of one record buffered per warp before they are placed
constexpr u32 RUN_BLOCK = 256;               // ... into stretches of the chunk's run buffer a warp reserves with one atomic each
// Warps per CTA of the lean kernel: compiled for up to LEAN_WARPS (one CTA of 24 warps per SM at 80 registers), launched with
// that many when the batch fills every SM with such CTAs and with LEAN_WARPS_SMALL (three CTAs per SM) when it does not.
// Same 24 warps per SM either way; measured on B200 (profiles/occupancy_r02.log), 10 M reads of configs[1] / the 2^28-key stress
// table: 6 warps x 4 CTAs 1 560 / 492 Mreads/s, 8 x 3 1 611 / 502, 12 x 2 1 649 / 509, 24 x 1 1 729 / 527; more warps at fewer
// registers lose (7 x 4 and 9 x 3 at 72 registers: 1 523 / 462 and 1 526 / 454).
#ifndef BNS_LEAN_WARPS
#define BNS_LEAN_WARPS 24
#endif
#ifndef BNS_CLASSIFY_U_MIN_CTAS
#define BNS_CLASSIFY_U_MIN_CTAS 1
#endif
constexpr int LEAN_WARPS = BNS_LEAN_WARPS;   // most warps a CTA of the lean kernel is launched with (blockDim.x / 32 is what it has)
constexpr int LEAN_WARPS_SMALL = 8;


__device__ __forceinline__ void ld_bucket8(const void *p, u32 (&w)[8]) {
    // one 32-byte sector per probe (LDG.E.256), as eight 32-bit words: {lo, hi} of slots 0..3
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

struct ProbeConst {                          // launch-uniform pieces of the slot format (bns_insert_kernel)
    const char *slots;
    u32 b, idx_shift;                        // bucket = h.hi >> idx_shift (LAYOUT_HASH)
    u32 tag_shift, fmt_bits, max_disp, layout;
    u32 hm;                                  // bits of the low word that belong to {remainder, disp}
    u32 flags_all;                           // the overflow flags of slot 0 (all set = no key homed here was displaced)
    u32 flag_shift, flag_mask;               // a key's flag: bit flag_shift + (hl & flag_mask)
    u32 val_mask;
};

// buckets after the home bucket for ONE key of a LAYOUT_HASH table (rare). (th, tl0): the left-aligned remainder. Returns the
// low word of the matching slot or ~tl0 ("no match").
__device__ __forceinline__ u32 probe_displaced32(const ProbeConst Pc, u32 home, u32 th, u32 tl0) {
    const u32 b = Pc.b;
    const u32 bmask = (b == 32) ? ~0u : ((1u << b) - 1);
    for(u32 d = 1; d <= Pc.max_disp; ++d) {
        u32 w[8];
        ld_bucket8(Pc.slots + ((u64)((home + d) & bmask) << 5), w);
        const u32 tl = tl0 | (d << Pc.tag_shift);
#pragma unroll
        for(int j = 0; j < 4; ++j)
            if(w[2 * j + 1] == th && ((w[2 * j] ^ tl) & Pc.hm) == 0) return w[2 * j];
        if((w[6] & w[7]) == ~0u) break;                               // a bucket with a free slot ends the run
    }
    return ~tl0;
}
// the overflow chain of ONE key of a LAYOUT_MINIMIZER table: 64-byte units from a scrambled image of the home unit on
__device__ __forceinline__ u32 probe_chain_loc(const ProbeConst Pc, u32 home, u32 th, u32 tl0, bool &exhausted) {
    const u32 ub = Pc.b - 1, umask = (1u << ub) - 1;
    const u32 u0 = nmix(home >> 1, ub);
    exhausted = false;
    for(u32 d = 1; d <= Pc.max_disp; ++d) {
        u32 w[16];
        const char *up = Pc.slots + ((u64)((u0 + (d - 1)) & umask) << 6);
        ld_bucket8(up, *reinterpret_cast<u32 (*)[8]>(&w[0]));
        ld_bucket8(up + 32, *reinterpret_cast<u32 (*)[8]>(&w[8]));
        const u32 tl = tl0 | (d << Pc.tag_shift);
#pragma unroll
        for(int j = 0; j < 8; ++j)
            if(w[2 * j + 1] == th && ((w[2 * j] ^ tl) & Pc.hm) == 0) return w[2 * j];
        if((w[14] & w[15]) == ~0u) return ~tl0;                       // a unit with a free slot ends the run
    }
    exhausted = true;                                                 // full to its end: the key may sit in the stash
    return ~tl0;
}

// 8 ASCII bases -> 16 bits of 2-bit codes (first base in the top bits); returns non-zero iff some byte is not ACGTacgt
__device__ __forceinline__ u32 pack8_fast(uint2 v, u32 &codes16) {
    u32 c0, c1;
    const u32 d = pack4_fast(v.x, c0) | pack4_fast(v.y, c1);
    codes16 = (c0 << 8) | c1;
    return d;
}
__device__ __forceinline__ void pack8(uint2 v, u32 &codes16, u32 &bad8) {
    u32 c0, c1, b0, b1, t0, t1;
    pack4(v.x, c0, b0, t0); pack4(v.y, c1, b1, t1);
    codes16 = (c0 << 8) | c1;
    bad8 = (b0 << 4) | b1;
}

// ---------------------------------------------------------------------------------------------
// Windowed minimizers without shared memory (single-tile records: at most TILE window elements).
// An element is the 128-bit key (score, k-mer) of ElScore::operator< (qmap.h:23); lane l holds elements 4l..4l+3.
// The minimum over the last W elements at element g = 4l+i is the minimum of
//   a suffix of lane l+f's four elements, the lane minima of lanes l+f+1 .. l-1, and a prefix of lane l's own elements,
// with f = floor((i-W+1)/4). Lane minima over runs of lanes come from a doubling table built with SHFL.UP; a run of c
// lanes is assembled from the set bits of c while the table is built, so no level is kept.
// ---------------------------------------------------------------------------------------------

// NW = 4: (score hi, score lo, element hi, element lo);  NW = 2: one 64-bit key (hi, lo) when the order of the pairs is the
// order of a single word: the Lex score is a bijection of the k-mer (the k-mer is recovered with lex_unscore), and with
// the saturating cast the entropy scores are non-decreasing in the k-mer (LEAN_KEY_ELEM, see pick_lean()).
template <int NW> struct WKey { u32 w[NW]; };
template <int NW> __device__ __forceinline__ WKey<NW> wk_inf() { WKey<NW> k; for(int i = 0; i < NW; ++i) k.w[i] = ~0u; return k; }
template <int NW> __device__ __forceinline__ WKey<NW> wk_min(const WKey<NW> &a, const WKey<NW> &b) {
    bool lt;
    const u64 a0 = ((u64)a.w[0] << 32) | a.w[1], b0 = ((u64)b.w[0] << 32) | b.w[1];
    if(NW == 2) lt = b0 < a0;
    else {
        const u64 a1 = ((u64)a.w[NW - 2] << 32) | a.w[NW - 1], b1 = ((u64)b.w[NW - 2] << 32) | b.w[NW - 1];
        lt = b0 < a0 || (b0 == a0 && b1 < a1);
    }
    WKey<NW> r;
#pragma unroll
    for(int i = 0; i < NW; ++i) r.w[i] = lt ? b.w[i] : a.w[i];
    return r;
}
template <int NW> __device__ __forceinline__ WKey<NW> wk_shfl_up(const WKey<NW> &a, u32 d) {
    WKey<NW> r;
#pragma unroll
    for(int i = 0; i < NW; ++i) r.w[i] = __shfl_up_sync(FULL, a.w[i], d);
    return r;
}
template <int NW> __device__ __forceinline__ WKey<NW> wk_shfl_xor(const WKey<NW> &a, u32 d) {
    WKey<NW> r;
#pragma unroll
    for(int i = 0; i < NW; ++i) r.w[i] = __shfl_xor_sync(FULL, a.w[i], d);
    return r;
}
// e[0..3]: this lane's elements (wk_inf() past the m-th). o[i] = minimum over elements [g-W+1, g], g = 4*lane+i.
// Returns the mask of i with W-1 <= g < m (the windows QueueMap::next_value returns an element for, qmap.h:79-87).
template <int NW>
__device__ __forceinline__ u32 window_min_tile(const WKey<NW> (&e)[PPL], u32 m, u32 W, u32 lane, WKey<NW> (&o)[PPL]) {
    typedef WKey<NW> K;
    K P[PPL], S[PPL];
    P[0] = e[0]; P[1] = wk_min(P[0], e[1]); P[2] = wk_min(P[1], e[2]); P[3] = wk_min(P[2], e[3]);
    S[3] = e[3]; S[2] = wk_min(e[2], S[3]); S[1] = wk_min(e[1], S[2]); S[0] = P[3];
    int f[PPL]; u32 c[PPL], off[PPL];
    u32 call = 0;
#pragma unroll
    for(int i = 0; i < PPL; ++i) {
        const int t = i - (int)W + 1;                                  // window start relative to 4*lane (<= 0)
        f[i] = t >= 0 ? 0 : -(int)(((u32)(-t) + 3u) >> 2);              // floor(t / 4)
        off[i] = (u32)(t - 4 * f[i]);                                  // first element's index inside lane l+f
        c[i] = f[i] < 0 ? (u32)(-f[i] - 1) : 0u;                        // whole lanes between the two partial ones
        call |= c[i];
    }
    // the partial first lane: a suffix of lane l+f's elements
    K R[PPL];
#pragma unroll
    for(int i = 0; i < PPL; ++i) {
        R[i] = wk_inf<NW>();
        if(f[i] < 0) {
            const K sel = off[i] == 0 ? S[0] : off[i] == 1 ? S[1] : off[i] == 2 ? S[2] : S[3];
            R[i] = wk_shfl_up(sel, (u32)(-f[i]));
        }
    }
    K U = S[0];                                                        // level j: minimum of the lane minima of lanes l-2^j+1 .. l
#pragma unroll
    for(int j = 0; j < 5; ++j) {
        if((call >> j) == 0) break;                                    // warp-uniform (W is)
#pragma unroll
        for(int i = 0; i < PPL; ++i)
            if((c[i] >> j) & 1u) R[i] = wk_min(R[i], wk_shfl_up(U, 1u + (c[i] & ((1u << j) - 1u))));
        if((call >> (j + 1)) != 0) U = wk_min(U, wk_shfl_up(U, 1u << j));
    }
    u32 mask = 0;
#pragma unroll
    for(int i = 0; i < PPL; ++i) {
        const u32 g = PPL * lane + i;
        if(f[i] < 0) o[i] = wk_min(R[i], P[i]);
        else {                                                         // W <= i+1: the window lies inside this lane
            K a = e[i];
#pragma unroll
            for(int q = 0; q < PPL; ++q) if(q < i && (u32)q >= off[i]) a = wk_min(a, e[q]);
            o[i] = a;
        }
        if(g + 1 >= W && g < m) mask |= 1u << i;
    }
    return mask;
}

// inverse of lex_score (bns_device.cuh): xor, rotate and an odd multiply are all invertible
__device__ __forceinline__ u64 lex_unscore(u64 s) {
    constexpr u64 MI = inv_odd(0x9a98567ed20c127dull);
    s ^= 0x691a9d706391077aull;
    s = (s >> 31) | (s << 33);
    return (s * MI) ^ 0x533f8c2151b20f97ull;
}

// score of a window element; same values as score_of() (bns_kernels.cu). With the saturating cast every k-mer that holds
// two different bases has H + 0.001 <= -0.138 (k <= 32), so kmer / (H + 0.001) <= -1.44 for kmer >= 2 (and kmer = 1 is
// A..AC, |H + 0.001| < 1): the cast gives ~0 without any floating point. Homopolymers (H = 0) go the exact way.
__device__ __forceinline__ u64 score_lean(const EncParams &cP, u64 x, u64 kmask) {
    if(cP.score_kind == SC_LEX) return lex_score(x);
    if(cP.score_kind == SC_ENT_ROLL && !cP.cast_wrap && ((x ^ (x >> 2)) & (kmask >> 2)) != 0) return ~0ull;
    return score_of(cP, x);
}

// resolve_tree (util.h:831-869) when the record's counts live one value id per lane (SV variants: dictionaries of at most 32
// values): score(t) = sum of the u16 counts over t's root path (Euler intervals), unique maximum wins, ties -> lca of the tied
// taxa. The ties are folded in value-id order, not first-seen order: lca is commutative and associative on the loader's
// well-formed taxonomy, so the result is the same.
__device__ __forceinline__ u32 resolve_lanes(u32 cnt_lane, const uint4 *vi, const TaxView &X, u32 n_values, u32 lane) {
    const u32 present = __ballot_sync(FULL, cnt_lane != 0);
    if(!present) return 0;
    if(!(present & (present - 1))) return vi[__ffs(present) - 1].w;
    const u32 ti = vi[min(lane, n_values - 1)].x;
    u32 sc = 0;
    for(u32 m = present; m; m &= m - 1) {
        const u32 u = __ffs(m) - 1;
        const uint4 iu = vi[u];
        const u32 cu = __shfl_sync(FULL, cnt_lane, u) & 0xffffu;         // the reference's counts are u16
        if(iu.x <= ti && ti < iu.y) sc += cu;
    }
    const bool mine = (present >> lane) & 1u;
    const u32 best = __reduce_max_sync(FULL, mine ? sc : 0u);
    u32 tied = __ballot_sync(FULL, mine && sc == best);
    u32 node = 0, ntied = 0, first_id = 0;
    while(tied) {
        const u32 l = __ffs(tied) - 1;
        tied &= tied - 1;
        const uint4 inf = vi[l];
        if(ntied == 0) { node = inf.z; first_id = l; }
        else {                                                         // lca(node, l): climb until the interval covers l (util.h:634-663)
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
    return ntied == 1 ? vi[first_id].w : X.node_info[node].w;
}

// MODE: LEAN_U every (canonical if CANON) k-mer                      encoder.h:240-272 (+ :218-232)
//       LEAN_K canonical k-mer at every position, invalid -> 0, windowed  encoder.h:211-217,622-628   (CANON must be true)
//       LEAN_R valid forward k-mers, windowed over the compacted sequence, tail flush, canonical on emit if CANON
//                                                                   encoder.h:273-353
//       LEAN_S every valid spaced k-mer (comb of at most 45 bases, window of one)        encoder.h:233-239,547-592,616-621
// KT: compile-time k (0 = use P.k). COUNTS: per-record hit / missing counts are wanted.
// Windowed modes handle records of at most TILE window elements here; longer ones (and the 32-T restart quirk of
// encoder.h:283) are appended to defer_idx and done by the generic stream kernel right after this one.
// KEY (windowed modes): how a window element is ordered -- LEAN_KEY_PAIR (score, k-mer) in full, LEAN_KEY_LEX the Lex score
// alone, LEAN_KEY_ELEM the k-mer alone (scores non-decreasing in the k-mer); chosen by pick_lean().
// LOC: the table is in LAYOUT_MINIMIZER (bns_device.cuh): bucket and remainder come from loc_encode instead of mix64.
// RUNS: the ordered hit list of every record leaves run-length encoded (what the Kraken-style text prints, classifier.h:46-61):
// (taxid << 32 | run length) words at runs_out[run_pos_out[r] .. + n_runs_out[r]). A record's runs are gathered in shared
// memory and placed into a stretch of runs_out the warp reserved with one atomic per RUN_BLOCK entries (runs_total = entries
// handed out so far, gaps included).
// PK: the bases arrive 2-bit packed by the host (bns_pack.cpp; 4x fewer bytes over PCIe): `bases` is the stream of 16-bit units
// (8 bases each, first base in the top bits; base index g of the batch sits in unit (g - pk.base0) >> 3), pk.susp one bit per
// unit ("holds a byte that is not ACGTacgt"), pk.exc the sorted (unit << 8 | invalid-bit mask) words of those units.
__device__ __forceinline__ u32 pk_invalid_mask(const PackedIn &pk, u64 unit) {     // rare: binary search of the exceptions
    u32 lo = 0, hi = pk.n_exc;
    while(lo < hi) { const u32 mid = (lo + hi) >> 1; if((pk.exc[mid] >> 8) < unit) lo = mid + 1; else hi = mid; }
    return (lo < pk.n_exc && (pk.exc[lo] >> 8) == unit) ? (u32)(pk.exc[lo] & 0xffu) : 0u;
}
// SV: the value dictionary has at most 32 entries: lane v keeps the record's count of value id v in a register (linear::counter::add
// is one predicated add, no list in shared memory), resolve_lanes() reads the counts by shuffle.
template <int MODE, bool CANON, int KT, bool COUNTS, int KEY, bool LOC, bool RUNS, bool PK = false, bool SV = false>
__global__ void __launch_bounds__(LEAN_WARPS * 32, BNS_CLASSIFY_U_MIN_CTAS)
bns_classify_u_kernel(const __grid_constant__ EncParams P, const char *__restrict__ bases, const u64 *__restrict__ offsets,
                      u64 n_records, TableView T, TaxView X, u32 *__restrict__ taxon_out, u32 *__restrict__ nhit_out,
                      u32 *__restrict__ nmiss_out, unsigned long long *__restrict__ counters, u32 *__restrict__ status,
                      u32 *__restrict__ defer_idx, unsigned long long *__restrict__ defer_cnt, u32 fixed_len, u64 fixed_base,
                      u32 mates, u32 *__restrict__ mate1_out, u64 *__restrict__ runs_out, u64 runs_cap,
                      unsigned long long *__restrict__ runs_total, u64 *__restrict__ run_pos_out, u32 *__restrict__ n_runs_out,
                      const PackedIn pk) {
    // n_records counts SEQUENCES here: a record is `mates` (1 or 2) consecutive sequences sharing one taxon counter
    // (classify_seq encodes the second mate into the same counter, classifier.h:233-236); a batch of 32 sequences holds
    // whole records. mate1_out: k-mers the first mate produced (classifier.h:232's first ambig_count term).
    // fixed_len != 0: every record has that many bases and record r starts at fixed_base + r * fixed_len; `offsets` is
    // not read (the host did not even copy it: 8 of the 158 bytes per 150 bp read that cross PCIe)
    __shared__ __align__(16) uint4 s_vi[VI_CAP];
    __shared__ __align__(8) unsigned long long s_mbar;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5, n_cta_warps = blockDim.x >> 5;
    const u32 k = KT ? (u32)KT : P.k;
    constexpr bool CANON_ELEM = MODE == LEAN_K || (MODE == LEAN_U && CANON);
    const bool staged = T.n_values > 0 && T.n_values <= (u32)VI_CAP;
    if(staged) tma_stage_val_info(s_vi, X.val_info, T.n_values * (u32)sizeof(uint4), &s_mbar);
    WarpSmem S;                                                        // only the distinct-taxon lists are used here
    S.ids = (u32 *)(g_smem + (size_t)wid * 4 * AGG_CAP * sizeof(u32));
    S.cnt = S.ids + AGG_CAP; S.tin = S.cnt + AGG_CAP; S.tout = S.tin + AGG_CAP;
    ClassifySink<false> sink;
    sink.T = T;
    sink.vi = staged ? s_vi : X.val_info;
    sink.begin(nullptr);

    ProbeConst Pc;
    Pc.slots = (const char *)T.slots;
    Pc.b = T.bucket_bits;
    Pc.idx_shift = 32 - T.bucket_bits;
    Pc.hm = ~0u << T.tag_shift;
    Pc.tag_shift = T.tag_shift; Pc.fmt_bits = T.fmt.fmt_bits; Pc.max_disp = T.fmt.max_disp(); Pc.layout = T.fmt.layout;
    Pc.flags_all = ((1u << T.tag_shift) - 1) & ~T.val_mask;
    Pc.flag_shift = T.flag_shift; Pc.flag_mask = T.flag_mask;
    Pc.val_mask = T.val_mask;
    const u32 c = MODE == LEAN_S ? P.c : k;                            // bases a k-mer spans (the comb)
    const u32 span = TILE + c - 1;
    const u32 down = 64 - 2 * k;
    const u32 kmask_lo = (u32)(~0ull >> down), kmask_hi = (u32)((~0ull >> down) >> 32);
    const u64 kmask = ~0ull >> down;
    const u32 W = P.W;
    u64 comb = 0;                                                      // LEAN_S: the comb's bases, first base at bit c-1
    if(MODE == LEAN_S)
        for(u32 sg = 0; sg < P.n_seg; ++sg) comb |= ((1ull << P.seg_len[sg]) - 1) << (c - P.seg_off[sg] - P.seg_len[sg]);

    // Per-warp staging area in shared memory, filled by cp.async (LDGSTS): no registers are held across the HBM latency
    // and -- unlike a register prefetch -- the wait is a cp.async group wait, not a scoreboard the compiler may share
    // with other loads (ncu showed the register prefetch of v2 stalling a full DRAM latency per record for that reason).
    //   s_off[2][RB+1]  offsets of the current / next batch of records
    //   s_rd[2][32]     8 bytes per lane of the first tile of the current / next record
    // (the run buffer of the RUNS variants follows the warp's staging area: one base address for both)
    constexpr u32 WARP_STAGE_BYTES = LEAN_STAGE_BYTES + (RUNS ? RUNBUF * sizeof(u64) : 0);
    u64 *s_off = (u64 *)(g_smem + (size_t)n_cta_warps * 4 * AGG_CAP * sizeof(u32) + (size_t)wid * WARP_STAGE_BYTES);
    uint2 *s_rd = (uint2 *)(s_off + 2 * (RB + 2));
    // run lists: this warp's record buffer, its reserved stretch of runs_out, and the run that is still open
    u64 *s_run = s_off + LEAN_STAGE_BYTES / sizeof(u64);
    u64 blk_pos = 0;
    u32 blk_left = 0, run_val = VAL_MISS, run_len = 0, n_runs_rec = 0;
    bool run_direct = false;                                          // this record's runs go straight to runs_out[blk_pos ...]
    u64 my_rpos = 0; u32 my_nruns = 0;
    auto put_run = [&](u32 widx, u32 val, u32 len) {                   // run number widx of the current record
        const u64 e = ((u64)sink.vi[val].w << 32) | len;
        if(!run_direct) s_run[widx] = e;
        else if(blk_pos + widx < runs_cap) runs_out[blk_pos + widx] = e;
    };
    auto reserve_runs = [&](u32 need) {                                // make the warp's stretch hold `need` more entries
        if(need > blk_left) {
            const u32 take = max(need, RUN_BLOCK);
            unsigned long long base = 0;
            if(lane == 0) base = atomicAdd(runs_total, (unsigned long long)take);
            blk_pos = __shfl_sync(FULL, base, 0);
            blk_left = take;
            if(blk_pos + take > runs_cap && lane == 0) atomicOr(status, 1u);
        }
    };
    s_rd[lane] = make_uint2(0x41414141u, 0x41414141u);                 // lanes past a tile read 'A's: code 0, never "suspicious"
    s_rd[32 + lane] = make_uint2(0x41414141u, 0x41414141u);
    __syncwarp();
    const u64 nwarps = (u64)gridDim.x * n_cta_warps;
    const u64 n_batches = (n_records + RB - 1) / RB;
    u64 bt = (u64)blockIdx.x * n_cta_warps + wid;
    auto async8 = [](void *dst, const void *src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((u32)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    auto async_commit = [] { asm volatile("cp.async.commit_group;" ::: "memory"); };
    auto async_wait_all = [] { asm volatile("cp.async.wait_group 0;" ::: "memory"); };
    // offsets[batch*RB .. batch*RB + RB] -> s_off[buf]
    auto fetch_offsets = [&](u64 batch, u32 buf) {
        const u64 r = batch * RB + lane;
        if(batch < n_batches) {
            if(fixed_len) {
                if(r <= n_records) s_off[buf * (RB + 2) + lane] = fixed_base + r * fixed_len;
                if(lane == 0 && r + RB <= n_records) s_off[buf * (RB + 2) + RB] = fixed_base + (r + RB) * fixed_len;
            } else {
                if(r <= n_records) async8(s_off + buf * (RB + 2) + lane, offsets + r);
                if(lane == 0 && r + RB <= n_records) async8(s_off + buf * (RB + 2) + RB, offsets + r + RB);
            }
        }
    };
    // the 8-byte block of a record's first tile this lane stages -> s_rd[buf]: bases [rb, rb + min(rl, span))
    auto fetch_tile = [&](u64 rb, u32 rl, u32 buf) {
        if(PK) {
            // packed input: the (at most 24) units of the tile are bytes [2 * U0, 2 * (U0 + nun + 1)) of the unit stream, fetched as
            // up to seven aligned 8-byte blocks; entry 8 of the buffer takes the two suspicious-bit words the units fall into
            if(!rl) return;
            const u64 g = rb - pk.base0, U0 = g >> 3, byte0 = (2 * U0) & ~7ull;
            const u32 nun = ((u32)(g & 7u) + min(rl, span) + 7) >> 3;
            const u32 nblk = (u32)((2 * (U0 + nun + 1) - byte0 + 7) >> 3);
            if(lane < nblk) async8(s_rd + buf * 32 + lane, bases + byte0 + 8 * lane);
            else if(lane == 8 || lane == 9)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((u32)__cvta_generic_to_shared((u32 *)(s_rd + buf * 32 + 8) + (lane - 8))),
                             "l"(pk.susp + (U0 >> 5) + (lane - 8)) : "memory");
            return;
        }
        const char *a0 = bases + rb;
        const u32 shift = (u32)((uintptr_t)a0 & 7u);
        const u32 nblk = rl ? (shift + min(rl, span) + 7) >> 3 : 0u;      // <= 22 for k <= 32
        if(lane < nblk) async8(s_rd + buf * 32 + lane, a0 - shift + 8 * lane);
        else s_rd[buf * 32 + lane] = make_uint2(0x41414141u, 0x41414141u);
    };
    // later tiles of a long record: plain loads (rare)
    auto tile_block = [&](u64 rb, u32 rl) -> uint2 {
        const char *a0 = bases + rb;
        const u32 shift = (u32)((uintptr_t)a0 & 7u);
        const u32 nblk = rl ? (shift + min(rl, span) + 7) >> 3 : 0u;
        uint2 v = make_uint2(0x41414141u, 0x41414141u);
        if(lane < nblk) v = __ldg(reinterpret_cast<const uint2 *>(a0 - shift) + lane);
        return v;
    };
    auto rec_len = [](u64 b, u64 e) -> u32 { const u64 len = e - b; return len < 0xffffffffull ? (u32)len : 0xffffffffu; };
    u32 pb = 0;                                                        // offsets buffer of the current batch
    fetch_offsets(bt, 0);
    async_commit();
    async_wait_all();
    __syncwarp();
    u64 rb = 0;                                                        // the record being processed (warp-uniform)
    u32 L = 0;
    if(bt < n_batches) { rb = s_off[0]; L = rec_len(rb, s_off[1]); }
    u32 tb = 0;                                                        // tile buffer of the current record
    fetch_tile(rb, L, 0);
    async_commit();
    if(staged) mbar_wait(&s_mbar, 0);

    for(; bt < n_batches; bt += nwarps) {
        fetch_offsets(bt + nwarps, pb ^ 1);                            // lands while this batch is processed
        async_commit();
        const u64 r0 = bt * RB;
        const u32 nrec = (u32)min((u64)RB, n_records - r0);
        const bool have_next_batch = bt + nwarps < n_batches;
        u32 my_taxon = 0, my_hit = 0, my_miss = 0, my_def = 0, my_m1 = 0;
        // ---- per-record state: linear::counter with its first key in registers --------------------------------------
        u32 nd = 0, id0 = 0, cnt0 = 0, n_hit = 0, n_emit = 0;
        u32 cnt_lane = 0;                                              // SV: this record's hits of value id `lane`
        bool spilled = false, deferred = false;
        const u32 msh = COUNTS ? mates - 1 : 0u;                       // record of sequence j: j >> msh (pairs always run a COUNTS variant)
        for(u32 j = 0; j < nrec; ++j) {
            const bool first_mate = (j & msh) == 0, last_mate = (j & msh) == msh;
            // this record's first tile (requested one record ago) and, from the second record on, the next batch's offsets
            // (the newest group at j == 0 is the next batch's offsets, requested a moment ago: not needed yet)
            if(j == 0) asm volatile("cp.async.wait_group 1;" ::: "memory"); else async_wait_all();
            __syncwarp();
            const uint2 pre = s_rd[tb * 32 + lane];
            u32 pre_c16 = 0, pre_susp = 0;                                 // PK: this lane's unit of the first tile and its suspicious bit
            if(PK && L != 0) {
                const u64 g = rb - pk.base0, U0 = g >> 3;
                const u32 nun = ((u32)(g & 7u) + min(L, span) + 7) >> 3;
                if(lane <= nun) {                                      // one more than the tile covers: a lane's word takes the next lane's unit
                    pre_c16 = reinterpret_cast<const unsigned short *>(s_rd + tb * 32)[(u32)(U0 & 3u) + lane];
                    const u32 *sw = reinterpret_cast<const u32 *>(s_rd + tb * 32 + 8);
                    const u64 un = U0 + lane;
                    pre_susp = (sw[(un >> 5) - (U0 >> 5)] >> (un & 31u)) & 1u;
                }
            }
            // request the first tile of the record after this one before working on this one
            u64 xb = 0; u32 xl = 0;
            {
                const u64 *o = nullptr;
                if(j + 1 < nrec) o = s_off + pb * (RB + 2) + j + 1;
                else if(have_next_batch) o = s_off + (pb ^ 1) * (RB + 2);
                if(o) { xb = o[0]; xl = rec_len(xb, o[1]); }
            }
            fetch_tile(xb, xl, tb ^ 1);
            async_commit();
            if(first_mate) { nd = 0; id0 = 0; cnt0 = 0; n_hit = 0; n_emit = 0; deferred = false; if(SV) cnt_lane = 0; }
            if(RUNS && first_mate) { run_val = VAL_MISS; run_len = 0; n_runs_rec = 0; run_direct = false; }
            if(L == 0xffffffffu) { if(lane == 0) atomicOr(status, 8u); }
            else if(deferred) {}                                       // the first mate already sent the record to the generic kernel
            else if((MODE == LEAN_K || MODE == LEAN_R) && L >= k && L - k + 1 > (u32)TILE) deferred = true;   // more than one tile of window elements
            else if(L >= c) {
                const u32 npos = L - c + 1;
                for(u32 p0 = 0; p0 < npos; p0 += TILE) {
                    // ---- stage: 8 bases per lane -> 16 bits per lane -> one 16-base word per lane ----------------
                    uint2 v = make_uint2(0x41414141u, 0x41414141u);
                    u32 shift, c16, susp;
                    u64 pk_unit = 0;
                    if(PK) {
                        // this lane's unit of the packed stream and its "not all ACGT" bit
                        const u64 g = rb - pk.base0 + p0;
                        shift = (u32)(g & 7u);
                        pk_unit = (g >> 3) + lane;
                        const u32 nun = (shift + min(L - p0, span) + 7) >> 3;          // units the tile covers (<= 23)
                        c16 = pre_c16; susp = pre_susp;
                        if(p0 && lane <= nun) {                                        // later tiles of a long record: plain loads (rare)
                            c16 = __ldg(reinterpret_cast<const unsigned short *>(bases) + pk_unit);
                            susp = (__ldg(pk.susp + (pk_unit >> 5)) >> (pk_unit & 31u)) & 1u;
                        } else if(p0) { c16 = 0; susp = 0; }
                    } else {
                        v = p0 ? tile_block(rb + p0, L - p0) : pre;
                        shift = (u32)((uintptr_t)(bases + rb + p0) & 7u);
                        susp = pack8_fast(v, c16);
                    }
                    const bool slow = __any_sync(FULL, susp != 0);
                    const u32 word = (c16 << 16) | __shfl_down_sync(FULL, c16, 1);   // bases [8*lane, 8*lane+16) of the tile
                    const u32 q0 = shift + PPL * lane, ci = q0 >> 3, s = (q0 & 7u) * 2u;
                    const u32 w0 = __shfl_sync(FULL, word, ci), w1 = __shfl_sync(FULL, word, ci + 2),
                              w2 = __shfl_sync(FULL, word, ci + 4), w3 = __shfl_sync(FULL, word, ci + 6);
                    const u32 left = npos - p0;
                    const u32 nlive = left > PPL * lane ? min((u32)PPL, left - PPL * lane) : 0u;
                    u32 mask = (1u << nlive) - 1;
                    if(slow) {                                         // some staged byte is not ACGTacgt (rare)
                        u32 b8 = 0, cc;
                        if(PK) { if(susp) b8 = pk_invalid_mask(pk, pk_unit); }
                        else pack8(v, cc, b8);
                        const u32 bw = (b8 << 8) | __shfl_down_sync(FULL, b8, 1);    // invalid bits of the same 16 bases, first at bit 15
                        const u32 b0 = __shfl_sync(FULL, bw, ci), b1 = __shfl_sync(FULL, bw, ci + 2),
                                  b2 = __shfl_sync(FULL, bw, ci + 4), b3 = __shfl_sync(FULL, bw, ci + 6);
                        const u64 B = ((u64)b0 << 48) | ((u64)b1 << 32) | ((u64)b2 << 16) | (u64)b3;   // coordinate 8*ci at bit 63
#pragma unroll
                        for(int i = 0; i < PPL; ++i)
                            if(MODE == LEAN_S ? ((((B << ((q0 & 7u) + i)) >> (64 - c)) & comb) != 0)
                                              : (((B << ((q0 & 7u) + i)) >> (64 - k)) != 0)) mask &= ~(1u << i);
                        if(COUNTS && (MODE == LEAN_U || MODE == LEAN_S)) n_emit += __reduce_add_sync(FULL, __popc(mask));
                    } else if(COUNTS && (MODE == LEAN_U || MODE == LEAN_S)) n_emit += min(left, (u32)TILE);
                    // ---- the lane's four k-mers (and reverse complements) out of one 96-bit window --------------
                    const u32 A = __funnelshift_l(w1, w0, s), B_ = __funnelshift_l(w2, w1, s), C = __funnelshift_l(w3, w2, s);
                    if(MODE == LEAN_R && P.t_restart) {
                        // for_each_uncanon_unspaced_windowed tests the 64-bit word against ~0 before masking (encoder.h:283):
                        // 32 consecutive T restart the rolling state. Leave such records to the generic kernel.
                        bool run = false;
#pragma unroll
                        for(int i = 0; i < PPL; ++i) run |= (__funnelshift_l(B_, A, 2 * i) & __funnelshift_l(C, B_, 2 * i)) == ~0u;
                        if(__any_sync(FULL, run && lane * PPL < left)) { deferred = true; break; }
                    }
                    u32 R0 = 0, R1 = 0, R2 = 0;
                    if(CANON_ELEM) {
                        R0 = __brev(C); R1 = __brev(B_); R2 = __brev(A);
                        R0 = ~(((R0 >> 1) & 0x55555555u) | ((R0 & 0x55555555u) << 1));
                        R1 = ~(((R1 >> 1) & 0x55555555u) | ((R1 & 0x55555555u) << 1));
                        R2 = ~(((R2 >> 1) & 0x55555555u) | ((R2 & 0x55555555u) << 1));
                    }
                    u32 xls[PPL], xhs[PPL];
                    u32 fwdm = 0xfu;                                   // bit i: k-mer i is used as read (not reverse-complemented)
                    if(MODE == LEAN_S) {
                        // Encoder::kmer(pos) (encoder.h:547-592) for the lane's four positions at once: per contiguous run of the
                        // comb, the run's bases plus three more come out of the 96-bit window as one 64-bit word; position i's piece
                        // is that word shifted by 2(3-i).
                        u64 xs[PPL] = {0, 0, 0, 0};
                        for(u32 sg = 0; sg < P.n_seg; ++sg) {
                            const u32 off = P.seg_off[sg], len = P.seg_len[sg];              // len <= 29 (checked by the dispatcher)
                            const u32 bo = 2 * off, wi = bo >> 5, sb = bo & 31u;
                            const u32 x0 = wi == 0 ? A : wi == 1 ? B_ : C, x1 = wi == 0 ? B_ : wi == 1 ? C : 0u, x2 = wi == 0 ? C : 0u;
                            const u64 V = ((((u64)__funnelshift_l(x1, x0, sb)) << 32) | __funnelshift_l(x2, x1, sb)) >> (58 - 2 * len);
                            const u64 pm = (1ull << (2 * len)) - 1;
#pragma unroll
                            for(int i = 0; i < PPL; ++i) xs[i] = (xs[i] << (2 * len)) | ((V >> (6 - 2 * i)) & pm);
                        }
#pragma unroll
                        for(int i = 0; i < PPL; ++i) {
                            xls[i] = (u32)xs[i]; xhs[i] = (u32)(xs[i] >> 32);
                            if(P.filter_none && xs[i] == ~0ull) mask &= ~(1u << i);          // `!= ENCODE_OVERFLOW` (k = 32, all T)
                        }
                    } else
#pragma unroll
                    for(int i = 0; i < PPL; ++i) {
                        u32 xl, xh;                                                          // forward k-mer: window bits [2i, 2i+2k)
                        if(KT == 31) {                                                       // ... = bits [2i, 2i+62): shifts fold
                            if(i == 0) { xl = __funnelshift_r(B_, A, 2); xh = A >> 2; }
                            else if(i == 1) { xl = B_; xh = A & 0x3fffffffu; }
                            else { xl = __funnelshift_l(C, B_, 2 * i - 2); xh = __funnelshift_l(B_, A, 2 * i - 2) & 0x3fffffffu; }
                        } else {
                            const u32 fh0 = __funnelshift_l(B_, A, 2 * i), fl0 = __funnelshift_l(C, B_, 2 * i);
                            if(down < 32) { xl = __funnelshift_r(fl0, fh0, down); xh = fh0 >> down; }
                            else { xl = fh0 >> ((down - 32) & 31u); xh = 0; }                      // k <= 16
                        }
                        if(CANON_ELEM) {
                            const u32 rl = __funnelshift_r(R2, R1, 2 * i) & kmask_lo, rh = __funnelshift_r(R1, R0, 2 * i) & kmask_hi;
                            const u64 f64 = ((u64)xh << 32) | xl, r64 = ((u64)rh << 32) | rl;
                            const bool lt = f64 < r64;
                            xl = lt ? xl : rl; xh = lt ? xh : rh;
                            if(LOC && !lt) fwdm &= ~(1u << i);
                        }
                        xls[i] = xl; xhs[i] = xh;
                    }
                    if(MODE == LEAN_K || MODE == LEAN_R) {
                        // ---- window elements -> minimizers (QueueMap, qmap.h:79-96) ---------------------------------
                        const u32 livem = (1u << nlive) - 1;                 // positions inside the record
                        constexpr int NW = KEY == LEAN_KEY_PAIR ? 4 : 2;
                        WKey<NW> e[PPL], o[PPL];
                        u32 m = npos;                                        // elements pushed
#pragma unroll
                        for(int i = 0; i < PPL; ++i) {
                            if(MODE == LEAN_K && !(mask >> i & 1u)) { xls[i] = 0; xhs[i] = 0; }   // invalid -> ~0 -> canonical 0 (encoder.h:622-628)
                            const bool on = MODE == LEAN_K ? (livem >> i & 1u) : (mask >> i & 1u);
                            e[i] = wk_inf<NW>();
                            if(on) {
                                const u64 xe = ((u64)xhs[i] << 32) | xls[i];
                                if(KEY == LEAN_KEY_ELEM) { e[i].w[0] = xhs[i]; e[i].w[1] = xls[i]; }
                                else {
                                    const u64 sc = KEY == LEAN_KEY_LEX ? lex_score(xe) : score_lean(P, xe, kmask);
                                    e[i].w[0] = (u32)(sc >> 32); e[i].w[1] = (u32)sc;
                                    if(KEY == LEAN_KEY_PAIR) { e[i].w[NW - 2] = xhs[i]; e[i].w[NW - 1] = xls[i]; }
                                }
                            }
                        }
                        if(MODE == LEAN_R && slow) {                         // only valid k-mers push: compact them (rare)
                            u32 tot;
                            u32 idx = warp_excl_scan(__popc(mask), lane, tot);
                            uint4 *scratch = (uint4 *)S.tin;                 // TILE x 16 bytes = tin + tout: only resolve() uses them (ids / cnt may hold mate 1's list)
#pragma unroll
                            for(int i = 0; i < PPL; ++i)
                                if(mask >> i & 1u) scratch[idx++] = make_uint4(e[i].w[0], e[i].w[1], e[i].w[NW - 2], e[i].w[NW - 1]);
                            __syncwarp();
#pragma unroll
                            for(int i = 0; i < PPL; ++i) {
                                const u32 g = PPL * lane + i;
                                e[i] = wk_inf<NW>();
                                if(g < tot) { const uint4 t4 = scratch[g]; e[i].w[0] = t4.x; e[i].w[1] = t4.y; e[i].w[NW - 2] = t4.z; e[i].w[NW - 1] = t4.w; }
                            }
                            __syncwarp();
                            m = tot;
                        }
                        mask = window_min_tile<NW>(e, m, W, lane, o);
                        if(MODE == LEAN_R && P.tail_flush && m > 0 && m < W) {
                            // a queue that never filled emits its minimum once (encoder.h:304-305,343-344)
                            WKey<NW> a = wk_min(wk_min(e[0], e[1]), wk_min(e[2], e[3]));
#pragma unroll
                            for(int d = 16; d; d >>= 1) a = wk_min(a, wk_shfl_xor(a, d));
                            o[0] = a;
                            mask = lane == 0 ? 1u : 0u;
                        }
#pragma unroll
                        for(int i = 0; i < PPL; ++i) {
                            if(KEY == LEAN_KEY_LEX) {
                                const u64 xe = lex_unscore(((u64)o[i].w[0] << 32) | o[i].w[1]);
                                xls[i] = (u32)xe; xhs[i] = (u32)(xe >> 32);
                            } else { xls[i] = o[i].w[NW - 1]; xhs[i] = o[i].w[NW - 2]; }
                            if(P.filter_none && (xls[i] & xhs[i]) == ~0u) mask &= ~(1u << i);      // `!= ENCODE_OVERFLOW`
                            if(MODE == LEAN_R && CANON) {                                          // canonical on emit (encoder.h:347-353)
                                const u64 cx = canonical(((u64)xhs[i] << 32) | xls[i], k);
                                xls[i] = (u32)cx; xhs[i] = (u32)(cx >> 32);
                            }
                        }
                        if(COUNTS) n_emit += __reduce_add_sync(FULL, __popc(mask));
                    }
                    u32 cand[PPL], nok = 0;                            // nok bit i: k-mer i was not found
                    if(LOC) {
                    // ---- LAYOUT_MINIMIZER: (home unit, left-aligned remainder) of the four keys --------------------------
                    u32 hb[PPL], th[PPL], tl[PPL];
                    // All four k-mers of every lane at once (records of at most 120 k-mers per tile: lanes 0..29 own k-mers, and
                    // lane l's 19 16-mers come from lanes l..l+2): each lane hashes the canonical 16-mers at offsets 0..3 and
                    // 8..11 of its window, two SHFL.DOWN steps bring in the other eleven, and the minimum of (27-bit hash :
                    // position) over each k-mer's 16 positions is a chain of 32-bit VIMNMX.
                    const bool loc_fast = MODE == LEAN_U && KT == 31 && left <= 120u;   // the offsets below are those of k = 31
                    if(loc_fast) {
                        u32 key[19];                                   // hashed canonical 16-mers at window offsets 0..18
#pragma unroll
                        for(int i = 0; i < 4; ++i) {
                            const u32 fa = __funnelshift_l(B_, A, 2 * i), fb = __funnelshift_l(B_, A, 16 + 2 * i);
                            key[i] = loc_hash(min(fa, rc16(fa))) & ~31u;
                            key[8 + i] = loc_hash(min(fb, rc16(fb))) & ~31u;
                        }
#pragma unroll
                        for(int i = 0; i < 4; ++i) {
                            key[4 + i] = __shfl_down_sync(FULL, key[i], 1);
                            key[12 + i] = __shfl_down_sync(FULL, key[8 + i], 1);
                            if(i < 3) key[16 + i] = __shfl_down_sync(FULL, key[8 + i], 2);
                        }
#pragma unroll
                        for(int j = 0; j < 19; ++j) key[j] |= (u32)j;                        // leftmost wins ties; ^31: rightmost
                        u32 coreL = key[3], coreR = key[3] ^ 31u;
#pragma unroll
                        for(int j = 4; j <= 15; ++j) { coreL = min(coreL, key[j]); coreR = min(coreR, key[j] ^ 31u); }
                        const u32 bl = Pc.b - LOC_GB, lmask = (1u << bl) - 1, xmask = ~((1u << (bl - 1)) - 1);
#pragma unroll
                        for(int i = 0; i < PPL; ++i) {
                            u32 mL = coreL, mR = coreR;
#pragma unroll
                            for(int j = 0; j < 19; ++j)
                                if(j >= i && j <= i + 15 && (j < 3 || j > 15)) { mL = min(mL, key[j]); mR = min(mR, key[j] ^ 31u); }
                            // the k-mer's canonical string reads left to right when it is the read's strand, else right to left:
                            // window offset js of the winner -> position bp in the canonical k-mer
                            const bool fw = (fwdm >> i) & 1u;
                            const u32 bp = fw ? (mL & 31u) - (u32)i : (mR & 31u) + (u32)i - 16u;   // = 15 - ((31 - (mR & 31)) - i)
                            const u32 xl = xls[i], xh = xhs[i];
                            const u32 m16 = __funnelshift_r(xl, xh, 30u - 2u * bp);            // the minimizer as it stands in the canonical k-mer
                            const u32 r16 = rc16(m16), pl = loc_place(loc_hash(min(m16, r16)));
                            hb[i] = ((pl & lmask) << LOC_GB) | ((bp & 3u) << 1);
                            tl[i] = ((pl >> 1) & xmask) | (r16 < m16 ? 0x80000000u : 0u);
                            th[i] = (__funnelshift_l(xh << 2, xl, 2u * bp) << 2) | (bp >> 2);   // the other 15 bases (right : left) | position >> 2
                        }
                    } else {
#pragma unroll
                        for(int i = 0; i < PPL; ++i) {
                            const TableHash t = loc_encode(((u64)xhs[i] << 32) | xls[i], k, Pc.b);
                            hb[i] = (u32)t.home; th[i] = (u32)(t.tag >> 32); tl[i] = (u32)t.tag;
                        }
                    }
                    // ---- probe: the 64-byte home unit of two k-mers at a time (four LDG.256 in flight per lane; the lane's four
                    // k-mers mostly share their group, so the second pair finds its lines in L2 or on their way) -------------
                    u32 more = 0;                                      // bit i: k-mer i must follow its overflow chain
#pragma unroll
                    for(int rnd = 0; rnd < 2; ++rnd) {
                        u32 w[2][16];
#pragma unroll
                        for(int j = 0; j < 2; ++j) {
                            const char *up = Pc.slots + ((u64)hb[2 * rnd + j] << 5);
                            ld_bucket8(up, *reinterpret_cast<u32 (*)[8]>(&w[j][0]));
                            ld_bucket8(up + 32, *reinterpret_cast<u32 (*)[8]>(&w[j][8]));
                        }
#pragma unroll
                        for(int j = 0; j < 2; ++j) {
                            const int i = 2 * rnd + j;
                            // first slot whose high word is the key's, verified on the low word; two entries of a unit sharing
                            // their high word (2^-30 per pair) send the lane through the exact loop
                            u32 c = ~tl[i];
#pragma unroll
                            for(int sl = 7; sl >= 0; --sl) c = w[j][2 * sl + 1] == th[i] ? w[j][2 * sl] : c;
                            if(((c ^ tl[i]) & Pc.hm) != 0 && c != ~tl[i]) {
                                c = ~tl[i];
#pragma unroll
                                for(int sl = 7; sl >= 0; --sl)
                                    if(w[j][2 * sl + 1] == th[i] && ((w[j][2 * sl] ^ tl[i]) & Pc.hm) == 0) c = w[j][2 * sl];
                            }
                            cand[i] = c;
                            const u32 miss = min((c ^ tl[i]) & Pc.hm, 1u);
                            nok |= miss << i;
                            // the key's overflow flag in slot 0 of the unit: cleared = a key homed here lives further down the chain
                            more |= (miss & ~(w[j][0] >> (Pc.flag_shift + ((th[i] >> 2) & Pc.flag_mask)))) << i;
                        }
                    }
                    more &= mask;
                    // ---- overflow chains (a few keys per record): compacted over the warp, one key per lane ------------------
                    if(__any_sync(FULL, more != 0)) {
                        u32 tot;
                        const u32 at0 = warp_excl_scan(__popc(more), lane, tot);
                        uint4 *q = (uint4 *)S.tin;                     // TILE x 16 bytes = tin + tout: only resolve() uses them
                        u32 at = at0;
#pragma unroll
                        for(int i = 0; i < PPL; ++i) if(more >> i & 1u) q[at++] = make_uint4(hb[i], th[i], tl[i], 0u);
                        __syncwarp();
                        for(u32 base = 0; base < tot; base += 32) {
                            const u32 t = base + lane;
                            if(t < tot) {
                                const uint4 e = q[t];
                                bool exhausted;
                                u32 c = probe_chain_loc(Pc, e.x, e.y, e.z, exhausted);
                                if(exhausted && T.stash) {
                                    const u32 v = probe_stash(T, loc_decode(e.x, ((u64)e.y << 32) | e.z, k, Pc.b));
                                    if(v != VAL_MISS) c = (e.z & Pc.hm) | v;
                                }
                                q[t].w = c;
                            }
                        }
                        __syncwarp();
                        at = at0;
#pragma unroll
                        for(int i = 0; i < PPL; ++i)
                            if(more >> i & 1u) {
                                const u32 c = q[at++].w;
                                if(c != ~tl[i]) { nok &= ~(1u << i); cand[i] = c; }
                            }
                        __syncwarp();
                    }
                    } else {
                    // ---- LAYOUT_HASH: mix64 as (hl, hh); bucket and tag words derive from it ---------------------------------
                    u32 hl[PPL], hh[PPL], w[PPL][8];
#pragma unroll
                    for(int i = 0; i < PPL; ++i) {
                        // mix64 (bns_device.cuh) on 32-bit halves
                        const u64 x = (((u64)xhs[i] << 32) | xls[i]) * 0xd6e8feb86659fd93ull;
                        hh[i] = (u32)(x >> 32); hl[i] = (u32)x ^ hh[i];
                        ld_bucket8(Pc.slots + ((u64)(hh[i] >> Pc.idx_shift) << 5), w[i]);
                    }
                    // ---- match: first slot whose high word equals the tag's, verified on the low word (upper words are unique
                    // within a bucket) ------------------------------------------------------------------------------------------
#pragma unroll
                    for(int i = 0; i < PPL; ++i) {
                        const u32 th = __funnelshift_lc(hl[i], hh[i], Pc.b);
                        const u32 tl = __funnelshift_lc(0u, hl[i], Pc.b);
                        u32 c = ~tl;
                        c = w[i][7] == th ? w[i][6] : c;
                        c = w[i][5] == th ? w[i][4] : c;
                        c = w[i][3] == th ? w[i][2] : c;
                        c = w[i][1] == th ? w[i][0] : c;                                   // slots fill in order: first match wins
                        cand[i] = c;
                        nok += min((c ^ tl) & Pc.hm, 1u) << i;
                    }
                    // keys displaced from a full home bucket (rare): cheap conservative test first -- did any of the 128 home
                    // buckets overflow at all? -- then the exact one
                    if(__any_sync(FULL, ((w[0][0] & w[1][0] & w[2][0] & w[3][0]) & Pc.flags_all) != Pc.flags_all)) {
                        u32 more = 0;
#pragma unroll
                        for(int i = 0; i < PPL; ++i)
                            if(!((w[i][0] >> (Pc.flag_shift + (hl[i] & Pc.flag_mask))) & 1u)) more |= 1u << i;
                        more &= nok & mask;
                        while(__any_sync(FULL, more != 0)) {
                            if(more) {
                                const u32 i = __ffs(more) - 1;
                                more &= more - 1;
                                const u32 l = i == 0 ? hl[0] : i == 1 ? hl[1] : i == 2 ? hl[2] : hl[3];
                                const u32 h = i == 0 ? hh[0] : i == 1 ? hh[1] : i == 2 ? hh[2] : hh[3];
                                const u32 th = __funnelshift_lc(l, h, Pc.b), tl0 = __funnelshift_lc(0u, l, Pc.b);
                                const u32 c = probe_displaced32(Pc, h >> Pc.idx_shift, th, tl0);
                                if(c != ~tl0) {
                                    nok &= ~(1u << i);
                                    if(i == 0) cand[0] = c; else if(i == 1) cand[1] = c; else if(i == 2) cand[2] = c; else cand[3] = c;
                                }
                            }
                        }
                    }
                    }
                    // ---- hits -> per-record distinct-taxon counts (linear::counter::add, linear.h:229) -----------
                    u32 todo = ~nok & mask;
                    u32 bal = __ballot_sync(FULL, todo != 0);
                    if(bal) {
                        if(COUNTS) n_hit += __reduce_add_sync(FULL, __popc(todo));
#pragma unroll
                        for(int i = 0; i < PPL; ++i) cand[i] &= Pc.val_mask;
                        const u32 todo0 = todo, bal0 = bal;
                        bool first_round = true;
                        do {
                            const u32 leader = __ffs(bal) - 1;
                            u32 fv = cand[3];
                            if(todo & 4u) fv = cand[2];
                            if(todo & 2u) fv = cand[1];
                            if(todo & 1u) fv = cand[0];
                            const u32 vv = __shfl_sync(FULL, fv, leader);
                            u32 e = (cand[0] == vv ? 1u : 0u) | (cand[1] == vv ? 2u : 0u) | (cand[2] == vv ? 4u : 0u) | (cand[3] == vv ? 8u : 0u);
                            e &= todo;
                            todo ^= e;
                            const u32 total = __reduce_add_sync(FULL, __popc(e));
                            if(SV) { if(lane == vv) cnt_lane += total; }
                            else if(nd == 0) { id0 = vv; cnt0 = total; nd = 1; }
                            else if(!spilled && vv == id0) cnt0 += total;
                            else {
                                if(!spilled) {                         // second distinct taxon: move the list to shared memory
                                    if(lane == 0) { S.ids[0] = id0; S.cnt[0] = cnt0; }
                                    __syncwarp();
                                    sink.n_distinct = 1;
                                    spilled = true;
                                }
                                sink.add(S, vv, total, lane);
                            }
                            bal = __ballot_sync(FULL, todo != 0);
                            if(RUNS && first_round) {
                                // The tile's hits in k-mer order (lane-major). A hit starts a run when no hit precedes it in the record
                                // or its value differs from the previous hit's; every start closes the run before it. Most tiles hit
                                // ONE value -- this round of the counting loop has just found that out (no hit is left) and knows the
                                // value and the count: the tile continues the open run or starts a single new one.
                                first_round = false;
                                if(!bal) {
                                    const u32 lead = vv, H1 = total;
                                    if(run_val == lead) run_len += H1;
                                    else {
                                        if(run_val != VAL_MISS) {
                                            if(!run_direct && n_runs_rec + 2 > (u32)RUNBUF) {
                                                const u32 rest = (npos - p0) + ((first_mate && !last_mate && xl >= c) ? xl - c + 1 : 0u) + 1u;
                                                reserve_runs(n_runs_rec + rest);
                                                for(u32 q = lane; q < n_runs_rec; q += 32) if(blk_pos + q < runs_cap) runs_out[blk_pos + q] = s_run[q];
                                                run_direct = true;
                                            }
                                            if(lane == 0) put_run(n_runs_rec, run_val, run_len);
                                            ++n_runs_rec;
                                        }
                                        run_val = lead; run_len = H1;
                                    }
                                } else {
                                    const u32 todo = todo0, bal = bal0;                   // the tile's hits as they were before this round
                                    u32 firstv = VAL_MISS, prevv = VAL_MISS, inner = 0;
#pragma unroll
                                    for(int i = 0; i < PPL; ++i)
                                        if(todo >> i & 1u) {
                                            if(prevv == VAL_MISS) firstv = cand[i]; else if(cand[i] != prevv) inner |= 1u << i;
                                            prevv = cand[i];
                                        }
                                    const u32 lower = bal & ((1u << lane) - 1);                    // lanes before this one that hold hits
                                    u32 pv = __shfl_sync(FULL, prevv, (31 - __clz(lower)) & 31);   // value of the last hit before this lane
                                    if(lower == 0) pv = run_val;
                                    u32 starts = inner;
                                    if(todo && (pv == VAL_MISS || firstv != pv)) starts |= todo & (0u - todo);
                                    const u32 m = __popc(todo), ns = __popc(starts);
                                    u32 tot;
                                    const u32 ex = warp_excl_scan(m | (ns << 16), lane, tot);
                                    const u32 hbase = ex & 0xffffu, sbase = ex >> 16, H = tot & 0xffffu, SX = tot >> 16;
                                    const bool carry = run_val != VAL_MISS;
                                    // more runs than the record buffer holds (long reads): from here on straight into a reserved stretch
                                    if(!run_direct && n_runs_rec + SX + 1 > (u32)RUNBUF) {
                                        const u32 rest = (npos - p0) + ((first_mate && !last_mate && xl >= c) ? xl - c + 1 : 0u) + 1u;
                                        reserve_runs(n_runs_rec + rest);
                                        for(u32 q = lane; q < n_runs_rec; q += 32) if(blk_pos + q < runs_cap) runs_out[blk_pos + q] = s_run[q];
                                        run_direct = true;
                                    }
                                    u32 last_idx = 0, last_val = 0;                               // this lane's last start
                                    {
                                        u32 rank = 0;
#pragma unroll
                                        for(int i = 0; i < PPL; ++i)
                                            if(todo >> i & 1u) { if(starts >> i & 1u) { last_idx = hbase + rank; last_val = cand[i]; } ++rank; }
                                    }
                                    const u32 sl = __ballot_sync(FULL, ns != 0), slower = sl & ((1u << lane) - 1);
                                    int open_start = (int)__shfl_sync(FULL, last_idx, (31 - __clz(slower)) & 31);   // where the run open before this lane began
                                    if(slower == 0) open_start = -(int)run_len;
                                    u32 open_val = pv, rank = 0, srank = 0;
#pragma unroll
                                    for(int i = 0; i < PPL; ++i)
                                        if(todo >> i & 1u) {
                                            if(starts >> i & 1u) {
                                                const u32 t = sbase + srank;                      // the t-th start of the tile
                                                if(carry || t > 0) put_run(n_runs_rec + t - (carry ? 0u : 1u), open_val, (u32)((int)(hbase + rank) - open_start));
                                                open_val = cand[i]; open_start = (int)(hbase + rank); ++srank;
                                            }
                                            ++rank;
                                        }
                                    if(SX) {
                                        const u32 top = 31 - __clz(sl);
                                        run_val = __shfl_sync(FULL, last_val, top);
                                        run_len = H - __shfl_sync(FULL, last_idx, top);
                                        n_runs_rec += SX - (carry ? 0u : 1u);
                                    } else run_len += H;
                                }
                                __syncwarp();
                            }
                        } while(bal);
                    }
                }
            }
            if(first_mate && !last_mate) {                              // between the mates of a pair
                if(lane == (j >> msh)) my_m1 = n_emit;
                rb = xb; L = xl; tb ^= 1;
                continue;
            }
            if((MODE == LEAN_K || MODE == LEAN_R) && deferred) {                           // the generic kernel redoes this record from scratch
                if(lane == 0) defer_idx[atomicAdd(defer_cnt, 1ull)] = (u32)((r0 + j) >> msh);
                nd = 0; n_hit = 0; n_emit = 0;
                if(spilled) { spilled = false; sink.n_distinct = 0; sink.overflow = 0; }
            }
            // ---- resolve_tree (util.h:831-869) -------------------------------------------------------------------
            u32 taxon = 0;
            if(SV) taxon = resolve_lanes(cnt_lane, sink.vi, X, T.n_values, lane);
            else if(spilled) {
                if(sink.overflow) {
                    // more distinct taxa than the shared-memory lists hold (only a database of more than AGG_CAP values can do
                    // that): the second pass redoes the record with its lists in global memory
                    if(defer_idx) { if(lane == 0) defer_idx[atomicAdd(defer_cnt, 1ull)] = (u32)((r0 + j) >> msh); deferred = true; }
                    else if(lane == 0) atomicOr(status, 2u);
                } else taxon = sink.resolve(S, X, lane);
                sink.n_distinct = 0; sink.overflow = 0;
                __syncwarp();
            } else if(nd) taxon = sink.vi[id0].w;
            if(lane == (j >> msh)) { my_taxon = taxon; my_def = deferred; if(COUNTS) { my_hit = n_hit; my_miss = n_emit - n_hit; if(msh == 0) my_m1 = n_emit; } }
            if(RUNS) {
                // the run still open ends with the record; then the record's runs take their place in the warp's stretch
                if(!run_direct && n_runs_rec == 0) {
                    // at most that one run (most records): straight to its place
                    if(run_val != VAL_MISS) {
                        reserve_runs(1);
                        if(lane == 0 && blk_pos < runs_cap) runs_out[blk_pos] = ((u64)sink.vi[run_val].w << 32) | run_len;
                        n_runs_rec = 1;
                    }
                } else {
                    if(run_val != VAL_MISS) {
                        if(!run_direct && n_runs_rec + 1 > (u32)RUNBUF) {
                            reserve_runs(n_runs_rec + 1);
                            for(u32 q = lane; q < n_runs_rec; q += 32) if(blk_pos + q < runs_cap) runs_out[blk_pos + q] = s_run[q];
                            run_direct = true;
                        }
                        if(lane == 0) put_run(n_runs_rec, run_val, run_len);
                        ++n_runs_rec;
                        __syncwarp();
                    }
                    if(!run_direct) {
                        reserve_runs(n_runs_rec);
                        for(u32 q = lane; q < n_runs_rec; q += 32) if(blk_pos + q < runs_cap) runs_out[blk_pos + q] = s_run[q];
                        __syncwarp();
                    }
                }
                if(lane == (j >> msh)) { my_rpos = blk_pos; my_nruns = n_runs_rec; }
                blk_pos += n_runs_rec; blk_left -= n_runs_rec;
            }
            spilled = false;
            rb = xb; L = xl; tb ^= 1;
        }
        // ---- one coalesced store per output array for the batch ------------------------------------------------------
        const u32 nout = nrec >> msh;                                  // records of the batch
        const u64 o0 = r0 >> msh;
        if(lane < nout) {
            taxon_out[o0 + lane] = my_taxon;
            if(COUNTS) {
                if(nhit_out) nhit_out[o0 + lane] = my_hit;
                if(nmiss_out) nmiss_out[o0 + lane] = my_miss;
                if(mate1_out) mate1_out[o0 + lane] = my_m1;
            }
            if(RUNS) { run_pos_out[o0 + lane] = my_rpos; n_runs_out[o0 + lane] = my_nruns; }
        }
        const u32 cls = __popc(__ballot_sync(FULL, lane < nout && my_taxon != 0));
        const u32 ndef = __popc(__ballot_sync(FULL, lane < nout && my_def != 0));
        if(lane == 0) {                                                // classified_[2] (classifier.h:138,238), once per batch
            atomicAdd(&counters[0], (unsigned long long)cls);
            atomicAdd(&counters[1], (unsigned long long)(nout - cls - ndef));
        }
        pb ^= 1;
    }
}

}  // namespace bns
