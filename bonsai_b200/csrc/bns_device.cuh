// bns_device.cuh -- device-side building blocks of the B200 classify path (sm_100a).
//
// What each piece replaces in the reference (dnbaker/bonsai @ 6741de9c, paths under the reference tree):
//   pack16 / tile staging      Encoder's per-base LUT walk            include/bonsai/encoder.h:240-272, alphabet.h:128
//   rc64 / canonical           reverse_complement / canonical_repr.   include/bonsai/kmerutil.h:83-90,137-140
//   lex_score                  FRev64                                  include/bonsai/encoder.h:47,59
//   table_probe                kh_get(c, ...) + kh_val                  include/bonsai/khash64.h:250-263
//   TaxAgg                     linear::counter<tax_t,u16>::add/count    linear/linear.h:229-244
//   resolve                    resolve_tree + lca                       include/bonsai/util.h:831-869,634-663
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bns {

typedef unsigned long long u64;
typedef uint32_t u32;

constexpr u32 FULL = 0xffffffffu;
constexpr u64 KMER_NONE = ~0ull;          // ENCODE_OVERFLOW, encoder.h:119
constexpr u32 VAL_MISS = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// Launch-constant parameters
// ---------------------------------------------------------------------------------------------
enum Family : u32 {
    FAM_U = 0,   // unspaced, unwindowed: rolling k-mers, N restarts              (encoder.h:240-272, 218-232)
    FAM_K = 1,   // kmer(pos)-based: canon-windowed (:211-217,622-628) and uncanon-spaced (:233-239,616-621)
    FAM_R = 2,   // rolling + window over the VALID k-mers, tail flush            (:273-306, :307-353)
    FAM_NONE = 3 // string overload with a spaced seed: emits nothing             (:437-440)
};
enum ScoreKind : u32 { SC_LEX = 0, SC_ENT_ROLL = 1, SC_ENT_NOTFULL = 2 };

constexpr int MAX_SEG = 32;
constexpr int TILE = 128;        // k-mer positions per warp tile (4 per lane, lane-contiguous)
constexpr int PPL = 4;           // positions per lane
constexpr int CMAX = 256;        // largest comb (spaced span) supported
constexpr int AGG_CAP = 256;     // distinct taxa tracked per record in shared memory (covers any read up to k+255 bases)
constexpr int DISP_BITS = 4;     // slot displacement field of LAYOUT_HASH; disp == 2^bits - 1 is reserved for the empty slot
constexpr int DISP_BITS_LOC = 3; // LAYOUT_MINIMIZER: probes are 64-byte units; chains of up to 6 units, then the stash

struct EncParams {
    u32 k, c, W;                 // W = w_ - c_ + 1  (QueueMap size, encoder.h:142)
    u32 family, score_kind, cast_wrap;
    u32 canon_elem;              // canonicalise the element before scoring (FAM_U canon, FAM_K a8)
    u32 canon_emit;              // canonicalise only on emit (entropy canon wrapper, encoder.h:347-353)
    u32 filter_none;             // drop window results equal to ~0 (FAM_K)
    u32 tail_flush;              // FAM_R
    u32 t_restart;               // FAM_R + Lex, k >= 31: 32 consecutive T restart the rolling state (encoder.h:283)
    u32 n_seg;                   // contiguous runs of the comb
    uint16_t seg_off[MAX_SEG];   // base offset of each run from the k-mer start
    uint16_t seg_len[MAX_SEG];   // bases in the run
    double plogp[33];            // n/k * log(n/k) evaluated by the HOST libm (entropy.h:44-48)
};

// How a key becomes (home bucket, remainder). Both are bijections key <-> (bucket, remainder), so hits and misses are exact.
//   LAYOUT_HASH      bucket = top b bits of mix64(key), remainder = the other 64-b bits. Consecutive k-mers of a read land
//                    in unrelated DRAM lines: one line per lookup once the table exceeds L2. A probe is one 32-byte bucket.
//   LAYOUT_MINIMIZER (23 <= k <= 31) a probe is a 64-byte UNIT of two adjacent buckets (8 slots). The home unit lies in a
//                    GROUP of four units (256 bytes, two 128-byte lines) chosen by the k-mer's canonical 16-mer minimizer, the
//                    unit inside the group by the minimizer's position mod 4; the remainder spells the k-mer relative to
//                    its minimizer (loc_pack). ~(k-14)/2 consecutive k-mers of a read share a minimizer and so a group:
//                    the lines one lookup brings into L2 serve the next ones (~30 lines per 150 bp read instead of 120).
//                    (A variant that stored a unit as a sector of eight high words followed by a sector of eight low words --
//                    one LDG.256 to find the candidate, one LDG.32 for its low word, 8F overflow flags per unit -- measured
//                    10 % slower on the 2^28-key stress table: profiles/ncu_r02_loc3_soa.txt.)
enum TableLayout : u32 { LAYOUT_HASH = 0, LAYOUT_MINIMIZER = 1 };
constexpr u32 LOC_L = 16;        // minimizer length: a k-mer holds k-15 <= 16 of them for k <= 31, so a full run of consecutive
                                 // k-mers sharing one minimizer puts at most 4 keys into each unit of its group
constexpr u32 LOC_MB = 2 * LOC_L; // bits of a minimizer
constexpr u32 LOC_GB = 3;         // a group is 2^3 buckets = four 64-byte units

struct TableFmt {
    u32 b;                       // bucket bits: 2^b buckets of 32 bytes
    u32 fmt_bits;                // 64 - (bits of the remainder): b for LAYOUT_HASH, loc_fmt_bits(k, b) for LAYOUT_MINIMIZER
    u32 F;                       // overflow flags per bucket (power of two)
    u32 layout;
    u32 kt;                      // k of the keys (LAYOUT_MINIMIZER spells keys as k-mers)
    u32 disp_bits;               // width of the displacement field
    __host__ __device__ u32 tag_shift() const { return fmt_bits - disp_bits; }
    __host__ __device__ u32 flag_shift() const { return fmt_bits - disp_bits - F; }
    __host__ __device__ u32 max_disp() const { return (1u << disp_bits) - 2u; }
    __host__ __device__ u32 unit_slots() const { return layout == LAYOUT_MINIMIZER ? 8u : 4u; }   // slots one probe covers
};

struct TableView {
    const u64 *slots;            // n_buckets * 4 u64
    u32 bucket_bits;             // b: 2^b buckets
    u32 tag_shift;               // fmt_bits - disp_bits: bits below it are {overflow flags, val}
    u32 val_mask;                // (1 << flag_shift) - 1
    u32 n_values;
    u32 flag_shift;              // tag_shift - F: the F overflow flags of slot 0 sit at [flag_shift, tag_shift)
    u32 flag_mask;               // F - 1 (F is a power of two): a key's flag is bit flag_shift + (flag selector & flag_mask)
    TableFmt fmt;
    // LAYOUT_MINIMIZER only: the keys whose overflow chain was full to its end (repeated minimizers of real genomes crowd
    // single groups) live in a small LAYOUT_HASH table of the same value ids, probed after an exhausted chain
    const u64 *stash;
    TableFmt sfmt;
};
// the stash as a table of its own
__host__ __device__ inline TableView stash_view(const TableView &T) {
    TableView S;
    S.slots = T.stash; S.bucket_bits = T.sfmt.b; S.fmt = T.sfmt;
    S.tag_shift = T.sfmt.tag_shift(); S.flag_shift = T.sfmt.flag_shift(); S.flag_mask = T.sfmt.F - 1;
    S.val_mask = (1u << S.flag_shift) - 1; S.n_values = T.n_values;
    S.stash = nullptr; S.sfmt = T.sfmt;
    return S;
}

// Overflow flags. A key that found its home bucket full is stored in a later bucket and CLEARS, in slot 0 of the home
// bucket, the one of F flag bits its hash selects (an empty slot is all ones and so reads "nothing displaced"). A lookup
// that misses in the home bucket goes on to the next bucket only if ITS flag is cleared, so a miss costs one sector
// unless a displaced key shares both its home bucket and its flag. F is what the value field leaves free, at most 8.
__host__ __device__ inline u32 flag_count_for(u32 fmt_bits, u32 disp_bits, u32 n_values) {
    u32 vb = 1;
    while((1u << vb) < n_values) ++vb;
    const u32 avail = fmt_bits - disp_bits > vb ? fmt_bits - disp_bits - vb : 1u;
    return avail >= 8 ? 8u : avail >= 4 ? 4u : avail >= 2 ? 2u : 1u;
}

struct TaxView {
    // one 16-byte record per distinct DB value: {tin, tout, node, taxid}
    const uint4 *val_info;
    // one per taxonomy node (index 0 = "no node"): {tin, tout, parent_node, taxid}
    const uint4 *node_info;
    u32 n_nodes;
    u32 node_of_one;             // node index of taxid 1 (lca's fall-through result, util.h:662)
};

// ---------------------------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 lex_score(u64 x) {
    x ^= 0x533f8c2151b20f97ull;
    x *= 0x9a98567ed20c127dull;
    x = (x << 31) | (x >> 33);
    return x ^ 0x691a9d706391077aull;
}

// reverse complement of a 2-bit packed k-mer: reverse all 64 bits, swap the two bits of every pair back,
// complement, shift the k groups down
__device__ __forceinline__ u64 rc64(u64 x, u32 k) {
    u64 r = __brevll(x);
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
    return (~r) >> (64 - 2 * k);
}
__device__ __forceinline__ u64 canonical(u64 x, u32 k) {
    const u64 r = rc64(x, k);
    return x < r ? x : r;
}

// the table's own hash: a bijection on u64 (an odd multiply and an xorshift are invertible), so the low 64-b bits of h
// identify the key once the top b bits chose the bucket. One multiply is enough: the top bits of x * M depend on every bit
// of x (multiplicative hashing), and on the 1.1 M- and 10.5 M-key databases of the 4-genome set it displaces exactly as
// many keys as a two-round mixer did (0.59 % / 0.91 % at 1.1 / 1.25 keys per bucket) with no upper-word clash, for half the
// instructions; the xorshift carries the well-mixed high word into the low word the overflow-flag selector uses.
__device__ __host__ __forceinline__ u64 mix64(u64 x) {
    x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
    return x;
}

// modular inverse of the odd multiplier (Newton iteration), for unmix64
__host__ __device__ constexpr u64 inv_odd(u64 a) {
    u64 x = a;                       // correct to 3 bits
    for(int i = 0; i < 6; ++i) x *= 2 - a * x;
    return x;
}
__device__ __host__ __forceinline__ u64 unmix64(u64 x) {
    constexpr u64 MI = inv_odd(0xd6e8feb86659fd93ull);
    x ^= x >> 32;
    x *= MI;
    return x;
}

// ---------------------------------------------------------------------------------------------
// LAYOUT_MINIMIZER: key <-> (bucket, remainder)
// ---------------------------------------------------------------------------------------------
// bijections on n-bit values (odd multiplies and xorshifts by >= n/2 are invertible mod 2^n)
__host__ __device__ __forceinline__ u32 nmix(u32 x, u32 n) {
    const u32 m = n >= 32 ? ~0u : ((1u << n) - 1), h = (n + 1) >> 1;
    x = (x * 0x9e3779b1u) & m; x ^= x >> h;
    x = (x * 0x85ebca6bu) & m; x ^= x >> h;
    return x;
}
__host__ __device__ __forceinline__ u32 nunmix(u32 x, u32 n) {
    const u32 m = n >= 32 ? ~0u : ((1u << n) - 1), h = (n + 1) >> 1;
    x ^= x >> h; x = (x * (u32)inv_odd(0x85ebca6bu)) & m;
    x ^= x >> h; x = (x * (u32)inv_odd(0x9e3779b1u)) & m;
    return x;
}
// reverse complement of a 16-mer held in 32 bits
__host__ __device__ __forceinline__ u32 rc16(u32 x) {
#ifdef __CUDA_ARCH__
    u32 r = __brev(x);
#else
    u32 r = x;
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    r = ((r >> 2) & 0x33333333u) | ((r & 0x33333333u) << 2);
    r = ((r >> 4) & 0x0f0f0f0fu) | ((r & 0x0f0f0f0fu) << 4);
    r = ((r >> 8) & 0x00ff00ffu) | ((r & 0x00ff00ffu) << 8);
    r = (r >> 16) | (r << 16);
#endif
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);     // bit-reversed pairs back in order
    return ~r;
}
// The minimizer's two hashes, both bijections on 32 bits and one instruction pair each:
//   loc_hash   orders the canonical 16-mers of a k-mer (its top 27 bits; the xor keeps poly-A from being everybody's minimum);
//   loc_place  folds the well-mixed high half into the low half: the GROUP is the low bits of loc_place(loc_hash(c)) -- the
//              minimum of 16 hashes is small in its top bits, which would crowd the keys into few groups -- and the bits above
//              them go into the remainder.
__host__ __device__ __forceinline__ u32 loc_hash(u32 c) { return (c ^ 0x5bd1e995u) * 0x9e3779b1u; }
__host__ __device__ __forceinline__ u32 loc_unhash(u32 h) { return (h * (u32)inv_odd(0x9e3779b1u)) ^ 0x5bd1e995u; }
__host__ __device__ __forceinline__ u32 loc_place(u32 h) { return h ^ (h >> 16); }                 // its own inverse
// remainder bits of a LAYOUT_MINIMIZER slot: [the other k-16 bases | position >> 2 | orientation | group hash above the group bits]
__host__ __device__ __forceinline__ u32 loc_rembits(u32 k, u32 b) { return 2 * (k - LOC_L) + 2 + 1 + (LOC_MB - (b - LOC_GB)); }
__host__ __device__ __forceinline__ u32 loc_fmt_bits(u32 k, u32 b) { return 64 - loc_rembits(k, b); }
struct TableHash { u64 home; u64 tag; u32 fsel; };                // home bucket, left-aligned remainder, overflow-flag selector

// (bucket, remainder) of the 2k-bit k-mer s once its minimizer is known: h = loc_hash(canonical 16-mer), at position bp of s
// (0 = the first, most significant base), bo = the 16-mer stands in s as the reverse complement of its canonical form.
// The other k-16 bases are taken as s rotated left by bp bases with the minimizer dropped: (right part : left part) -- two
// shifts on 32-bit halves in the classify kernel.
__host__ __device__ __forceinline__ TableHash loc_pack(u64 s, u32 k, u32 b, u32 h, u32 bp, u32 bo) {
    const u32 nf = 2 * (k - LOC_L), bl = b - LOC_GB;
    const u64 fmask = (1ull << nf) - 1;
    const u64 flank = ((s << (2 * bp)) & fmask) | (bp ? (s >> (2 * (k - bp))) : 0ull);
    const u32 p = loc_place(h);
    const u32 line = bl >= LOC_MB ? p : (p & ((1u << bl) - 1)), mrest = bl >= LOC_MB ? 0u : (p >> bl);
    const u64 rem = ((((flank << 2) | (bp >> 2)) << 1 | bo) << (LOC_MB - bl)) | mrest;
    TableHash t;
    t.home = ((u64)line << LOC_GB) | ((bp & 3u) << 1);            // the unit's first (even) bucket
    t.tag = rem << (64 - loc_rembits(k, b));
    t.fsel = (u32)(t.tag >> 34);                                   // k = 31: the other 15 bases
    return t;
}
// the order of the minimizer: the top 27 bits of loc_hash of the canonical 16-mer (value and position sort as one 32-bit
// word in the classify kernel), LEFTMOST position of the k-mer on ties
__host__ __device__ inline TableHash loc_encode(u64 s, u32 k, u32 b) {
    const u32 J = k - LOC_L + 1;
    u32 best = ~0u, best27 = ~0u, bp = 0, bo = 0;
    for(u32 p = 0; p < J; ++p) {
        const u32 f = (u32)(s >> (2 * (k - LOC_L - p)));                     // position 0 = the first (most significant) base
        const u32 r = rc16(f), c = f < r ? f : r, mh = loc_hash(c);
        if((mh >> 5) < best27) { best27 = mh >> 5; best = mh; bp = p; bo = r < f; }
    }
    return loc_pack(s, k, b, best, bp, bo);
}
// inverse: home bucket (any bucket of the unit) + left-aligned remainder -> key
__host__ __device__ inline u64 loc_decode(u64 home, u64 tag, u32 k, u32 b) {
    const u32 bl = b - LOC_GB;
    const u64 rem = tag >> (64 - loc_rembits(k, b));
    const u32 mrest = bl >= LOC_MB ? 0u : (u32)(rem & ((1ull << (LOC_MB - bl)) - 1));
    const u64 up = rem >> (LOC_MB - bl);
    const u32 bo = (u32)up & 1u, bp = (((u32)(up >> 1) & 3u) << 2) | (((u32)home >> 1) & 3u);
    const u64 flank = up >> 3;
    const u32 p = bl >= LOC_MB ? (u32)(home >> LOC_GB) : ((mrest << bl) | (u32)(home >> LOC_GB));
    const u32 c = loc_unhash(loc_place(p)), m = bo ? rc16(c) : c;
    const u32 nr = 2 * (k - LOC_L - bp);                                     // bits right of the minimizer
    const u64 left = bp ? (flank & ((1ull << (2 * bp)) - 1)) : 0ull, right = flank >> (2 * bp);
    return (bp ? (left << (2 * (k - bp))) : 0ull) | ((u64)m << nr) | right;
}
// the d-th probe of a key, as the index of its first bucket. LAYOUT_HASH: the next buckets. LAYOUT_MINIMIZER: the home unit,
// then a run of units starting at a scrambled image of the home unit: keys that overflow a crowded group scatter instead of
// piling into the next group, so probe sequences stay short.
__host__ __device__ __forceinline__ u64 probe_bucket(u32 layout, u64 home, u32 d, u32 b) {
    const u64 bmask = b >= 64 ? ~0ull : ((1ull << b) - 1);
    if(layout == LAYOUT_MINIMIZER) {
        if(d == 0) return home & ~1ull;
        return (((u64)nmix((u32)(home >> 1), b - 1) + (d - 1)) << 1) & bmask;
    }
    return (home + d) & bmask;
}
// ... and back: the home bucket of an entry found in `bucket` with displacement d
__host__ __device__ __forceinline__ u64 probe_home(u32 layout, u64 bucket, u32 d, u32 b) {
    const u64 bmask = b >= 64 ? ~0ull : ((1ull << b) - 1);
    if(layout == LAYOUT_MINIMIZER) {
        if(d == 0) return bucket & ~1ull;
        return (u64)nunmix((u32)((((bucket >> 1) - (d - 1)) & (bmask >> 1))), b - 1) << 1;
    }
    return (bucket - d) & bmask;
}
// key -> (home bucket, remainder) of either layout. Keys that are not k-mers of the table's k cannot be in a
// LAYOUT_MINIMIZER table: `possible` is false and the caller reports a miss.
__host__ __device__ inline TableHash table_hash(const TableFmt &f, u64 key, bool &possible) {
    possible = true;
    if(f.layout == LAYOUT_MINIMIZER) {
        if(f.kt < 32 && (key >> (2 * f.kt)) != 0) { possible = false; key &= (1ull << (2 * f.kt)) - 1; }
        return loc_encode(key, f.kt, f.b);
    }
    const u64 h = mix64(key);
    TableHash t;
    t.home = h >> (64 - f.b); t.tag = h << f.b; t.fsel = (u32)h;
    return t;
}

__device__ __forceinline__ void ld_bucket(const u64 *p, u64 &a, u64 &b, u64 &c, u64 &d) {
    // one 32-byte sector per probe (LDG.E.256)
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

// (u64)double as the two x86-64 code paths of the reference execute it
__device__ __forceinline__ u64 cast_u64(double x, u32 wrap) {
    if(!wrap) {                                   // AVX-512 vcvttsd2usi
        if(!(x > -1.0) || x >= 18446744073709551616.0) return ~0ull;   // also NaN
        if(x < 0.0) return 0;
        return __double2ull_rz(x);
    }
    if(x != x) return 0x8000000000000000ull;
    if(x < 9223372036854775808.0) {
        if(!(x > -9223372036854775809.0)) return 0x8000000000000000ull;
        return (u64)__double2ll_rz(x);
    }
    x -= 9223372036854775808.0;
    if(!(x < 9223372036854775808.0)) return 0ull;        // 0x8000.. ^ 0x8000..
    return ((u64)__double2ll_rz(x)) ^ 0x8000000000000000ull;
}

// ---------------------------------------------------------------------------------------------
// 16 ASCII bases -> one u32 of 2-bit codes (first base in the top bits) + 16 "invalid" bits (first
// base in bit 15). Codes: A/a 0, C/c 1, G/g 2, T/t 3; everything else invalid (alphabet.h:128).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack4(u32 w, u32 &code8, u32 &bad4, u32 &t4) {
    u32 t = (w >> 1) & 0x03030303u;
    t ^= (t >> 1) & 0x01010101u;                       // per byte: A0 C1 G2 T3
    code8 = (t * 0x40100401u) >> 24;                   // b0<<6 | b1<<4 | b2<<2 | b3
    const u32 hi = (t >> 1) & 0x01010101u, lo = t & 0x01010101u;
    const u32 expect = 0x41414141u + 2u * t + 2u * hi + 11u * (hi & lo);   // 'A','C','G','T'
    const u32 diff = expect ^ (w & 0xdfdfdfdfu);
    const u32 nz = (((diff & 0x7f7f7f7fu) + 0x7f7f7f7fu) | diff) & 0x80808080u;
    bad4 = (((nz >> 7) * 0x08040201u) >> 24) & 0xfu;   // byte0 -> bit 3
    t4 = ((((hi & lo) * 0x08040201u) >> 24) & 0xfu) & ~bad4;   // valid T
}
// fast path: codes only, plus a word that is non-zero iff some byte is not one of ACGTacgt
__device__ __forceinline__ u32 pack4_fast(u32 w, u32 &code8) {
    u32 t = (w >> 1) & 0x03030303u;
    t ^= (t >> 1) & 0x01010101u;
    code8 = (t * 0x40100401u) >> 24;
    const u32 hi = (t >> 1) & 0x01010101u, lo = t & 0x01010101u;
    const u32 expect = 0x41414141u + 2u * t + 2u * hi + 11u * (hi & lo);
    return expect ^ (w & 0xdfdfdfdfu);
}
__device__ __forceinline__ u32 pack16_fast(uint4 v, u32 &codes) {
    u32 c0, c1, c2, c3;
    const u32 d = pack4_fast(v.x, c0) | pack4_fast(v.y, c1) | pack4_fast(v.z, c2) | pack4_fast(v.w, c3);
    codes = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    return d;
}
__device__ __forceinline__ void pack16(uint4 v, u32 &codes, u32 &bad, u32 &tmask) {
    u32 c0, c1, c2, c3, b0, b1, b2, b3, t0, t1, t2, t3;
    pack4(v.x, c0, b0, t0); pack4(v.y, c1, b1, t1); pack4(v.z, c2, b2, t2); pack4(v.w, c3, b3, t3);
    codes = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    bad = (b0 << 12) | (b1 << 8) | (b2 << 4) | b3;
    tmask = (t0 << 12) | (t1 << 8) | (t2 << 4) | t3;
}

// n (1..32) bases starting at base coordinate q of a big-endian 2-bit word array in shared memory
__device__ __forceinline__ u64 extract_bases(const u32 *codes, u32 q, u32 n) {
    const u32 wi = q >> 4, s = (q & 15u) * 2u;
    const u32 w0 = codes[wi], w1 = codes[wi + 1], w2 = codes[wi + 2];
    const u32 hi = __funnelshift_l(w1, w0, s), lo = __funnelshift_l(w2, w1, s);
    const u64 v = ((u64)hi << 32) | lo;
    return v >> (64 - 2 * n);
}
// n (1..32) invalid-bits starting at coordinate q; bad words hold 16 bases each in their LOW 16 bits
__device__ __forceinline__ u32 any_bad(const u32 *bad, u32 q, u32 n) {
    const u32 wi = q >> 4, s = q & 15u;
    const u64 v = ((u64)(bad[wi] & 0xffffu) << 32) | ((u64)(bad[wi + 1] & 0xffffu) << 16) | (u64)(bad[wi + 2] & 0xffffu);
    // 48 bits, coordinate wi*16 at bit 47
    const u64 win = (v << (16 + s)) >> (64 - n);
    return win != 0;
}

}  // namespace bns
