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
constexpr int DISP_BITS = 4;     // slot displacement field; disp == 2^DISP_BITS - 1 is reserved for the empty slot
constexpr int MAX_DISP = (1 << DISP_BITS) - 2;

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

struct TableView {
    const u64 *slots;            // n_buckets * 4 u64
    u32 bucket_bits;             // b: bucket = h >> (64-b)
    u32 tag_shift;               // b - DISP_BITS: bits below it are {overflow flags, val}
    u32 val_mask;                // (1 << flag_shift) - 1
    u32 n_values;
    u32 flag_shift;              // tag_shift - F: the F overflow flags of slot 0 sit at [flag_shift, tag_shift)
    u32 flag_mask;               // F - 1 (F is a power of two): a key's flag is bit flag_shift + (low word of mix64 & flag_mask)
};

// Overflow flags. A key that found its home bucket full is stored in a later bucket and CLEARS, in slot 0 of the home
// bucket, the one of F flag bits its hash selects (an empty slot is all ones and so reads "nothing displaced"). A lookup
// that misses in the home bucket goes on to the next bucket only if ITS flag is cleared, so a miss costs one sector
// unless a displaced key shares both its home bucket and its flag. F is what the value field leaves free, at most 8.
__host__ __device__ inline u32 flag_count_for(u32 b, u32 n_values) {
    u32 vb = 1;
    while((1u << vb) < n_values) ++vb;
    const u32 avail = b - DISP_BITS > vb ? b - DISP_BITS - vb : 1u;
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

// the table's own hash: a bijection on u64 (xorshift and odd multiplies are invertible), so the low
// 64-b bits of h identify the key once the top b bits chose the bucket
__device__ __host__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
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
    x ^= x >> 32;
    x *= MI;
    x ^= x >> 32;
    return x;
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
