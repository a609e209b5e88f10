// bns_classify_loc.cuh -- what `bonsai classify` runs (every canonical 31-mer, no window; bin/bonsai.cpp:152) against a
// LAYOUT_MINIMIZER table, i.e. a table far beyond L2 (BASELINE configs[4]), as a two-stage pipeline inside each warp.
// Included by bns_kernels.cu after bns_classify_u.cuh, whose building blocks it uses; same algorithm, same table, same results
// as bns_classify_u_kernel<LEAN_U, true, 31, COUNTS, 0, true, false>.
//
// Why a second kernel. ncu on the lean kernel against the 34 GB table (profiles/ncu_r02_final_stress.txt): 4.0 of the 10.4
// warp-cycles per issued instruction are spent waiting for the home units to arrive from DRAM, because a warp computes a tile's
// minimizers, asks for the units and has nothing else to do until they are there; the same kernel against a table that lives
// in L2 is 24 % faster. Here a warp works on two tiles at once:
//   stage A (tile t+1)  bases -> 2-bit window -> canonical k-mers -> minimizers -> (home unit, remainder) of the lane's four
//                       keys, written to shared memory, and ONE L2 prefetch per key;
//   stage B (tile t)    reads its keys back, loads the home units (by now in L2 or on their way: stage A of the next tile and
//                       stage B of the previous one ran in between), matches, follows overflow chains, counts the hits.
// The distinct-taxon list of a record is not a list either: with a value dictionary of at most 32 entries (what the kernel is
// used for) lane v keeps the count of value id v in a register; linear::counter::add (linear.h:229) is one predicated add and
// resolve_tree (util.h:831-869) reads the counts by shuffle. The tied maxima are folded with lca in value-id order instead of
// first-seen order -- lca is commutative and associative on the loader's well-formed taxonomy, so the taxon is the same.
#pragma once

namespace bns {

constexpr int LOCP_OFF_BYTES = 2 * (RB + 2) * 8;                       // offsets of the current / next batch
constexpr int LOCP_RING = 3;                                         // first-tile staging buffers: the tile is requested two records ahead
constexpr int LOCP_RD_BYTES = LOCP_RING * 32 * 8;
constexpr int LOCP_STATE_BYTES = 2 * (3 * 32 * 16 + 32 * 4);         // two tiles' keys: home unit / high word / low word of four keys per lane + live mask
constexpr int LOCP_Q_BYTES = TILE * 16;                              // keys that follow their overflow chain, compacted over the warp
constexpr int LOCP_WARP_BYTES = LOCP_OFF_BYTES + LOCP_RD_BYTES + LOCP_STATE_BYTES + LOCP_Q_BYTES;
constexpr u32 LOCP_MAX_VALUES = 32;                                  // one lane per value id

template <bool COUNTS>
__global__ void __launch_bounds__(LEAN_WARPS * 32, BNS_CLASSIFY_U_MIN_CTAS)
bns_classify_loc_kernel(const char *__restrict__ bases, const u64 *__restrict__ offsets, u64 n_records, TableView T, TaxView X,
                        u32 *__restrict__ taxon_out, u32 *__restrict__ nhit_out, u32 *__restrict__ nmiss_out,
                        unsigned long long *__restrict__ counters, u32 *__restrict__ status, u32 fixed_len, u64 fixed_base) {
    __shared__ __align__(16) uint4 s_vi[LOCP_MAX_VALUES];
    __shared__ __align__(8) unsigned long long s_mbar;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    constexpr u32 k = 31, span = TILE + k - 1;
    tma_stage_val_info(s_vi, X.val_info, T.n_values * (u32)sizeof(uint4), &s_mbar);
    const uint4 *vi = s_vi;

    ProbeConst Pc;
    Pc.slots = (const char *)T.slots;
    Pc.b = T.bucket_bits;
    Pc.idx_shift = 32 - T.bucket_bits;
    Pc.hm = ~0u << T.tag_shift;
    Pc.tag_shift = T.tag_shift; Pc.fmt_bits = T.fmt.fmt_bits; Pc.max_disp = T.fmt.max_disp(); Pc.layout = T.fmt.layout;
    Pc.flags_all = ((1u << T.tag_shift) - 1) & ~T.val_mask;
    Pc.flag_shift = T.flag_shift; Pc.flag_mask = T.flag_mask;
    Pc.val_mask = T.val_mask;
    const u32 kmask_lo = 0xffffffffu, kmask_hi = 0x3fffffffu;

    unsigned char *wbase = g_smem + (size_t)wid * LOCP_WARP_BYTES;
    u64 *s_off = (u64 *)wbase;
    uint2 *s_rd = (uint2 *)(wbase + LOCP_OFF_BYTES);
    uint4 *s_st = (uint4 *)(wbase + LOCP_OFF_BYTES + LOCP_RD_BYTES);              // [2][3][32]
    u32 *s_msk = (u32 *)(s_st + 2 * 3 * 32);                                      // [2][32]
    uint4 *s_q = (uint4 *)(wbase + LOCP_OFF_BYTES + LOCP_RD_BYTES + LOCP_STATE_BYTES);

    const u64 nwarps = (u64)gridDim.x * LEAN_WARPS;
    const u64 n_batches = (n_records + RB - 1) / RB;
    u64 bt = (u64)blockIdx.x * LEAN_WARPS + wid;
    auto async8 = [](void *dst, const void *src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((u32)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    auto async_commit = [] { asm volatile("cp.async.commit_group;" ::: "memory"); };
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
    auto fetch_tile = [&](u64 rb, u32 rl, u32 buf) {                    // bases [rb, rb + min(rl, span)) -> s_rd[buf]
        const char *a0 = bases + rb;
        const u32 shift = (u32)((uintptr_t)a0 & 7u);
        const u32 nblk = rl ? (shift + min(rl, span) + 7) >> 3 : 0u;
        if(lane < nblk) async8(s_rd + buf * 32 + lane, a0 - shift + 8 * lane);
        else s_rd[buf * 32 + lane] = make_uint2(0x41414141u, 0x41414141u);
    };
    auto tile_block = [&](u64 rb, u32 rl) -> uint2 {                    // later tiles of a long record: plain loads
        const char *a0 = bases + rb;
        const u32 shift = (u32)((uintptr_t)a0 & 7u);
        const u32 nblk = rl ? (shift + min(rl, span) + 7) >> 3 : 0u;
        uint2 v = make_uint2(0x41414141u, 0x41414141u);
        if(lane < nblk) v = __ldg(reinterpret_cast<const uint2 *>(a0 - shift) + lane);
        return v;
    };
    auto rec_len = [](u64 b, u64 e) -> u32 { const u64 len = e - b; return len < 0xffffffffull ? (u32)len : 0xffffffffu; };

    // ---- stage B state of the record being counted -------------------------------------------------------------------
    u32 cnt_lane = 0;                                                  // hits of value id `lane` (linear::counter, one entry per lane)
    u32 n_hit = 0;
    u32 my_taxon = 0, my_hit = 0, my_miss = 0;                         // results of the batch's record `lane`
    // ---- the tile between the stages (warp-uniform) -------------------------------------------------------------------
    bool pending = false, pend_first = false, pend_last = false;
    u32 pend_slot = 0, pend_j = 0, pend_emit = 0;

    // stage B: probe, count, and at the record's last tile resolve
    auto stage_b = [&]() {
        const u32 slot = pend_slot;
        if(pend_first) { cnt_lane = 0; n_hit = 0; }
        const uint4 h4 = s_st[(slot * 3 + 0) * 32 + lane], t4 = s_st[(slot * 3 + 1) * 32 + lane], l4 = s_st[(slot * 3 + 2) * 32 + lane];
        const u32 mask = s_msk[slot * 32 + lane];
        const u32 hb[PPL] = {h4.x, h4.y, h4.z, h4.w}, th[PPL] = {t4.x, t4.y, t4.z, t4.w}, tl[PPL] = {l4.x, l4.y, l4.z, l4.w};
        u32 cand[PPL], nok = 0, more = 0;
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
                u32 c = ~tl[i];
#pragma unroll
                for(int sl = 7; sl >= 0; --sl) c = w[j][2 * sl + 1] == th[i] ? w[j][2 * sl] : c;
                if(((c ^ tl[i]) & Pc.hm) != 0 && c != ~tl[i]) {         // two entries of the unit share a high word (2^-30): exact pass
                    c = ~tl[i];
#pragma unroll
                    for(int sl = 7; sl >= 0; --sl)
                        if(w[j][2 * sl + 1] == th[i] && ((w[j][2 * sl] ^ tl[i]) & Pc.hm) == 0) c = w[j][2 * sl];
                }
                cand[i] = c;
                const u32 miss = min((c ^ tl[i]) & Pc.hm, 1u);
                nok |= miss << i;
                more |= (miss & ~(w[j][0] >> (Pc.flag_shift + ((th[i] >> 2) & Pc.flag_mask)))) << i;
            }
        }
        more &= mask;
        if(__any_sync(FULL, more != 0)) {                              // overflow chains: compacted over the warp, one key per lane
            u32 tot;
            const u32 at0 = warp_excl_scan(__popc(more), lane, tot);
            u32 at = at0;
#pragma unroll
            for(int i = 0; i < PPL; ++i) if(more >> i & 1u) s_q[at++] = make_uint4(hb[i], th[i], tl[i], 0u);
            __syncwarp();
            for(u32 base = 0; base < tot; base += 32) {
                const u32 t = base + lane;
                if(t < tot) {
                    const uint4 e = s_q[t];
                    bool exhausted;
                    u32 c = probe_chain_loc(Pc, e.x, e.y, e.z, exhausted);
                    if(exhausted && T.stash) {
                        const u32 v = probe_stash(T, loc_decode(e.x, ((u64)e.y << 32) | e.z, k, Pc.b));
                        if(v != VAL_MISS) c = (e.z & Pc.hm) | v;
                    }
                    s_q[t].w = c;
                }
            }
            __syncwarp();
            at = at0;
#pragma unroll
            for(int i = 0; i < PPL; ++i)
                if(more >> i & 1u) {
                    const u32 c = s_q[at++].w;
                    if(c != ~tl[i]) { nok &= ~(1u << i); cand[i] = c; }
                }
            __syncwarp();
        }
        u32 todo = ~nok & mask;
        u32 bal = __ballot_sync(FULL, todo != 0);
        if(bal) {
            if(COUNTS) n_hit += __reduce_add_sync(FULL, __popc(todo));
#pragma unroll
            for(int i = 0; i < PPL; ++i) cand[i] &= Pc.val_mask;
            do {                                                       // one distinct value of the tile per round
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
                if(lane == vv) cnt_lane += total;
                bal = __ballot_sync(FULL, todo != 0);
            } while(bal);
        }
        if(pend_last) {
            // ---- resolve_tree over the lanes that hold a count ---------------------------------------------------------
            const u32 present = __ballot_sync(FULL, cnt_lane != 0);
            u32 taxon = 0;
            if(present & (present - 1)) {
                const u32 ti = vi[min(lane, T.n_values - 1)].x;
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
                    else {                                             // lca(node, l): climb until the interval covers l (util.h:634-663)
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
                taxon = ntied == 1 ? vi[first_id].w : X.node_info[node].w;
            } else if(present) taxon = vi[__ffs(present) - 1].w;
            if(lane == pend_j) { my_taxon = taxon; if(COUNTS) { my_hit = n_hit; my_miss = pend_emit - n_hit; } }
        }
        pending = false;
    };

    u32 pb = 0;                                                        // offsets buffer of the current batch
    fetch_offsets(bt, 0);
    async_commit();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    u32 tb = 0, slot_a = 0;                                            // ring index of the current record's first tile; stage A's state slot
    u64 rb = 0, nb_ = 0;                                               // the current record and the one after it
    u32 L = 0, nl_ = 0;
    if(bt < n_batches) {
        const u32 nrec0 = (u32)min((u64)RB, n_records - bt * RB);
        rb = s_off[0]; L = rec_len(rb, s_off[1]);
        if(nrec0 > 1) { nb_ = s_off[1]; nl_ = rec_len(nb_, s_off[2]); }
    }
    fetch_tile(rb, L, 0);
    async_commit();
    fetch_tile(nb_, nl_, 1);
    async_commit();
    mbar_wait(&s_mbar, 0);

    for(; bt < n_batches; bt += nwarps) {
        fetch_offsets(bt + nwarps, pb ^ 1);
        async_commit();
        const u64 r0 = bt * RB;
        const u32 nrec = (u32)min((u64)RB, n_records - r0);
        const bool have_next_batch = bt + nwarps < n_batches;
        const u32 nrec_next = have_next_batch ? (u32)min((u64)RB, n_records - (bt + nwarps) * RB) : 0u;
        my_taxon = 0; my_hit = 0; my_miss = 0;
        // j == nrec is the flush step: no record, stage B takes the batch's last tile so that its results leave with the batch
        for(u32 j = 0; j <= nrec; ++j) {
            u32 npos = 0, emit = 0;
            uint2 pre = make_uint2(0u, 0u);
            u64 zb = 0; u32 zl = 0;
            if(j < nrec) {
                // pending copies, newest first: [next batch's offsets at j == 0,] the tile of record j+1, the tile of record j
                if(j == 0) asm volatile("cp.async.wait_group 2;" ::: "memory"); else asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncwarp();
                pre = s_rd[tb * 32 + lane];
                // the record after the next: its first tile is requested now
                const u32 idx = j + 2;
                const u64 *o = nullptr;
                if(idx < nrec) o = s_off + pb * (RB + 2) + idx;
                else if(idx - nrec < nrec_next) o = s_off + (pb ^ 1) * (RB + 2) + (idx - nrec);
                if(o) { zb = o[0]; zl = rec_len(zb, o[1]); }
                fetch_tile(zb, zl, (tb + 2) % LOCP_RING);
                async_commit();
                if(L == 0xffffffffu) { if(lane == 0) atomicOr(status, 8u); }
                else if(L >= k) npos = L - k + 1;
            }
            // a record without a k-mer has no tile (its results stay 0); the flush step has one empty iteration
            const u32 p_end = j < nrec ? npos : 1u;
            for(u32 p0 = 0; p0 < p_end; p0 += TILE) {
                if(npos) {
                    // ================= stage A: this tile's keys =================================================================
                    const uint2 v = p0 ? tile_block(rb + p0, L - p0) : pre;
                    const u32 shift = (u32)((uintptr_t)(bases + rb + p0) & 7u);
                    u32 c16;
                    const u32 susp = pack8_fast(v, c16);
                    const bool slow = __any_sync(FULL, susp != 0);
                    const u32 word = (c16 << 16) | __shfl_down_sync(FULL, c16, 1);
                    const u32 q0 = shift + PPL * lane, ci = q0 >> 3, s = (q0 & 7u) * 2u;
                    const u32 w0 = __shfl_sync(FULL, word, ci), w1 = __shfl_sync(FULL, word, ci + 2),
                              w2 = __shfl_sync(FULL, word, ci + 4), w3 = __shfl_sync(FULL, word, ci + 6);
                    const u32 left = npos - p0;
                    const u32 nlive = left > PPL * lane ? min((u32)PPL, left - PPL * lane) : 0u;
                    u32 mask = (1u << nlive) - 1;
                    if(slow) {                                             // some staged byte is not ACGTacgt (rare)
                        u32 b8, cc;
                        pack8(v, cc, b8);
                        const u32 bw = (b8 << 8) | __shfl_down_sync(FULL, b8, 1);
                        const u32 b0 = __shfl_sync(FULL, bw, ci), b1 = __shfl_sync(FULL, bw, ci + 2),
                                  b2 = __shfl_sync(FULL, bw, ci + 4), b3 = __shfl_sync(FULL, bw, ci + 6);
                        const u64 B = ((u64)b0 << 48) | ((u64)b1 << 32) | ((u64)b2 << 16) | (u64)b3;
    #pragma unroll
                        for(int i = 0; i < PPL; ++i)
                            if(((B << ((q0 & 7u) + i)) >> (64 - k)) != 0) mask &= ~(1u << i);
                        if(COUNTS) emit += __reduce_add_sync(FULL, __popc(mask));
                    } else if(COUNTS) emit += min(left, (u32)TILE);
                    const u32 A = __funnelshift_l(w1, w0, s), B_ = __funnelshift_l(w2, w1, s), C = __funnelshift_l(w3, w2, s);
                    u32 R0 = __brev(C), R1 = __brev(B_), R2 = __brev(A);
                    R0 = ~(((R0 >> 1) & 0x55555555u) | ((R0 & 0x55555555u) << 1));
                    R1 = ~(((R1 >> 1) & 0x55555555u) | ((R1 & 0x55555555u) << 1));
                    R2 = ~(((R2 >> 1) & 0x55555555u) | ((R2 & 0x55555555u) << 1));
                    u32 xls[PPL], xhs[PPL], fwdm = 0xfu;
    #pragma unroll
                    for(int i = 0; i < PPL; ++i) {
                        u32 xl, xh;                                        // forward k-mer: window bits [2i, 2i + 62)
                        if(i == 0) { xl = __funnelshift_r(B_, A, 2); xh = A >> 2; }
                        else if(i == 1) { xl = B_; xh = A & 0x3fffffffu; }
                        else { xl = __funnelshift_l(C, B_, 2 * i - 2); xh = __funnelshift_l(B_, A, 2 * i - 2) & 0x3fffffffu; }
                        const u32 rl = __funnelshift_r(R2, R1, 2 * i) & kmask_lo, rh = __funnelshift_r(R1, R0, 2 * i) & kmask_hi;
                        const u64 f64 = ((u64)xh << 32) | xl, r64 = ((u64)rh << 32) | rl;
                        const bool lt = f64 < r64;
                        xl = lt ? xl : rl; xh = lt ? xh : rh;
                        if(!lt) fwdm &= ~(1u << i);
                        xls[i] = xl; xhs[i] = xh;
                    }
                    u32 hb[PPL], th[PPL], tl[PPL];
                    if(left <= 120u) {
                        // minimizers of all four k-mers from 19 hashed canonical 16-mers (8 computed, 11 by SHFL.DOWN): bns_classify_u.cuh
                        u32 key[19];
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
                        for(int j2 = 0; j2 < 19; ++j2) key[j2] |= (u32)j2;
                        u32 coreL = key[3], coreR = key[3] ^ 31u;
    #pragma unroll
                        for(int j2 = 4; j2 <= 15; ++j2) { coreL = min(coreL, key[j2]); coreR = min(coreR, key[j2] ^ 31u); }
                        const u32 bl = Pc.b - LOC_GB, lmask = (1u << bl) - 1, xmask = ~((1u << (bl - 1)) - 1);
    #pragma unroll
                        for(int i = 0; i < PPL; ++i) {
                            u32 mL = coreL, mR = coreR;
    #pragma unroll
                            for(int j2 = 0; j2 < 19; ++j2)
                                if(j2 >= i && j2 <= i + 15 && (j2 < 3 || j2 > 15)) { mL = min(mL, key[j2]); mR = min(mR, key[j2] ^ 31u); }
                            const bool fw = (fwdm >> i) & 1u;
                            const u32 bp = fw ? (mL & 31u) - (u32)i : (mR & 31u) + (u32)i - 16u;
                            const u32 xl = xls[i], xh = xhs[i];
                            const u32 m16 = __funnelshift_r(xl, xh, 30u - 2u * bp);
                            const u32 r16 = rc16(m16), pl = loc_place(loc_hash(min(m16, r16)));
                            hb[i] = ((pl & lmask) << LOC_GB) | ((bp & 3u) << 1);
                            tl[i] = ((pl >> 1) & xmask) | (r16 < m16 ? 0x80000000u : 0u);
                            th[i] = (__funnelshift_l(xh << 2, xl, 2u * bp) << 2) | (bp >> 2);
                        }
                    } else {
    #pragma unroll
                        for(int i = 0; i < PPL; ++i) {
                            const TableHash t = loc_encode(((u64)xhs[i] << 32) | xls[i], k, Pc.b);
                            hb[i] = (u32)t.home; th[i] = (u32)(t.tag >> 32); tl[i] = (u32)t.tag;
                        }
                    }
    #ifndef LOCP_PREFETCH_MASK
#define LOCP_PREFETCH_MASK 0xf
#endif
#pragma unroll
                    for(int i = 0; i < PPL; ++i)
                        if(((LOCP_PREFETCH_MASK) >> i & 1) && (mask >> i & 1u)) asm volatile("prefetch.global.L2 [%0];" ::"l"(Pc.slots + ((u64)hb[i] << 5)));
                    s_st[(slot_a * 3 + 0) * 32 + lane] = make_uint4(hb[0], hb[1], hb[2], hb[3]);
                    s_st[(slot_a * 3 + 1) * 32 + lane] = make_uint4(th[0], th[1], th[2], th[3]);
                    s_st[(slot_a * 3 + 2) * 32 + lane] = make_uint4(tl[0], tl[1], tl[2], tl[3]);
                    s_msk[slot_a * 32 + lane] = mask;
                }
                // ================= stage B: the tile before this one ==========================================================
                if(pending) stage_b();
                if(npos) {
                    pending = true; pend_slot = slot_a; pend_j = j; pend_first = p0 == 0; pend_last = p0 + TILE >= npos; pend_emit = emit;
                    slot_a ^= 1u;
                }
            }
            if(j < nrec) {
                rb = nb_; L = nl_; nb_ = zb; nl_ = zl;
                tb = (tb + 1) % LOCP_RING;
            }
        }
        // ---- one coalesced store per output array for the batch ------------------------------------------------------
        if(lane < nrec) {
            taxon_out[r0 + lane] = my_taxon;
            if(COUNTS) {
                if(nhit_out) nhit_out[r0 + lane] = my_hit;
                if(nmiss_out) nmiss_out[r0 + lane] = my_miss;
            }
        }
        const u32 cls = __popc(__ballot_sync(FULL, lane < nrec && my_taxon != 0));
        if(lane == 0) {                                                // classified_[2] (classifier.h:138,238), once per batch
            atomicAdd(&counters[0], (unsigned long long)cls);
            atomicAdd(&counters[1], (unsigned long long)(nrec - cls));
        }
        pb ^= 1;
    }
}

}  // namespace bns
