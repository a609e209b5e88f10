"""GPU (-m gpu): the CUDA path, called through the C ABI (bonsai_b200/capi.py -> libbonsai_b200.so), against
the CPU oracle on the same seeded inputs and against the golden vectors the unmodified reference produced.
Bit-exact everywhere: k-mer streams, hit/miss + value of every lookup, per-record taxon / hit / missing counts
and the ordered hit lists."""
import hashlib
import random

import numpy as np
import pytest

import helpers as H
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from bonsai_b200 import capi as m
    m.load_library()          # fails loudly if the extension was not built
    return m


def hx(a):
    return [format(int(x), "x") for x in a]


def gpu_encode_one(capi, seq, k, w, gaps, score, canon, api, cast):
    with capi.Context(k, w, gaps, score, canon, api, cast) as ctx:
        b, o = po.pack_reads([seq])
        return ctx.encode_lists(b, o)[0]


def supported(k, w, gaps, score, canon, api):
    """every Spacer/Encoder combination the reference accepts is supported"""
    return True


def test_encode_small_golden(capi, golden):
    ctxs = {}
    n = 0
    for e in golden["encode_small"]:
        if not supported(e["k"], e["w"], e["gaps"], e["score"], e["canon"], e["api"]):
            with pytest.raises(capi.BnsError):
                capi.Context(e["k"], e["w"], e["gaps"], e["score"], e["canon"], e["api"])
            continue
        for cast, key in ((capi.CAST_SATURATE, "saturate"), (capi.CAST_WRAP, "wrap")):
            exp = e[key] if e[key] is not None else e["saturate"]
            ck = (e["k"], e["w"], tuple(e["gaps"] or ()), e["score"], e["canon"], e["api"], cast)
            if ck not in ctxs:
                ctxs[ck] = capi.Context(e["k"], e["w"], e["gaps"], e["score"], e["canon"], e["api"], cast)
            b, o = po.pack_reads([e["seq"]])
            got = hx(ctxs[ck].encode_lists(b, o)[0])
            assert got == exp, (e, key)
            n += 1
    assert n > 1000
    for c in ctxs.values():
        c.close()


@pytest.mark.parametrize("tag", ["saturate", "wrap"])
def test_encode_streams_golden(capi, golden, genomes, reads2000, tag):
    bases, offs, _ = reads2000
    cast = capi.CAST_SATURATE if tag == "saturate" else capi.CAST_WRAP
    pb, poff = po.pack_reads([bytes(genomes["phix"])])
    for name, s in golden["streams"].items():
        if not name.endswith(":" + tag):
            continue
        with capi.Context(s["k"], s["w"], s["gaps"], s["score"], s["canon"], s["api"], cast) as ctx:
            kmers, oo, cnt = ctx.encode(bases, offs)
            allk = np.concatenate([kmers[int(oo[i]):int(oo[i]) + int(cnt[i])] for i in range(cnt.size)])
            assert list(H.digest(allk)) == s["reads"], name
            px = ctx.encode_lists(pb, poff)[0]
            assert list(H.digest(px)) == s["phix"], name
            assert int(np.unique(px).size) == s["phix_distinct"], name


def test_encode_fuzz_vs_oracle(capi, oracle):
    rng = random.Random(99)
    for it in range(60):
        k = rng.choice([1, 2, 3, 5, 7, 13, 16, 21, 31, 31, 31, 32])
        gaps = None
        if rng.random() < 0.35 and k > 1:
            gaps = [rng.choice([0, 0, 0, 1, 2, 3]) for _ in range(k - 1)]
        c = k + (sum(gaps) if gaps else 0)
        w = rng.choice([0, k, c, c + 1, c + 3, c + 19, c + 50, c + 200])
        seqs = []
        for _ in range(40):
            L = rng.choice([0, 1, k - 1, k, c, c + 1, c + 5, 60, 150, 300, 700, 1500])
            alphabet = rng.choice(["ACGT", "ACGT", "ACGTN", "ACGTacgtNnUu-", "AT", "A", "T", "AC"])
            s = "".join(rng.choice(alphabet) for _ in range(L))
            if rng.random() < 0.2 and L > 40:
                p = rng.randrange(L - 35)
                s = s[:p] + rng.choice("ACGT") * 35 + s[p + 35:]
            seqs.append(s)
        b, o = po.pack_reads(seqs)
        for score in (0, 1):
            for canon in (0, 1):
                for api in (0, 1):
                    if not supported(k, w, gaps, score, canon, api):
                        continue
                    cast = rng.choice([capi.CAST_SATURATE, capi.CAST_WRAP])
                    with capi.Context(k, w, gaps, score, canon, api, cast) as ctx:
                        got = ctx.encode_lists(b, o)
                    for s, g in zip(seqs, got):
                        exp = oracle.encode(s, k, w, gaps, score, canon, api, cast_mode=cast)
                        assert np.array_equal(g, exp), dict(k=k, w=w, gaps=gaps, score=score, canon=canon, api=api, cast=cast, seq=s)


def test_lookup_matches_kh_get(capi, oracle, dbcache):
    db = dbcache.get("ent_k31_w50")
    keys, vals, flags, nb, size = oracle.db_arrays(db)
    with capi.Context(31) as ctx:
        ctx.load_table(keys, vals, flags, nb)          # the raw khash arrays, as Database::db_ holds them
        info = ctx.table_info()
        assert info["n_keys"] == size
        k, v = oracle.db_pairs(db)
        gv, gf = ctx.lookup(k)
        assert gf.all() and np.array_equal(gv, v)
        rng = np.random.default_rng(3)
        q = rng.integers(0, 2**62, 200000, dtype=np.uint64)
        q[::7] = k[rng.integers(0, k.size, q[::7].size)]
        gv, gf = ctx.lookup(q)
        pos = np.searchsorted(k, q)
        pos[pos == k.size] = 0
        hit = k[pos] == q
        assert np.array_equal(gf, hit)
        assert np.array_equal(gv[hit], v[pos][hit])
        # load_pairs builds the same table
        with capi.Context(31) as ctx2:
            ctx2.load_pairs(k, v)
            gv2, gf2 = ctx2.lookup(q)
            assert np.array_equal(gf2, hit) and np.array_equal(gv2[hit], v[pos][hit])


def test_lookup_small_and_adversarial(capi):
    """tiny tables, many distinct values, keys that collide in the low / high bits"""
    rng = np.random.default_rng(17)
    for n, nvals in ((0, 1), (1, 1), (5, 5), (1000, 1000), (70000, 50000), (300000, 7)):
        keys = np.unique(rng.integers(0, 2**64 - 1, n, dtype=np.uint64))
        if n >= 1000:
            keys[: n // 4] = np.arange(n // 4, dtype=np.uint64) << np.uint64(40)     # only high bits differ
            keys[n // 4: n // 2] = np.arange(n // 4, n // 2, dtype=np.uint64)          # dense small integers
            keys = np.unique(keys)
        vals = rng.integers(1, 2**32 - 1, max(nvals, 1), dtype=np.uint64).astype(np.uint32)[rng.integers(0, max(nvals, 1), keys.size)]
        with capi.Context(31) as ctx:
            ctx.load_pairs(keys, vals)
            assert ctx.table_info()["n_keys"] == keys.size
            gv, gf = ctx.lookup(keys)
            assert gf.all() and np.array_equal(gv, vals)
            miss = keys ^ np.uint64(1 << 63) if keys.size else np.array([5], np.uint64)
            miss = miss[~np.isin(miss, keys)]
            _, gf = ctx.lookup(miss)
            assert not gf.any()


def test_resolve_golden(capi, golden):
    c, p = H.toy_tax_arrays()
    with capi.Context(31) as ctx:
        vals = np.array([1, 2, 10, 11, 12, 13, 20], np.uint32)
        ctx.load_pairs(np.arange(vals.size, dtype=np.uint64), vals)
        ctx.load_taxonomy(c, p)
        cases = [x for x, _ in golden["resolve"]]
        got = ctx.resolve(cases)
        assert got.tolist() == [e for _, e in golden["resolve"]]


def test_resolve_fuzz_vs_oracle(capi, oracle):
    rng = np.random.default_rng(5)
    ids = np.unique(rng.integers(2, 500000, 3000))[:2500]
    nodes = np.concatenate([[1], ids]).astype(np.uint32)
    parent = np.zeros(nodes.size, np.uint32)
    parent[0] = 1
    for i in range(1, nodes.size):
        parent[i] = nodes[rng.integers(max(0, i - 40), i)]       # deep-ish tree
    T = oracle.tax_from_pairs(nodes, parent)
    with capi.Context(31) as ctx:
        ctx.load_pairs(np.arange(nodes.size, dtype=np.uint64), nodes)
        ctx.load_taxonomy(nodes, parent)
        cases = []
        for _ in range(3000):
            n = int(rng.integers(1, 40))
            taxa = rng.choice(nodes, n, replace=False)
            cnt = rng.integers(1, 4, n)
            cases.append(list(zip(taxa.tolist(), cnt.tolist())))
        got = ctx.resolve(cases)
        exp = [oracle.resolve(T, [a for a, _ in cs], [b for _, b in cs]) for cs in cases]
        assert got.tolist() == exp


@pytest.fixture(scope="module")
def gpu_dbs(capi, oracle, dbcache, golden):
    """one context per golden classify case, loaded from the oracle-built khash arrays"""
    c, p = H.toy_tax_arrays()
    out = {}

    def make(cname):
        if cname not in out:
            spec = golden["classify"][cname]
            keys, vals, flags, nb, _ = oracle.db_arrays(dbcache.get(spec["db"]))
            ctx = capi.Context(spec["k"], spec["w"], spec["gaps"], capi.SCORE_LEX, spec["canon"], spec["api"])
            ctx.load_table(keys, vals, flags, nb)
            ctx.load_taxonomy(c, p)
            out[cname] = ctx
        return out[cname]
    yield make
    for ctx in out.values():
        ctx.close()


@pytest.mark.parametrize("cname", ["config1_lex_w31", "config2_entdb", "config4_spaced", "windowed_lex_w50_on_reads",
                                   "windowed_lex_w50_nocanon"])
def test_classify_golden(capi, golden, gpu_dbs, reads2000, cname):
    bases, offs, _ = reads2000
    spec = golden["classify"][cname]
    ctx = gpu_dbs(cname)
    taxon, nhit, nmiss, lists = ctx.classify(bases, offs, want_taxa=True)
    assert taxon.tolist() == spec["taxon"]
    assert nhit.tolist() == spec["nhit"]
    assert nmiss.tolist() == spec["nmiss"]
    assert hashlib.md5(np.concatenate(lists).tobytes()).hexdigest() == spec["taxa_md5"]
    # the lean kernel (no hit list; with and without counts) agrees
    t2, _, _ = ctx.classify(bases, offs, want_counts=False)
    assert np.array_equal(t2, taxon)
    t3, h3, m3 = ctx.classify(bases, offs)
    assert np.array_equal(t3, taxon) and h3.tolist() == spec["nhit"] and m3.tolist() == spec["nmiss"]
    st = ctx.stats()
    assert st["kernel_launches"] > 0


def test_classify_runs_match_hit_lists(capi, golden, gpu_dbs, reads2000):
    """bns_b200_classify_batch_runs: the device-side run-length encoding of the ordered hit lists equals the encoding of
    the lists bns_b200_classify_batch returns (single-end and paired), and too small a run buffer is reported."""
    bases, offs, _ = reads2000
    ctx = gpu_dbs("config1_lex_w31")
    for paired in (False, True):
        taxon, nhit, nmiss, lists = ctx.classify(bases, offs, paired=paired, want_taxa=True)
        t2, h2, m2, runs = ctx.classify_runs(bases, offs, paired=paired)
        assert np.array_equal(t2, taxon) and np.array_equal(h2, nhit) and np.array_equal(m2, nmiss)
        for lst, rn in zip(lists, runs):
            if lst.size == 0:
                assert rn.shape[0] == 0
                continue
            cut = np.flatnonzero(np.diff(lst.astype(np.int64)) != 0) + 1
            starts = np.concatenate([[0], cut])
            lens = np.diff(np.concatenate([starts, [lst.size]]))
            assert np.array_equal(rn[:, 0], lst[starts]) and np.array_equal(rn[:, 1], lens.astype(np.uint32))
    with pytest.raises(capi.BnsError):
        ctx.classify_runs(bases, offs, cap=10)


def _rle(lst):
    if lst.size == 0:
        return np.zeros((0, 2), np.uint32)
    starts = np.concatenate([[0], np.flatnonzero(np.diff(lst.astype(np.int64)) != 0) + 1])
    lens = np.diff(np.concatenate([starts, [lst.size]]))
    return np.stack([lst[starts], lens.astype(np.uint32)], axis=1)


@pytest.mark.parametrize("layout", ["hash", "minimizer"])
def test_classify_runs_long_records_many_runs(capi, oracle, toy_tax, genomes, monkeypatch, layout):
    """Run lists out of the lean kernel where one record holds far more runs than its per-warp buffer (k-mers of a 6 kb
    stretch valued 11 / 12 / 13 by position, so nearly every hit starts a run), runs that continue across tiles and across
    the mates of a pair, records without hits and empty records between them -- against the oracle's ordered hit lists."""
    monkeypatch.setenv("BNS_B200_LAYOUT", layout)
    b, _ = H.genome_records(genomes, 0)
    region = b[10_000:16_000]
    lut = np.zeros(256, np.uint64)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    c = lut[region]
    n = c.size - 30
    f = np.zeros(n, np.uint64); r = np.zeros(n, np.uint64)
    for j in range(31):
        f = (f << np.uint64(2)) | c[j:j + n]
        r = r | ((np.uint64(3) - c[j:j + n]) << np.uint64(2 * j))
    km = np.minimum(f, r)
    keys, first = np.unique(km, return_index=True)
    pos = first                                               # value by the position of the k-mer's first occurrence
    vals = np.where(pos % 7 < 3, 11, np.where(pos % 7 < 5, 12, 13)).astype(np.uint32)
    vals[(pos // 400) % 3 == 1] = 12                          # and stretches of equal values: long runs across tile borders
    dbo = oracle.db_from_pairs(keys, vals)
    rng = np.random.default_rng(5)
    reads = [bytes(region), b"", bytes(region[100:250]), bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 300)]),
             bytes(region[2000:2900]), bytes(region[3000:3100]), bytes(region[:31]), bytes(region[5:35])]
    reads += [bytes(region[s:s + int(l)]) for s, l in zip(rng.integers(0, 5000, 200), rng.integers(0, 700, 200))]
    bases, offs = po.pack_reads(reads)
    tc, tp = H.toy_tax_arrays()
    with capi.Context(31, 31) as ctx:
        ctx.load_pairs(keys, vals)
        ctx.load_taxonomy(tc, tp)
        assert ctx.table_info()["layout"] == (1 if layout == "minimizer" else 0)
        for paired in (False, True):
            et, eh, em, lists = oracle.classify(dbo, toy_tax, bases, offs, 31, 31, paired=paired, want_taxa=True)
            t, h, m, runs = ctx.classify_runs(bases, offs, paired=paired)
            assert np.array_equal(t, et) and np.array_equal(h, eh) and np.array_equal(m, em)
            assert max(len(x) for x in runs) > 1000            # the long record really overflows the per-warp buffer
            for lst, rn in zip(lists, runs):
                assert np.array_equal(rn, _rle(lst))


def test_records_with_more_distinct_taxa_than_shared_memory_holds(capi, oracle, genomes):
    """linear::counter has no limit on distinct taxa (linear.h:229); the kernels keep 256 per record in shared memory. A
    database of 900 values and records that hit up to 900 of them: the second pass (lists in global memory) must give the
    reference's answer for every entry point -- taxon only, counts, hit lists, run lists, pairs, a windowed encoder."""
    b, _ = H.genome_records(genomes, 1)
    region = b[50_000:50_930]
    lut = np.zeros(256, np.uint64)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    c = lut[region]
    n = c.size - 30
    f = np.zeros(n, np.uint64); r = np.zeros(n, np.uint64)
    for j in range(31):
        f = (f << np.uint64(2)) | c[j:j + n]
        r = r | ((np.uint64(3) - c[j:j + n]) << np.uint64(2 * j))
    keys, first = np.unique(np.minimum(f, r), return_index=True)
    vals = (1000 + first).astype(np.uint32)                       # one taxon per k-mer position
    tc = np.concatenate([[1], 50 + np.arange(7), 1000 + np.arange(n)]).astype(np.uint32)
    tp = np.concatenate([[1], np.ones(7), 50 + (np.arange(n) % 7)]).astype(np.uint32)
    T = oracle.tax_from_pairs(tc, tp)
    dbo = oracle.db_from_pairs(keys, vals)
    reads = [bytes(region), bytes(region[:400]), bytes(region[:150]), b"", bytes(region[300:930]), bytes(region[100:420])] * 3
    reads.append(bytes(region[:286]))                             # 256 hits: just fits
    reads.append(bytes(region[:287]))                             # 257: the first that does not
    bases, offs = po.pack_reads(reads)
    with capi.Context(31, 31) as ctx:
        ctx.load_pairs(keys, vals)
        ctx.load_taxonomy(tc, tp)
        for paired in (False, True):
            et, eh, em, lists = oracle.classify(dbo, T, bases, offs, 31, 31, paired=paired, want_taxa=True)
            assert eh.max() > 256
            t, h, m = ctx.classify(bases, offs, paired=paired)
            assert np.array_equal(t, et) and np.array_equal(h, eh) and np.array_equal(m, em)
            t, h, m, gl = ctx.classify(bases, offs, paired=paired, want_taxa=True)
            assert np.array_equal(t, et) and all(np.array_equal(a, x) for a, x in zip(lists, gl))
            t, h, m, runs = ctx.classify_runs(bases, offs, paired=paired)
            assert np.array_equal(t, et) and np.array_equal(h, eh) and all(np.array_equal(rn, _rle(x)) for rn, x in zip(runs, lists))
        st = ctx.stats()
        assert st["n_classified"] + st["n_unclassified"] == 3 * (len(reads) + len(reads) // 2)    # every record counted exactly once
    with capi.Context(31, 40) as ctx:                             # windowed (lean LEAN_K + second pass)
        ctx.load_pairs(keys, vals)
        ctx.load_taxonomy(tc, tp)
        et, eh, em = oracle.classify(dbo, T, bases, offs, 31, 40)
        t, h, m = ctx.classify(bases, offs)
        assert eh.max() > 256 and np.array_equal(t, et) and np.array_equal(h, eh) and np.array_equal(m, em)


def test_classify_phix_and_paired(capi, golden, gpu_dbs, reads2000, genomes):
    ctx = gpu_dbs("config1_lex_w31")
    pb, poff = po.pack_reads([bytes(genomes["phix"])])
    taxon, nhit, nmiss = ctx.classify(pb, poff)
    g = golden["classify"]["phix"]
    assert (int(taxon[0]), int(nhit[0]), int(nmiss[0])) == (g["taxon"], g["nhit"], g["nmiss"]) == (0, 0, 5356)
    bases, offs, _ = reads2000
    g = golden["classify"]["paired_lex_w31"]
    taxon, nhit, nmiss = ctx.classify(bases, offs, paired=True)
    assert taxon.tolist() == g["taxon"] and nhit.tolist() == g["nhit"] and nmiss.tolist() == g["nmiss"]


def test_classify_ragged_vs_oracle(capi, oracle, dbcache, toy_tax, gpu_dbs, genomes):
    """empty reads, reads shorter than k, long reads (multi-tile), contig-sized records"""
    bases, offs, _ = H.make_reads(3000, seed=7, ragged=True)
    ctx = gpu_dbs("config1_lex_w31")
    db = dbcache.get("lex_k31_w31")
    exp = oracle.classify(db, toy_tax, bases, offs, 31, 31, want_taxa=True)
    got = ctx.classify(bases, offs, want_taxa=True)
    for a, b in zip(exp[:3], got[:3]):
        assert np.array_equal(a, b)
    assert all(np.array_equal(a, b) for a, b in zip(exp[3], got[3]))
    # a few long records: 20 kb slices of the genomes (hundreds of tiles per warp)
    g = genomes
    rng = np.random.default_rng(1)
    recs = []
    for _ in range(12):
        s = int(rng.integers(0, g["bases"].size - 20000))
        recs.append(bytes(g["bases"][s:s + int(rng.integers(2000, 20000))]))
    b, o = po.pack_reads(recs)
    exp = oracle.classify(db, toy_tax, b, o, 31, 31)
    got = ctx.classify(b, o)
    for a, bb in zip(exp, got):
        assert np.array_equal(a, bb)
    # the ragged reads as mate pairs (lean kernel, two sequences per record)
    ne = ((offs.size - 1) // 2) * 2
    expp = oracle.classify(db, toy_tax, bases[:int(offs[ne])], offs[:ne + 1], 31, 31, paired=True)
    gotp = ctx.classify(bases[:int(offs[ne])], offs[:ne + 1], paired=True)
    for a, bb in zip(expp, gotp):
        assert np.array_equal(a, bb)
    # empty batch
    t, h, m = ctx.classify(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert t.size == 0


@pytest.mark.parametrize("mode", ["pack", "hybrid", "auto"])
def test_host_packed_chunks_vs_oracle(capi, oracle, dbcache, toy_tax, genomes, golden, reads2000, monkeypatch, mode):
    """bns_b200_classify_batch with host packing threads (bns_b200_config.host_pack_threads): chunks packed to 2 bits on the
    host and read by the packed-input kernel, alone ("pack") and next to chunks that cross as ASCII ("hybrid"), against the
    oracle and the ASCII-only call: ragged, empty and long reads, N / lower case / arbitrary bytes, records that straddle
    units, chunks and suspicious-bit words, a fixed-length batch (offsets generated on the device), mate pairs, both layouts."""
    monkeypatch.setenv("BNS_B200_PACK_MIN_BASES", "1")
    monkeypatch.setenv("BNS_B200_PACK_CHUNK_READS", "700")
    if mode != "auto":                                    # auto: numpy arrays are pageable memory, every chunk is packed
        monkeypatch.setenv("BNS_B200_HOST_PACK_MODE", mode)
    else:
        monkeypatch.delenv("BNS_B200_HOST_PACK_MODE", raising=False)
    db = dbcache.get("lex_k31_w31")
    keys, vals, flags, nb, _ = oracle.db_arrays(db)
    rb, ro, _ = H.make_reads(6000, seed=41, ragged=True)
    reads = [bytes(rb[int(ro[i]):int(ro[i + 1])]) for i in range(ro.size - 1)]
    rng = np.random.default_rng(5)
    g = genomes
    for i in range(0, len(reads), 7):                     # lower case, runs of N, arbitrary bytes
        r = bytearray(reads[i])
        if not r:
            continue
        if i % 3 == 0:
            r = bytearray(bytes(r).lower())
        elif i % 3 == 1:
            a, nn = int(rng.integers(0, len(r))), int(rng.integers(1, 12))
            r[a:a + nn] = b"N" * len(r[a:a + nn])
        else:
            for _ in range(3):
                r[int(rng.integers(0, len(r)))] = int(rng.integers(0, 256))
        reads[i] = bytes(r)
    for _ in range(6):                                    # long records: many tiles, later tiles read the unit stream directly
        s0 = int(rng.integers(0, g["bases"].size - 9000))
        r = bytearray(g["bases"][s0:s0 + int(rng.integers(500, 9000))].tobytes())
        r[len(r) // 2] = ord("n")
        reads.insert(int(rng.integers(0, len(reads))), bytes(r))
    bases, offs = po.pack_reads(reads)
    fixed_b, fixed_o, _ = H.make_reads(5000, seed=43)
    # the spaced-seed encoder (BASELINE configs[3]) reads packed chunks too: the reference-generated golden answers
    spec = golden["classify"]["config4_spaced"]
    sk, sv, sf, snb, _ = oracle.db_arrays(dbcache.get(spec["db"]))
    with capi.Context(spec["k"], spec["w"], spec["gaps"], capi.SCORE_LEX, spec["canon"], spec["api"], host_pack_threads=3) as sctx:
        sctx.load_table(sk, sv, sf, snb)
        sctx.load_taxonomy(*H.toy_tax_arrays())
        h0 = sctx.stats()["h2d_bytes"]
        t, nh, nm = sctx.classify(reads2000[0], reads2000[1])
        assert t.tolist() == spec["taxon"] and nh.tolist() == spec["nhit"] and nm.tolist() == spec["nmiss"]
        if mode != "hybrid":
            assert sctx.stats()["h2d_bytes"] - h0 < reads2000[0].size * 0.3 + 4096 * 16
        db4 = dbcache.get(spec["db"])
        exp4 = oracle.classify(db4, toy_tax, bases, offs, spec["k"], spec["w"], spec["gaps"], 0, spec["canon"], spec["api"])
        got4 = sctx.classify(bases, offs)
        for a, b in zip(exp4, got4):
            assert np.array_equal(a, b)
    for layout in ("hash", "minimizer"):
        monkeypatch.setenv("BNS_B200_LAYOUT", layout)
        with capi.Context(31, 31, host_pack_threads=3) as ctx, capi.Context(31, 31, host_pack_threads=0) as plain:
            for c in (ctx, plain):
                c.load_table(keys, vals, flags, nb)
                c.load_taxonomy(*H.toy_tax_arrays())
            assert ctx.table_info()["layout"] == (1 if layout == "minimizer" else 0)
            exp = oracle.classify(db, toy_tax, bases, offs, 31, 31)
            h0 = ctx.stats()["h2d_bytes"]
            got = ctx.classify(bases, offs)
            packed_bytes = ctx.stats()["h2d_bytes"] - h0
            ref = plain.classify(bases, offs)
            for a, b, c in zip(exp, got, ref):
                assert np.array_equal(a, b) and np.array_equal(a, c)
            if mode != "hybrid":                          # the bases crossed as 2-bit units (+ offsets): well under one byte each
                assert packed_bytes < bases.size * 0.25 + offs.size * 8 + 4096 * 16
            t_only, _, _ = ctx.classify(bases, offs, want_counts=False)
            assert np.array_equal(t_only, exp[0])
            # the worker threads can be switched off and on again on a live context (bns_b200_set_host_pack_threads)
            assert ctx.host_pack_threads() == 3 and plain.host_pack_threads() == 0
            ctx.set_host_pack_threads(0)
            h1 = ctx.stats()["h2d_bytes"]
            off_run = ctx.classify(bases, offs)
            assert all(np.array_equal(a, b) for a, b in zip(exp, off_run)) and ctx.stats()["h2d_bytes"] - h1 >= bases.size
            ctx.set_host_pack_threads(2)
            assert ctx.host_pack_threads() == 2
            on_run = ctx.classify(bases, offs)
            assert all(np.array_equal(a, b) for a, b in zip(exp, on_run))
            ne = ((offs.size - 1) // 2) * 2
            expp = oracle.classify(db, toy_tax, bases[:int(offs[ne])], offs[:ne + 1], 31, 31, paired=True)
            gotp = ctx.classify(bases[:int(offs[ne])], offs[:ne + 1], paired=True)
            for a, b in zip(expp, gotp):
                assert np.array_equal(a, b)
            expf = oracle.classify(db, toy_tax, fixed_b, fixed_o, 31, 31)
            gotf = ctx.classify(fixed_b, fixed_o)
            for a, b in zip(expf, gotf):
                assert np.array_equal(a, b)
            st, sp = ctx.stats(), plain.stats()
            assert st["n_classified"] == sp["n_classified"] + 3 * int((exp[0] != 0).sum()) + int((expp[0] != 0).sum()) + int((expf[0] != 0).sum())


@pytest.mark.parametrize("k,canon", [(21, True), (31, False), (32, True), (16, False)])
def test_host_packed_other_k_and_strands(capi, oracle, toy_tax, genomes, monkeypatch, k, canon):
    """The packed-input variants for k other than 31 (run-time k) and for uncanonical k-mers: a database of the first 120 kb of
    each genome built on the device for that encoder, ragged reads with N / lower case, packing threads on vs the oracle."""
    from bonsai_b200 import dbbuild, workload as W
    monkeypatch.setenv("BNS_B200_PACK_MIN_BASES", "1")
    monkeypatch.setenv("BNS_B200_PACK_CHUNK_READS", "900")
    monkeypatch.setenv("BNS_B200_HOST_PACK_MODE", "pack")
    tc, tp = H.toy_tax_arrays()
    g = W.load_genomes()
    gen = []
    for gi in range(4):
        b, off = W.genome_records(g, gi)
        gen.append((b[:120_000].copy(), np.array([0, 120_000], np.uint64)))
    with capi.Context(k, k, canonicalize=canon) as bctx:
        dbbuild.build_on_device(bctx, gen, W.GENOME_TAXIDS, tc, tp, k, k, canonicalize=canon)
        keys, vals = bctx.table_dump()
    db = oracle.db_from_pairs(keys, vals)
    rng = np.random.default_rng(k)
    reads = []
    for i in range(5000):
        src = gen[i % 4][0]
        L = int(rng.integers(0, 260))
        s0 = int(rng.integers(0, src.size - 300))
        r = bytearray(src[s0:s0 + L].tobytes())
        if r and i % 9 == 0:
            r[int(rng.integers(0, len(r)))] = ord("N")
        if i % 13 == 0:
            r = bytearray(bytes(r).lower())
        if i % 2:
            r = bytearray(bytes(r)[::-1].translate(bytes.maketrans(b"ACGTacgt", b"TGCAtgca")))
        reads.append(bytes(r))
    bases, offs = po.pack_reads(reads)
    exp = oracle.classify(db, toy_tax, bases, offs, k, k, None, 0, canon, capi.API_STRING)
    with capi.Context(k, k, canonicalize=canon, host_pack_threads=3) as ctx:
        ctx.load_pairs(keys, vals)
        ctx.load_taxonomy(tc, tp)
        h0 = ctx.stats()["h2d_bytes"]
        got = ctx.classify(bases, offs)
        assert ctx.stats()["h2d_bytes"] - h0 < bases.size * 0.3 + offs.size * 8 + 4096 * 16      # it did go packed
        for a, b in zip(exp, got):
            assert np.array_equal(a, b)
    assert int((exp[0] != 0).sum()) > 1000


def test_chimeric_reads_tied_taxa_vs_oracle(capi, oracle, dbcache, toy_tax, gpu_dbs, genomes):
    """resolve_tree's tie rule (equal root-path scores -> lca of the tied taxa, util.h:850-866) on reads built to tie: two to four
    segments of equal length from different genomes, so that sibling taxa, cousins and three-way ties occur; single reads and
    pairs, with and without counts. The small value dictionary of this database runs the lean kernel's SV variants (counts in
    lane registers, ties folded in value-id order), which must give the reference's taxon."""
    g = genomes
    rng = np.random.default_rng(77)
    coff = g["contig_off"]
    by_g = [np.nonzero((g["contig_genome"] == x) & ((coff[1:] - coff[:-1]) >= 400))[0] for x in range(4)]
    reads = []
    for i in range(4000):
        nseg = int(rng.integers(2, 5))
        seg = int(rng.choice([45, 60, 75]))
        gs = rng.choice(4, size=nseg, replace=False)
        parts = []
        for x in gs:
            c = int(by_g[x][rng.integers(0, by_g[x].size)])
            s0 = int(coff[c]) + int(rng.integers(0, int(coff[c + 1] - coff[c]) - seg))
            parts.append(g["bases"][s0:s0 + seg].tobytes())
        reads.append(b"".join(parts))
    bases, offs = po.pack_reads(reads)
    db = dbcache.get("lex_k31_w31")
    ctx = gpu_dbs("config1_lex_w31")
    assert ctx.table_info()["n_values"] <= 32
    exp = oracle.classify(db, toy_tax, bases, offs, 31, 31)
    got = ctx.classify(bases, offs)
    for a, b in zip(exp, got):
        assert np.array_equal(a, b)
    t_only, _, _ = ctx.classify(bases, offs, want_counts=False)
    assert np.array_equal(t_only, exp[0])
    expp = oracle.classify(db, toy_tax, bases, offs, 31, 31, paired=True)
    gotp = ctx.classify(bases, offs, paired=True)
    for a, b in zip(expp, gotp):
        assert np.array_equal(a, b)
    # the reads do tie: inner taxonomy nodes win a good share of them
    leaves = np.isin(exp[0], [11, 12, 13, 20])
    assert int((~leaves & (exp[0] != 0)).sum()) > 300


def test_classify_device_and_replication(capi, golden, gpu_dbs, reads2000):
    """device-resident call + the broadcast path (header, segments, commit) into a second context"""
    import torch
    bases, offs, _ = reads2000
    spec = golden["classify"]["config1_lex_w31"]
    src = gpu_dbs("config1_lex_w31")
    hdr = src.db_export_header()
    with capi.Context(31, 31) as dst:
        dst.db_alloc_from_header(hdr)
        for (sp, sb), (dp, db) in zip(src.db_segments(), dst.db_segments()):
            assert sb == db
            if sb:           # plain device-to-device copies stand in for the NCCL broadcast
                torch.as_tensor(capi.DevMem(dp, db), device="cuda").copy_(torch.as_tensor(capi.DevMem(sp, sb), device="cuda"))
        torch.cuda.synchronize()
        dst.db_commit()
        d_b = torch.from_numpy(bases).cuda()
        d_o = torch.from_numpy(offs.astype(np.int64)).cuda()
        n = offs.size - 1
        d_t = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_h = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_m = torch.zeros(n, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        dst.classify_device(d_b.data_ptr(), d_o.data_ptr(), n, d_t.data_ptr(), d_h.data_ptr(), d_m.data_ptr(), stream=st)
        torch.cuda.synchronize()
        assert d_t.cpu().numpy().astype(np.uint32).tolist() == spec["taxon"]
        assert d_h.cpu().numpy().tolist() == spec["nhit"]
        assert d_m.cpu().numpy().tolist() == spec["nmiss"]
        s = dst.stats()
        assert s["n_classified"] + s["n_unclassified"] == n
        assert s["n_unclassified"] == spec["taxon"].count(0)


def test_full_size_properties(capi, gpu_dbs, golden, oracle, dbcache, toy_tax):
    """BASELINE-scale batch (1M reads): size-independent properties -- determinism, batch-split invariance,
    reverse-complement invariance of canonical classification, and hit + missing == valid windows."""
    ctx = gpu_dbs("config1_lex_w31")
    bases, offs, _ = H.make_reads(1_000_000, seed=123, frac_n=0.0)
    t1, h1, m1 = ctx.classify(bases, offs)
    t2, h2, m2 = ctx.classify(bases, offs)
    assert np.array_equal(t1, t2) and np.array_equal(h1, h2)
    assert ((h1 + m1) == 120).all()                      # no N: every window is looked up exactly once
    half = 400_003
    ta, _, _ = ctx.classify(bases[: int(offs[half])], offs[: half + 1])
    tb, _, _ = ctx.classify(bases[int(offs[half]):], offs[half:] - offs[half])
    assert np.array_equal(np.concatenate([ta, tb]), t1)
    # reverse-complementing every read leaves canonical k-mer sets, hence taxa, unchanged
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    rc = comp[bases.reshape(-1, 150)[:, ::-1]].reshape(-1)
    t3, h3, m3 = ctx.classify(np.ascontiguousarray(rc), offs)
    assert np.array_equal(t3, t1) and np.array_equal(h3, h1)
    # and the first 20k reads agree with the CPU oracle record by record
    n = 20000
    et, eh, em = oracle.classify(dbcache.get("lex_k31_w31"), toy_tax, bases[: int(offs[n])], offs[: n + 1], 31, 31, nthreads=0)
    assert np.array_equal(et, t1[:n]) and np.array_equal(eh, h1[:n]) and np.array_equal(em, m1[:n])


def test_stress_device_build_vs_oracle(capi, oracle, toy_tax):
    """BASELINE config 5 at test scale: table built from device-resident keys, reads cut from the key stream."""
    import torch
    from bonsai_b200 import workload as W
    dev = torch.device("cuda", 0)
    n_keys = (1 << 20) + 12345
    stream, d_keys, d_vals = W.make_stress_db(n_keys, seed=5, device=dev)
    d_bases, d_offs, from_db = W.make_stress_reads(stream, 20000, seed=6, device=dev)
    c, p = H.toy_tax_arrays()
    with capi.Context(31, 31) as ctx:
        ctx.load_pairs_device(d_keys.data_ptr(), d_vals.data_ptr(), n_keys, W.STRESS_VALUES)
        ctx.load_taxonomy(c, p)
        info = ctx.table_info()
        keys = d_keys.cpu().numpy().astype(np.uint64)
        vals = d_vals.cpu().numpy().astype(np.uint32)
        assert info["n_keys"] == np.unique(keys).size
        assert info["n_keys"] / info["n_buckets"] > 0.6
        gv, gf = ctx.lookup(keys)
        assert gf.all() and np.array_equal(gv, vals)
        bases = d_bases.cpu().numpy()
        offs = d_offs.cpu().numpy().astype(np.uint64)
        t, h, m = ctx.classify(bases, offs)
        fd = from_db.cpu().numpy()
        assert (h[fd] == 120).all() and (h[~fd] == 0).all()
        uk, ui = np.unique(keys, return_index=True)
        D = oracle.db_from_pairs(uk, vals[ui])
        et, eh, em = oracle.classify(D, toy_tax, bases, offs, 31, 31)
        assert np.array_equal(t, et) and np.array_equal(h, eh) and np.array_equal(m, em)
        oracle.db_free(D)


def test_bns_python_surface(capi, oracle, genomes, tmp_path):
    from bonsai_b200 import bns
    s = bytes(genomes["phix"]).decode()
    assert np.array_equal(bns.from_str(s, 31), oracle.encode(s, 31, 31))
    assert np.array_equal(bns.from_str(s, 31, "", 50, False), oracle.encode(s, 31, 50, canon=False))
    p = tmp_path / "x.fa"
    p.write_text(">a\n%s\n>b desc\n%s\n%s\n" % (s[:300], s[300:400], s[400:650]))
    gaps = "1x2,0x28"
    lists = bns.seqlist(str(p), 31, gaps, 0, True)
    exp = [oracle.encode(x, 31, 0, bns.parse_spacing(gaps, 31), 0, True, po.API_PATH) for x in (s[:300], s[300:650])]
    assert len(lists) == 2 and all(np.array_equal(a, b) for a, b in zip(lists, exp))
    assert np.array_equal(bns.from_fasta(str(p), 31, unique=True), np.unique(np.concatenate(
        [oracle.encode(x, 31, 31, None, 0, True, po.API_PATH) for x in (s[:300], s[300:650])])))
    # seqdict (python/bns.cpp:175-199): the reference maps every record name to the k-mers of the whole file; per_record=True
    # gives each its own
    whole = np.concatenate([oracle.encode(x, 31, 31, None, 0, True, po.API_PATH) for x in (s[:300], s[300:650])])
    d = bns.seqdict(str(p), 31)
    assert list(d) == ["a", "b"] and all(np.array_equal(v, whole) for v in d.values())
    d = bns.seqdict(str(p), 31, per_record=True)
    assert np.array_equal(d["a"], oracle.encode(s[:300], 31, 31, None, 0, True, po.API_PATH))
    assert np.array_equal(d["b"], oracle.encode(s[300:650], 31, 31, None, 0, True, po.API_PATH))


@pytest.mark.parametrize("name", ["lex_k31_w31", "ent_k31_w50", "spaced_k31_c40"])
def test_device_build_matches_reference(capi, golden, genomes, name):
    """bonsai build on the GPU (encode -> insert-or-LCA-merge) reproduces the reference-built database bit for bit"""
    from bonsai_b200 import dbbuild
    spec = golden["dbs"][name]
    c, p = H.toy_tax_arrays()
    gs = [H.genome_records(genomes, gi) for gi in range(4)]
    with capi.Context(31, 31) as ctx:
        info = dbbuild.build_on_device(ctx, gs, H.GENOME_TAXIDS, c, p, spec["k"], spec["w"], spec["gaps"], spec["score"], spec["canon"])
        assert info["n_keys"] == spec["size"]
        k, v = ctx.table_dump()
        vals, cnts = np.unique(v, return_counts=True)
        assert {int(a): int(b) for a, b in zip(vals, cnts)} == {int(a): b for a, b in spec["hist"].items()}
        h = hashlib.md5()
        h.update(k.tobytes()); h.update(v.tobytes())
        assert h.hexdigest() == spec["md5"]
        # and the table is immediately usable for classification after switching the encoder back
        ctx.reconfigure(31, 31)
        gv, gf = ctx.lookup(k[::1000])
        assert gf.all() and np.array_equal(gv, v[::1000])
    # the host-merge builder agrees too (small slice: first 200 kb of each genome)
    small = [(b[:200_000].copy(), np.array([0, 200_000], np.uint64)) for b, _ in gs]
    k1, v1 = dbbuild.build_db(small, H.GENOME_TAXIDS, c, p, spec["k"], spec["w"], spec["gaps"], spec["score"], spec["canon"])
    with capi.Context(31, 31) as ctx:
        dbbuild.build_on_device(ctx, small, H.GENOME_TAXIDS, c, p, spec["k"], spec["w"], spec["gaps"], spec["score"], spec["canon"])
        k2, v2 = ctx.table_dump()
    assert np.array_equal(k1, k2) and np.array_equal(v1, v2)


def test_error_paths(capi):
    """every failure is an error code + message, never a crash"""
    with pytest.raises(capi.BnsError) as e:
        capi.Context(0)
    assert e.value.code == -1
    with pytest.raises(capi.BnsError):
        capi.Context(33)
    with pytest.raises(capi.BnsError) as e:
        capi.Context(31, 31, [300] + [0] * 29)             # comb beyond the supported span
    assert "comb" in str(e.value)
    with capi.Context(31) as ctx:
        b, o = po.pack_reads(["ACGT" * 40])
        with pytest.raises(capi.BnsError) as e:             # classify before any table
            ctx.classify(b, o)
        assert e.value.code == -4
        ctx.load_pairs(np.array([1, 2, 3], np.uint64), np.array([5, 6, 7], np.uint32))
        with pytest.raises(capi.BnsError) as e:             # table but no taxonomy
            ctx.classify(b, o)
        assert e.value.code == -4
        # a cycle in the taxonomy is rejected
        ctx.load_taxonomy(np.array([5, 6, 7], np.uint32), np.array([6, 7, 5], np.uint32))
        with pytest.raises(capi.BnsError) as e:
            ctx.classify(b, o)
        assert e.value.code == -5
        # lenient where the reference is UB: a value (7) whose parent is not a node, values missing from the taxonomy
        ctx.load_taxonomy(np.array([5, 6], np.uint32), np.array([1, 5], np.uint32))
        t, h, m = ctx.classify(b, o)
        assert t.tolist() == [0] and h.tolist() == [0] and m.tolist() == [130]
        with pytest.raises(capi.BnsError):                  # build_add_genome without build_begin
            ctx.build_add_genome(b, o, 5)


def test_many_distinct_taxa_and_big_taxonomy(capi, oracle):
    """records with dozens of distinct taxa on a 300 k-node, 2 000-deep taxonomy (iterative Euler tour, lca climbs)"""
    rng = np.random.default_rng(9)
    n = 300_000
    nodes = np.arange(1, n + 1, dtype=np.uint32) * 3 + 1            # taxids 4, 7, 10, ... plus the root 1
    nodes[0] = 1
    parent = np.zeros(n, np.uint32)
    parent[0] = 1
    # a 2 000-long spine, everything else hangs off random earlier nodes
    for i in range(1, 2000):
        parent[i] = nodes[i - 1]
    parent[2000:] = nodes[(rng.random(n - 2000) * np.arange(2000, n)).astype(np.int64)]
    T = oracle.tax_from_pairs(nodes, parent)
    # DB: random 31-mers of a random sequence, values drawn from 5 000 taxa (some deep on the spine)
    seq = "".join(rng.choice(list("ACGT"), 40_000))
    km = np.unique(oracle.encode(seq, 31, 31))
    pool = np.concatenate([nodes[1500:2000], rng.choice(nodes, 4500, replace=False)])
    vals = pool[rng.integers(0, pool.size, km.size)].astype(np.uint32)
    D = oracle.db_from_pairs(km, vals)
    reads = [seq[s:s + int(rng.integers(100, 280))] for s in rng.integers(0, len(seq) - 400, 300)]
    b, o = po.pack_reads(reads)
    with capi.Context(31) as ctx:
        ctx.load_pairs(km, vals)
        ctx.load_taxonomy(nodes, parent)
        t, h, m = ctx.classify(b, o)
    et, eh, em = oracle.classify(D, T, b, o, 31, 31)
    assert np.array_equal(t, et) and np.array_equal(h, eh) and np.array_equal(m, em)
    assert len(set(t.tolist())) > 50
    # more than 256 distinct taxa in one record (what shared memory holds): the second pass, lists in global memory
    long_read = seq[:3000]
    b2, o2 = po.pack_reads([long_read, seq[5000:5600], long_read[::-1]])
    vals2 = (np.arange(km.size) % 4000 + 1).astype(np.uint32) * 3 + 1
    D2 = oracle.db_from_pairs(km, vals2)
    with capi.Context(31) as ctx:
        ctx.load_pairs(km, vals2)
        ctx.load_taxonomy(nodes, parent)
        t, h, m = ctx.classify(b2, o2)
    et, eh, em = oracle.classify(D2, T, b2, o2, 31, 31)
    assert eh[0] > 2000 and np.array_equal(t, et) and np.array_equal(h, eh) and np.array_equal(m, em)


@pytest.mark.parametrize("mode", ["lex_canon", "lex_nocanon", "ent_canon_sat", "ent_canon_wrap", "ent_nocanon_sat",
                                  "path_ent_canon", "lex_canon_w33", "ent_canon_k21_w60"])
def test_classify_lean_windowed_vs_oracle(capi, oracle, dbcache, toy_tax, genomes, mode):
    """The lean kernel's windowed modes (shuffle-only sliding minima; counts, no hit list) against the oracle on reads
    that exercise its corners: reads shorter than the window (tail flush), empty reads, N (compaction / k-mer 0), runs of
    >= 32 T (the encoder.h:283 restart: deferred to the generic kernel) and multi-tile records (deferred as well)."""
    cfg = {
        "lex_canon":         dict(k=31, w=50, score=po.SCORE_LEX, canon=True, api=po.API_STRING, cast=po.CAST_SATURATE),
        "lex_nocanon":       dict(k=31, w=50, score=po.SCORE_LEX, canon=False, api=po.API_STRING, cast=po.CAST_SATURATE),
        "ent_canon_sat":     dict(k=31, w=50, score=po.SCORE_ENTROPY, canon=True, api=po.API_STRING, cast=po.CAST_SATURATE),
        "ent_canon_wrap":    dict(k=31, w=50, score=po.SCORE_ENTROPY, canon=True, api=po.API_STRING, cast=po.CAST_WRAP),
        "ent_nocanon_sat":   dict(k=31, w=50, score=po.SCORE_ENTROPY, canon=False, api=po.API_STRING, cast=po.CAST_SATURATE),
        "path_ent_canon":    dict(k=31, w=50, score=po.SCORE_ENTROPY, canon=True, api=po.API_PATH, cast=po.CAST_SATURATE),
        "lex_canon_w33":     dict(k=31, w=33, score=po.SCORE_LEX, canon=True, api=po.API_STRING, cast=po.CAST_SATURATE),
        "ent_canon_k21_w60": dict(k=21, w=60, score=po.SCORE_ENTROPY, canon=True, api=po.API_STRING, cast=po.CAST_SATURATE),
    }[mode]
    k, w = cfg["k"], cfg["w"]
    rng = np.random.default_rng(11)
    bases, offs, _ = H.make_reads(1500, seed=21, ragged=True)
    reads = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(offs.size - 1)]
    fixed, foffs, _ = H.make_reads(1500, seed=22)
    reads += [bytes(fixed[int(foffs[i]):int(foffs[i + 1])]) for i in range(foffs.size - 1)]
    g = genomes["bases"]
    for j in range(40):                                   # T runs of 30..70 inside genome reads; homopolymers; many N
        s = int(rng.integers(0, g.size - 400))
        r = bytearray(g[s:s + 150].tobytes())
        p, n = int(rng.integers(0, 80)), int(rng.integers(30, 71))
        r[p:p + n] = (b"T" if j % 4 else b"A") * n
        if j % 5 == 0:
            r[int(rng.integers(0, 150))] = ord("N")
        reads.append(bytes(r[:150]))
    for j in range(12):                                   # multi-tile records
        s = int(rng.integers(0, g.size - 3000))
        reads.append(bytes(g[s:s + int(rng.integers(160 + k, 2500))]))
    reads += [b"T" * 150, b"A" * 150, b"ACGT" * 40, b"N" * 150, b"", b"ACGTN" * 30]
    b, o = po.pack_reads(reads)
    dbname = "lex_k31_w31" if k == 31 else None
    if dbname:
        db = dbcache.get(dbname)
    else:                                                 # a k=21 DB of the first genome only
        db = oracle.db_new()
        oracle.db_add_genome(db, toy_tax, H.genome_records(genomes, 0), H.GENOME_TAXIDS[0], k, k)
    keys, vals, flags, nb, _ = oracle.db_arrays(db)
    c, p = H.toy_tax_arrays()
    with capi.Context(k, w, None, cfg["score"], cfg["canon"], cfg["api"], entropy_cast=cfg["cast"]) as ctx:
        ctx.load_table(keys, vals, flags, nb)
        ctx.load_taxonomy(c, p)
        got = ctx.classify(b, o)                          # counts, no hit list: the lean kernel
        got_taxon_only, _, _ = ctx.classify(b, o, want_counts=False)
        full = ctx.classify(b, o, want_taxa=True)         # the generic kernel (ordered hit lists)
        # the same reads as mate pairs (two sequences per record in the lean kernel; a deferred mate defers the pair)
        ne = (len(reads) // 2) * 2
        pb_, po_ = po.pack_reads(reads[:ne])
        got_pair = ctx.classify(pb_, po_, paired=True)
    exp_pair = oracle.classify(db, toy_tax, pb_, po_, k, w, None, cfg["score"], cfg["canon"], cfg["api"], cast_mode=cfg["cast"], paired=True)
    for name, a, bb in zip(("taxon", "nhit", "nmiss"), exp_pair, got_pair):
        bad = np.nonzero(a != bb)[0]
        assert bad.size == 0, "paired %s differs for %d records, first %d: oracle %d gpu %d" % (name, bad.size, bad[0], a[bad[0]], bb[bad[0]])
    exp = oracle.classify(db, toy_tax, b, o, k, w, None, cfg["score"], cfg["canon"], cfg["api"], cast_mode=cfg["cast"])
    for name, a, bb in zip(("taxon", "nhit", "nmiss"), exp, got):
        bad = np.nonzero(a != bb)[0]
        assert bad.size == 0, "%s differs for %d records, first %d (len %d): oracle %d gpu %d" % (
            name, bad.size, bad[0], len(reads[bad[0]]), a[bad[0]], bb[bad[0]])
    assert np.array_equal(got_taxon_only, exp[0])
    for a, bb in zip(exp, full[:3]):
        assert np.array_equal(a, bb)


def test_minimizer_layout_vs_oracle(capi, oracle, toy_tax, monkeypatch):
    """The opt-in minimizer-bucketed table layout (line by the k-mer's canonical 16-mer minimizer, structured remainder): same
    kh_get results, same dump, same classification as the oracle, through the lean kernel (fast warp-level encode), the
    generic kernel (hit lists) and the windowed modes (per-key encode); and a key set of another k falls back cleanly."""
    import torch
    from bonsai_b200 import workload as W
    monkeypatch.setenv("BNS_B200_LAYOUT", "minimizer")
    dev = torch.device("cuda", 0)
    n_keys = (1 << 19) + 777
    stream, d_keys, d_vals = W.make_stress_db(n_keys, seed=15, device=dev)
    d_bases, d_offs, from_db = W.make_stress_reads(stream, 12000, seed=16, device=dev)
    keys = d_keys.cpu().numpy().astype(np.uint64)
    vals = d_vals.cpu().numpy().astype(np.uint32)
    uk, ui = np.unique(keys, return_index=True)
    D = oracle.db_from_pairs(uk, vals[ui])
    # reads: the stress reads plus ragged / N / long ones
    rb, ro, _ = H.make_reads(1500, seed=23, ragged=True)
    reads = [bytes(rb[int(ro[i]):int(ro[i + 1])]) for i in range(ro.size - 1)]
    sb, so = d_bases.cpu().numpy(), d_offs.cpu().numpy().astype(np.uint64)
    reads += [bytes(sb[int(so[i]):int(so[i + 1])]) for i in range(so.size - 1)]
    st = np.frombuffer(b"ACGT", np.uint8)[stream[:20000].cpu().numpy()]      # the key stream, as bases: long all-hit records
    for j in range(8):
        s0 = 1000 * j
        reads.append(bytes(st[s0:s0 + 700 + 37 * j]))
    b, o = po.pack_reads(reads)
    c, p = H.toy_tax_arrays()
    with capi.Context(31, 31) as ctx:
        ctx.load_pairs_device(d_keys.data_ptr(), d_vals.data_ptr(), n_keys, W.STRESS_VALUES)
        ctx.load_taxonomy(c, p)
        info = ctx.table_info()
        assert info["layout"] == 1, "the minimizer layout was not taken"
        gv, gf = ctx.lookup(keys)
        assert gf.all() and np.array_equal(gv, vals)
        rng = np.random.default_rng(3)
        probe = rng.integers(0, 1 << 62, 200000, dtype=np.uint64)
        pv, pf = ctx.lookup(probe)
        present = np.isin(probe, uk)
        assert np.array_equal(pf, present)
        dk, dv = ctx.table_dump()
        order = np.argsort(dk)
        assert np.array_equal(dk[order], uk) and np.array_equal(dv[order], vals[ui])
        exp = oracle.classify(D, toy_tax, b, o, 31, 31, want_taxa=True)
        got = ctx.classify(b, o)                                  # lean kernel, fast encode
        for a, bb in zip(exp[:3], got):
            assert np.array_equal(a, bb)
        full = ctx.classify(b, o, want_taxa=True)                 # generic kernel
        for a, bb in zip(exp[:3], full[:3]):
            assert np.array_equal(a, bb)
        assert all(np.array_equal(a, bb) for a, bb in zip(exp[3], full[3]))
        # windowed minimizers on the reads against the same table (lean windowed kernel, per-key encode)
        ctx.reconfigure(31, 50, None, capi.SCORE_LEX, True, capi.API_STRING)
        expw = oracle.classify(D, toy_tax, b, o, 31, 50)
        gotw = ctx.classify(b, o)
        for a, bb in zip(expw, gotw):
            assert np.array_equal(a, bb)
        # a shorter k against the k=31 table: keys that are not 31-mers of the table cannot be there
        ctx.reconfigure(25, 25, None, capi.SCORE_LEX, True, capi.API_STRING)
        t25, h25, m25 = ctx.classify(b, o)
        e25 = oracle.classify(D, toy_tax, b, o, 25, 25)
        for a, bb in zip(e25, (t25, h25, m25)):
            assert np.array_equal(a, bb)
    # k = 21 is outside the layout's range: the hash layout is used
    with capi.Context(21, 21) as ctx:
        ctx.load_pairs(uk[:1000] & np.uint64((1 << 42) - 1), vals[ui][:1000])
        assert ctx.table_info()["layout"] == 0
    oracle.db_free(D)


def test_minimizer_layout_real_genomes_use_the_stash(capi, oracle, dbcache, toy_tax, reads2000, monkeypatch):
    """The 10.5 M canonical 31-mers of the four genomes in the minimizer layout: repeated 16-mers crowd single groups, the keys
    whose six-unit chain is full go to the stash (a small hash-layout table) instead of sending the whole table back to the
    hash layout. Every key is found with its value, absent keys are not, the dump is the key set, classification is the
    oracle's -- also on a replica made from the exported segments."""
    monkeypatch.setenv("BNS_B200_LAYOUT", "minimizer")
    db = dbcache.get("lex_k31_w31")
    keys, vals = oracle.db_pairs(db)
    c, p = H.toy_tax_arrays()
    bases, offs, _ = reads2000
    exp = oracle.classify(db, toy_tax, bases, offs, 31, 31, want_taxa=True)
    with capi.Context(31, 31) as ctx:
        ctx.load_pairs(keys, vals)
        ctx.load_taxonomy(c, p)
        info = ctx.table_info()
        assert info["layout"] == 1 and info["n_stash"] > 0 and info["n_keys"] == keys.size, info
        assert info["max_disp"] <= 6
        gv, gf = ctx.lookup(keys)
        assert gf.all() and np.array_equal(gv, vals)
        rng = np.random.default_rng(4)
        probe = rng.integers(0, 1 << 62, 300000, dtype=np.uint64)
        pv, pf = ctx.lookup(probe)
        assert np.array_equal(pf, np.isin(probe, keys))
        dk, dv = ctx.table_dump()
        assert np.array_equal(dk, keys) and np.array_equal(dv, vals)
        got = ctx.classify(bases, offs)
        assert all(np.array_equal(a, b) for a, b in zip(exp[:3], got))
        full = ctx.classify(bases, offs, want_taxa=True)
        assert all(np.array_equal(a, b) for a, b in zip(exp[3], full[3]))
        t, h, m, runs = ctx.classify_runs(bases, offs)
        assert np.array_equal(t, exp[0]) and all(np.array_equal(rn, _rle(x)) for rn, x in zip(runs, exp[3]))
        # a replica from the exported header + segments (the stash is the fifth segment)
        import torch
        with capi.Context(31, 31) as rep:
            rep.db_alloc_from_header(ctx.db_export_header())
            segs_a, segs_b = ctx.db_segments(), rep.db_segments()
            assert len(segs_a) == 5 and [n for _, n in segs_a] == [n for _, n in segs_b]
            for (pa, na), (pb, nb) in zip(segs_a, segs_b):
                torch.as_tensor(capi.DevMem(pb, nb), device="cuda").copy_(torch.as_tensor(capi.DevMem(pa, na), device="cuda"))
            torch.cuda.synchronize()
            rep.db_commit()
            assert rep.table_info()["n_stash"] == info["n_stash"]
            got = rep.classify(bases, offs)
            assert all(np.array_equal(a, b) for a, b in zip(exp[:3], got))


# ---- the call-by-call Encoder surface (BNS_API_ITER) ---------------------------------------------------------------
def test_iterator_surface_golden_and_fuzz(capi, oracle):
    """assign / has_next_kmer / next_minimizer / next_canonicalized_minimizer (encoder.h:201-206,594-628) as one unfiltered
    stream per sequence: against the vectors of the unmodified reference (tests/golden/golden_iter.json, saturating cast) and,
    batched and in both cast modes, against the oracle on seeded random sequences (N runs, T runs, multi-tile lengths)."""
    import json
    import os
    with open(os.path.join(H.GOLDEN, "golden_iter.json")) as f:
        gold = json.load(f)
    ctxs = {}

    def ctx_for(k, w, gaps, score, canon, cast):
        key = (k, w, tuple(gaps or ()), score, canon, cast)
        if key not in ctxs:
            ctxs[key] = capi.Context(k, w, gaps, score, bool(canon), capi.API_ITER, cast)
        return ctxs[key]
    for c in gold["cases"]:
        b, o = po.pack_reads([c["seq"].encode()])
        got = ctx_for(c["k"], c["w"], c["gaps"], c["score"], c["canon"], capi.CAST_SATURATE).encode_lists(b, o)[0]
        assert hx(got) == c["values"], {k: c[k] for k in ("k", "w", "gaps", "score", "canon", "seq")}
    rng = random.Random(4321)
    six = [0] * 30
    for i, g in ((2, 1), (7, 2), (11, 1), (15, 1), (20, 3), (25, 1)):
        six[i] = g
    for k, w, gaps in ((31, 0, None), (31, 50, None), (31, 0, six), (31, 70, six), (21, 33, None), (32, 45, None), (13, 200, None)):
        seqs = []
        for _ in range(60):
            L = rng.choice([0, 5, k, 64, 150, 151, 400, 1100])
            s = "".join(rng.choice(rng.choice(["ACGT", "ACGTN", "ACGTacgt", "T"])) for _ in range(L))
            if L > 60 and rng.random() < 0.3:
                p0 = rng.randrange(L - 40)
                s = s[:p0] + "T" * 40 + s[p0 + 40:]
            seqs.append(s.encode())
        b, o = po.pack_reads(seqs)
        for score in (0, 1):
            for canon in (0, 1):
                for cast, ocast in ((capi.CAST_SATURATE, po.CAST_SATURATE), (capi.CAST_WRAP, po.CAST_WRAP)):
                    if score == 0 and cast == capi.CAST_WRAP:
                        continue
                    got = ctx_for(k, w, gaps, score, canon, cast).encode_lists(b, o)
                    for sq, g in zip(seqs, got):
                        exp = oracle.encode(sq, k, w, gaps, score, canon, po.API_ITER, cast_mode=ocast)
                        assert np.array_equal(g, exp), dict(k=k, w=w, gaps=gaps, score=score, canon=canon, cast=cast, seq=sq)
    # an encode-only configuration: classify refuses it
    c0 = ctx_for(31, 0, None, 0, 1, capi.CAST_SATURATE)
    c0.load_pairs(np.array([1, 2, 3], np.uint64), np.array([11, 11, 12], np.uint32))
    tc, tp = H.toy_tax_arrays()
    c0.load_taxonomy(tc, tp)
    b, o = po.pack_reads([b"ACGT" * 40])
    with pytest.raises(capi.BnsError):
        c0.classify(b, o)
    for c in ctxs.values():
        c.close()


def test_cpp_encoder_surface(capi, oracle, tmp_path):
    """The C++ mirror of the reference's Encoder (include/bonsai_b200/bonsai.hpp): for_each(fn, str, l), the record overloads and
    assign / has_next_kmer / next_kmer / next_minimizer / next_canonicalized_minimizer, compiled into tests/host/encoder_api.cpp
    and linked against the library, against streams the oracle computes here."""
    import os
    import shutil
    import subprocess
    from bonsai_b200 import build
    here = os.path.dirname(os.path.abspath(__file__))
    libdir = os.path.dirname(build.build())
    exe = str(tmp_path / "encoder_api")
    r = subprocess.run([shutil.which("g++") or "/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, os.path.join(here, "host", "encoder_api.cpp"),
                        "-L" + libdir, "-lbonsai_b200", "-lz", "-lpthread", "-Wl,-rpath," + libdir], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    # the mirror picks the cast of THIS host's CPU (AVX-512 or not), as a -march=native build of the reference would
    cast = po.CAST_SATURATE if po.host_has_avx512() else po.CAST_WRAP
    BF = (1 << 64) - 1
    rng = random.Random(99)
    gapped = [1, 2] + [0] * 28
    lines = []
    for k, w, gaps in ((31, 31, None), (31, 50, None), (31, 31, gapped), (31, 60, gapped), (16, 24, None)):
        cc, ww = oracle.spacer(k, w, gaps)[:2]
        for trial in range(6):
            L = rng.choice([0, 20, 34, 150, 300])
            seq = "".join(rng.choice("ACGT" if trial % 3 else "ACGTN") for _ in range(L))
            if trial == 0:
                seq = "ACATGCTAGCATGCTGACTGACTGATCGATCGTA"                                     # test/encoding.cpp:17
            npos = max(0, len(seq) - cc + 1)
            for score in (0, 1):
                for canon in (0, 1):
                    for kind in range(5):
                        if kind == 0:
                            exp = [int(x) for x in oracle.encode(seq, k, w, gaps, score, canon, po.API_STRING, cast_mode=cast)]
                        elif kind == 1:
                            exp = [int(x) for x in oracle.encode(seq, k, w, gaps, score, canon, po.API_PATH, cast_mode=cast)]
                        elif kind == 2:                                                         # kmer(pos) for every position
                            exp = [int(x) for x in oracle.encode(seq, k, 0, gaps, score, 0, po.API_ITER, cast_mode=cast)]
                        else:                                                                   # W - 1 calls before the first full window
                            v = [int(x) for x in oracle.encode(seq, k, w, gaps, score, int(kind == 4), po.API_ITER, cast_mode=cast)]
                            exp = [BF] * min(npos, ww - cc) + v
                        if kind >= 2:
                            assert len(exp) == npos
                        lines.append("%d %d %d %d %d %s %s %d %s" % (kind, k, w, score, canon, ",".join(map(str, gaps)) if gaps else "-", seq or "-", len(exp),
                                                                     " ".join("%x" % x for x in exp)))
    cases = tmp_path / "cases.txt"
    cases.write_text("\n".join(lines) + "\n")
    r = subprocess.run([exe, str(cases)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout and ("%d cases, 0 failures" % len(lines)) in r.stdout, r.stdout[-3000:]
