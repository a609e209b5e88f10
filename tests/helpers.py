"""Shared test fixtures: the 4-genome set, the toy taxonomy, seeded read generator, k-mer digests.

The genome fixture (tests/golden/genomes4.npz) is DERIVED data: the four archaeal assemblies the
reference ships as test/GCF_*.fna.gz plus phiX (test/phix.fa), 2-bit packed by
tests/golden/make_golden.py (all five are pure upper-case ACGT). Nothing here reads /root/reference.
"""
import gzip
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

# toy taxonomy (child, parent) and genome -> taxid in fixture order
TOY_TAX = [(1, 1), (2, 1), (10, 2), (11, 10), (12, 10), (13, 10), (20, 2)]
GENOME_TAXIDS = [11, 12, 13, 20]
# BASELINE.json configs[3]: k=31 with 6 non-zero gaps, comb 40
SPACED_GAPS = [0] * 30
for _i, _g in ((2, 1), (7, 2), (11, 1), (15, 1), (20, 3), (25, 1)):
    SPACED_GAPS[_i] = _g

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def unpack2bit(packed, n):
    """uint8 packed (4 bases/byte, first base in the low bits) -> ASCII uint8[n]"""
    b = np.asarray(packed, dtype=np.uint8)
    codes = np.empty(b.size * 4, dtype=np.uint8)
    for j in range(4):
        codes[j::4] = (b >> (2 * j)) & 3
    return _ACGT[codes[:n]]


def pack2bit(ascii_u8):
    lutv = np.full(256, 255, np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lutv[ch] = i
    codes = lutv[np.asarray(ascii_u8, dtype=np.uint8)]
    assert (codes != 255).all(), "fixture genomes are pure ACGT"
    pad = (-codes.size) % 4
    codes = np.concatenate([codes, np.zeros(pad, np.uint8)])
    return (codes[0::4] | (codes[1::4] << 2) | (codes[2::4] << 4) | (codes[3::4] << 6)).astype(np.uint8)


_genomes = None


def load_genomes():
    """-> dict(bases=uint8 ASCII of all contigs concatenated, contig_off=uint64[n+1],
              contig_genome=uint8[n] (0..3), phix=uint8 ASCII)"""
    global _genomes
    if _genomes is None:
        z = np.load(os.path.join(GOLDEN, "genomes4.npz"))
        lens = z["contig_len"].astype(np.uint64)
        off = np.zeros(lens.size + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        _genomes = dict(bases=unpack2bit(z["packed"], int(off[-1])), contig_off=off,
                        contig_genome=z["contig_genome"], phix=unpack2bit(z["phix_packed"], int(z["phix_len"])))
    return _genomes


def genome_records(g, gi):
    """(bases, offsets) of the contigs of genome gi, as a contiguous slice of the fixture."""
    idx = np.nonzero(g["contig_genome"] == gi)[0]
    lo, hi = int(idx[0]), int(idx[-1]) + 1
    assert (idx == np.arange(lo, hi)).all()
    base0 = g["contig_off"][lo]
    offs = (g["contig_off"][lo:hi + 1] - base0).astype(np.uint64)
    return g["bases"][int(base0):int(g["contig_off"][hi])], offs


def toy_tax_arrays():
    c = np.array([a for a, _ in TOY_TAX], np.uint32)
    p = np.array([b for _, b in TOY_TAX], np.uint32)
    return c, p


_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def make_reads(n, seed, L=150, genomes=None, frac_random=0.10, sub_rate=0.01, frac_n=0.05, frac_rc=0.5,
               ragged=False):
    """Seeded synthetic reads after SURVEY App. C.1 / 8(d) config 2: 90 % sampled from the 4 genomes
    (contigs >= L), 1 % substitutions, 5 % of reads get one N, 50 % reverse-complemented; 10 % uniform
    random ACGT. Returns (bases uint8, offsets uint64[n+1], origin int8[n] = genome index or -1).
    ragged=True draws lengths in [0, 2L] to exercise short / empty reads."""
    g = genomes or load_genomes()
    rng = np.random.default_rng(seed)
    coff = g["contig_off"]
    clen = (coff[1:] - coff[:-1]).astype(np.int64)
    lens = np.full(n, L, np.int64) if not ragged else rng.integers(0, 2 * L + 1, n)
    ok = np.nonzero(clen >= max(int(lens.max()) if n else 1, 1))[0]
    is_rand = rng.random(n) < frac_random
    # genome uniformly, then contig uniformly within the genome, then offset uniformly
    gi = rng.integers(0, 4, n)
    by_g = [ok[g["contig_genome"][ok] == x] for x in range(4)]
    ci = np.empty(n, np.int64)
    for x in range(4):
        m = gi == x
        ci[m] = by_g[x][rng.integers(0, by_g[x].size, int(m.sum()))]
    start = (coff[ci].astype(np.int64) + (rng.random(n) * (clen[ci] - lens + 1)).astype(np.int64))
    offs = np.zeros(n + 1, np.uint64)
    offs[1:] = np.cumsum(lens)
    total = int(offs[-1])
    rid = np.repeat(np.arange(n), lens)
    pos = np.arange(total, dtype=np.int64) - offs[:-1].astype(np.int64)[rid]
    bases = g["bases"][start[rid] + pos].copy()
    rnd = _ACGT[rng.integers(0, 4, total)]
    sub = rng.random(total) < sub_rate
    take_rnd = is_rand[rid] | sub
    bases[take_rnd] = rnd[take_rnd]
    # one N in frac_n of the reads
    has_n = (rng.random(n) < frac_n) & (lens > 0)
    npos = (rng.random(n) * lens).astype(np.int64)
    bases[(offs[:-1].astype(np.int64) + npos)[has_n]] = ord("N")
    # reverse complement
    do_rc = rng.random(n) < frac_rc
    rcm = do_rc[rid]
    src = offs[:-1].astype(np.int64)[rid] + (lens[rid] - 1 - pos)
    out = bases.copy()
    out[rcm] = _COMP[bases[src[rcm]]]
    origin = np.where(is_rand, -1, gi).astype(np.int8)
    return out, offs, origin


def digest(kmers):
    """n, xor of (km * 0x9E3779B97F4A7C15 + index), sum -- all mod 2^64."""
    km = np.asarray(kmers, dtype=np.uint64)
    n = km.size
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        x = np.bitwise_xor.reduce(km * np.uint64(0x9E3779B97F4A7C15) + idx) if n else np.uint64(0)
        s = km.sum(dtype=np.uint64) if n else np.uint64(0)
    return int(n), int(x), int(s)


def bgzf_bytes(data, block=0xff00):
    """`data` as bgzip writes it: independent gzip members with a "BC" extra subfield (block size - 1), then an empty block."""
    import struct
    import zlib
    out = bytearray()
    for x in list(range(0, len(data), block)) + [None]:
        chunk = b"" if x is None else data[x:x + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25)
        out += comp + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return bytes(out)


def write_fastq(path, names, seqs, quals=None):
    if str(path).endswith(".bgz"):
        txt = "".join(">%s extra\n%s\n" % (n, s) if quals is None else "@%s extra\n%s\n+\n%s\n" % (n, s, quals[i])
                      for i, (n, s) in enumerate(zip(names, seqs)))
        with open(path, "wb") as f:
            f.write(bgzf_bytes(txt.encode(), block=3000))
        return
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "wt") as f:
        for i, (n, s) in enumerate(zip(names, seqs)):
            if quals is None:
                f.write(">%s extra\n%s\n" % (n, s))
            else:
                f.write("@%s extra\n%s\n+\n%s\n" % (n, s, quals[i]))
