"""Synthetic workloads of BASELINE.json: the 4-genome fixture, the toy taxonomy and the seeded read generator
(SURVEY App. C.1 / section 8(d)), generated on the GPU with torch so a 10 M-read batch costs milliseconds.

The genome fixture is tests/golden/genomes4.npz (2-bit packed, derived from the reference's test genomes by
tests/golden/make_golden.py). If it is missing, four seeded random genomes of the same sizes stand in and the
returned dict says so (`synthetic_genomes: True`).
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "genomes4.npz")

TOY_TAX = [(1, 1), (2, 1), (10, 2), (11, 10), (12, 10), (13, 10), (20, 2)]
GENOME_TAXIDS = [11, 12, 13, 20]
SPACED_GAPS = [0] * 30
for _i, _g in ((2, 1), (7, 2), (11, 1), (15, 1), (20, 3), (25, 1)):
    SPACED_GAPS[_i] = _g

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def toy_tax_arrays():
    return (np.array([a for a, _ in TOY_TAX], np.uint32), np.array([b for _, b in TOY_TAX], np.uint32))


def _unpack2bit(packed, n):
    b = np.asarray(packed, dtype=np.uint8)
    codes = np.empty(b.size * 4, dtype=np.uint8)
    for j in range(4):
        codes[j::4] = (b >> (2 * j)) & 3
    return _ACGT[codes[:n]]


def load_genomes(path=FIXTURE):
    """-> dict(bases ASCII uint8 (all contigs concatenated), contig_off uint64[n+1], contig_genome uint8[n])"""
    if os.path.exists(path):
        z = np.load(path)
        lens = z["contig_len"].astype(np.uint64)
        off = np.zeros(lens.size + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        return dict(bases=_unpack2bit(z["packed"], int(off[-1])), contig_off=off, contig_genome=z["contig_genome"],
                    synthetic_genomes=False)
    rng = np.random.default_rng(2024)
    sizes = [2_684_267, 2_449_987, 2_478_074, 5_192_569]
    off = np.zeros(5, np.uint64)
    off[1:] = np.cumsum(sizes)
    return dict(bases=_ACGT[rng.integers(0, 4, int(off[-1]))], contig_off=off,
                contig_genome=np.arange(4, dtype=np.uint8), synthetic_genomes=True)


def genome_records(g, gi):
    idx = np.nonzero(g["contig_genome"] == gi)[0]
    lo, hi = int(idx[0]), int(idx[-1]) + 1
    base0 = g["contig_off"][lo]
    return (g["bases"][int(base0):int(g["contig_off"][hi])], (g["contig_off"][lo:hi + 1] - base0).astype(np.uint64))


def make_reads_torch(g, n, seed, device, L=150, frac_random=0.10, sub_rate=0.01, frac_n=0.05, frac_rc=0.5,
                     chunk=1 << 20):
    """Config-2 reads on the GPU: 90 % sampled uniformly from the 4 genomes (contigs >= L), 1 % substitutions,
    5 % of reads get one N, 50 % reverse-complemented; 10 % uniform random ACGT. Fixed length L.
    -> (bases uint8 [n*L] on `device`, offsets int64 [n+1] on `device`)"""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    genome = torch.from_numpy(g["bases"]).to(device)
    coff = g["contig_off"].astype(np.int64)
    clen = coff[1:] - coff[:-1]
    ok = np.nonzero(clen >= L)[0]
    cg = g["contig_genome"][ok]
    # per genome: table of eligible contigs, padded
    per = [ok[cg == x] for x in range(4)]
    width = max(len(p) for p in per)
    tab = np.zeros((4, width), np.int64)
    cnt = np.zeros(4, np.int64)
    for x, p in enumerate(per):
        tab[x, :len(p)] = p
        cnt[x] = len(p)
    tab_t, cnt_t = torch.from_numpy(tab).to(device), torch.from_numpy(cnt).to(device)
    coff_t, clen_t = torch.from_numpy(coff).to(device), torch.from_numpy(clen).to(device)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    out = torch.empty(n * L, dtype=torch.uint8, device=device)
    ar = torch.arange(L, device=device)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        u = torch.rand((7, m), generator=gen, device=device)
        gi = (u[0] * 4).long().clamp_(max=3)
        ci = tab_t[gi, (u[1] * cnt_t[gi]).long().clamp_(max=width - 1) % cnt_t[gi]]
        start = coff_t[ci] + (u[2] * (clen_t[ci] - L + 1)).long()
        rd = genome[start[:, None] + ar[None, :]]
        rnd = acgt[torch.randint(0, 4, (m, L), generator=gen, device=device)]
        sub = torch.rand((m, L), generator=gen, device=device) < sub_rate
        take = sub | (u[3] < frac_random)[:, None]
        rd = torch.where(take, rnd, rd)
        has_n = u[4] < frac_n
        npos = (u[5] * L).long().clamp_(max=L - 1)
        rows = torch.nonzero(has_n).squeeze(1)
        rd[rows, npos[rows]] = ord("N")
        do_rc = u[6] < frac_rc
        rc = comp[rd.flip(1).long()]
        rd = torch.where(do_rc[:, None], rc, rd)
        out[s * L:(s + m) * L] = rd.reshape(-1)
    offs = torch.arange(n + 1, device=device, dtype=torch.int64) * L
    return out, offs


# ------------------------------------------------------------------------------------------------------------------
# BASELINE config 5: HBM-bound lookup stress. DB = canonical 31-mers of a seeded random base stream (n_keys of them,
# values uniform over the toy taxids), reads = 50 % windows of that stream (every k-mer hits), 50 % random (every
# k-mer misses), half of all reads reverse-complemented. Everything is produced on the device with torch.
# ------------------------------------------------------------------------------------------------------------------
STRESS_VALUES = np.array([10, 11, 12, 13, 20, 2], np.uint32)


def _kmers_torch(codes, k, canonical=True):
    """codes: uint8 tensor of 2-bit codes (length n + k - 1) -> int64 tensor of n k-mers (two's complement view of u64)"""
    import torch
    n = codes.numel() - k + 1
    c64 = codes.to(torch.int64)
    f = torch.zeros(n, dtype=torch.int64, device=codes.device)
    for j in range(k):
        f = (f << 2) | c64[j:j + n]
    if not canonical:
        return f
    r = torch.zeros(n, dtype=torch.int64, device=codes.device)
    for j in range(k):
        r = r | ((3 - c64[j:j + n]) << (2 * j))
    # k <= 31: both are < 2^62, signed compare == unsigned compare
    return torch.minimum(f, r)


def make_stress_db(n_keys, seed, device, k=31, chunk=1 << 25):
    """-> (stream codes uint8 [n_keys + k - 1] on device, keys int64 [n_keys], vals int32 [n_keys])"""
    import torch
    assert k <= 31
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    stream = torch.randint(0, 4, (n_keys + k - 1,), generator=gen, device=device, dtype=torch.uint8)
    keys = torch.empty(n_keys, dtype=torch.int64, device=device)
    for s in range(0, n_keys, chunk):
        m = min(chunk, n_keys - s)
        keys[s:s + m] = _kmers_torch(stream[s:s + m + k - 1], k)
    vals_tab = torch.from_numpy(STRESS_VALUES.astype(np.int32)).to(device)
    vals = vals_tab[(keys % len(STRESS_VALUES)).long()].contiguous()
    return stream, keys, vals


def make_stress_reads(stream, n, seed, device, L=150, frac_db=0.5, frac_rc=0.5, chunk=1 << 20):
    """-> (bases uint8 ASCII [n*L], offsets int64 [n+1], from_db bool [n])"""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    out = torch.empty(n * L, dtype=torch.uint8, device=device)
    from_db = torch.empty(n, dtype=torch.bool, device=device)
    ar = torch.arange(L, device=device)
    nmax = stream.numel() - L
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        u = torch.rand((3, m), generator=gen, device=device, dtype=torch.float64)
        start = (u[0] * nmax).long()
        codes = stream[start[:, None] + ar[None, :]]
        rnd = torch.randint(0, 4, (m, L), generator=gen, device=device, dtype=torch.uint8)
        db = u[1] < frac_db
        codes = torch.where(db[:, None], codes, rnd)
        rc = (3 - codes).flip(1)
        codes = torch.where((u[2] < frac_rc)[:, None], rc, codes)
        out[s * L:(s + m) * L] = acgt[codes.long()].reshape(-1)
        from_db[s:s + m] = db
    offs = torch.arange(n + 1, device=device, dtype=torch.int64) * L
    return out, offs, from_db
