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
