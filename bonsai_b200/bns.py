"""`bns`-style convenience functions over the batched-encode ABI (the reference's pybind11 module python/bns.cpp:52-146
is SURVEY 8(f)-4, a second consumer of bns_b200_encode_batch). Same argument names and meaning:

    from_str(str, k=31, spacing="", w=0, canon=True)            -> uint64 array, Encoder<>::for_each(fn, str, len)
    from_fasta(path, k=31, spacing="", w=0, canon=True, unique=False) -> uint64 array over every record (record overloads)
    seqlist(path, k=31, spacing="", w=0, canon=True, unique=False)    -> list of uint64 arrays, one per record
"""
import gzip

import numpy as np

from . import capi


def parse_spacing(spacing, k):
    """"1x3,0x5" -> gap list of k-1 entries (include/bonsai/spacer.h:29-47)"""
    if not spacing:
        return [0] * (k - 1)
    out = []
    for item in spacing.split(","):
        if "x" in item:
            g, n = item.split("x", 1)
            out += [int(g)] * max(int(n), 1)
        else:
            out.append(int(item))
    return out


def _records(path):
    op = gzip.open if path.endswith(".gz") else open
    seq = []
    with op(path, "rt") as f:
        fastq = None
        lines = iter(f)
        for line in lines:
            line = line.rstrip("\r\n")
            if fastq is None and line[:1] in ">@":
                fastq = line[0] == "@"
            if line[:1] == ">" or (fastq and line[:1] == "@"):
                if seq:
                    yield "".join(seq)
                seq = []
                if fastq:
                    s = next(lines).rstrip("\r\n")
                    next(lines)
                    next(lines)
                    yield s
            elif not fastq:
                seq.append(line)
    if seq:
        yield "".join(seq)


def _pack(seqs):
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs), np.uint8).copy() if bs else np.zeros(0, np.uint8), offs


def from_str(s, k=31, spacing="", w=0, canon=True, device=-1):
    with capi.Context(k, w, parse_spacing(spacing, k), capi.SCORE_LEX, canon, capi.API_STRING, device=device) as ctx:
        b, o = _pack([s])
        return ctx.encode_lists(b, o)[0]


def seqlist(path, k=31, spacing="", w=0, canon=True, unique=False, device=-1):
    with capi.Context(k, w, parse_spacing(spacing, k), capi.SCORE_LEX, canon, capi.API_PATH, device=device) as ctx:
        b, o = _pack(list(_records(path)))
        out = ctx.encode_lists(b, o)
    return [np.unique(x) for x in out] if unique else out


def from_fasta(path, k=31, spacing="", w=0, canon=True, unique=False, device=-1):
    parts = seqlist(path, k, spacing, w, canon, False, device)
    allk = np.concatenate(parts) if parts else np.zeros(0, np.uint64)
    return np.unique(allk) if unique else allk
