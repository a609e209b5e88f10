"""`bns`-style convenience functions over the batched-encode ABI (the reference's pybind11 module python/bns.cpp:52-146
is SURVEY 8(f)-4, a second consumer of bns_b200_encode_batch). Same argument names and meaning:

    from_str(str, k=31, spacing="", w=0, canon=True)            -> uint64 array, Encoder<>::for_each(fn, str, len)
    from_fasta(path, k=31, spacing="", w=0, canon=True, unique=False) -> uint64 array over every record (record overloads)
    seqlist(path, k=31, spacing="", w=0, canon=True, unique=False)    -> list of uint64 arrays, one per record
    seqdict(path, k=31, spacing="", w=0)                              -> {record name: uint64 array}   (python/bns.cpp:175-199)
    repack(vallist, ngenomes)                                         -> float32 [pairs, len(vallist)]  (python/bns.cpp:130-150)
The `_r` variants of the reference (from_fasta_r, seqdict_r, seqlist(rolling=True)) hash k-mers with RollingHasher, a cyclic
polynomial hash from the `rollinghash` dependency: not an Encoder and not on the classify path; they are not
offered here rather than offered on the CPU.
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


def _named_records(path):
    """(name, sequence) per record; name = header up to the first white space, as kseq"""
    op = gzip.open if path.endswith(".gz") else open
    name, seq = None, []
    with op(path, "rt") as f:
        fastq = None
        lines = iter(f)
        for line in lines:
            line = line.rstrip("\r\n")
            if fastq is None and line[:1] in ">@":
                fastq = line[0] == "@"
            if line[:1] == ">" or (fastq and line[:1] == "@"):
                if name is not None and not fastq:
                    yield name, "".join(seq)
                name, seq = (line[1:].split() or [""])[0], []
                if fastq:
                    s = next(lines).rstrip("\r\n")
                    next(lines)
                    next(lines)
                    yield name, s
                    name = None
            elif not fastq:
                seq.append(line)
    if name is not None:
        yield name, "".join(seq)


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


def seqdict(path, k=31, spacing="", w=0, per_record=False, device=-1):
    """python/bns.cpp:175-199. The reference hands `path.data()` -- the FILE NAME, not the record -- to Encoder::for_each, which
    takes it for the path overload: every record name maps to the k-mers of the WHOLE file. That is what per_record=False
    returns (one array shared by all names); per_record=True gives each name the k-mers of its own record."""
    recs = list(_named_records(path))
    if per_record:
        with capi.Context(k, w, parse_spacing(spacing, k), capi.SCORE_LEX, True, capi.API_PATH, device=device) as ctx:
            b, o = _pack([s for _, s in recs])
            out = ctx.encode_lists(b, o)
        return {n: a for (n, _), a in zip(recs, out)}
    whole = from_fasta(path, k, spacing, w, True, False, device)
    return {n: whole for n, _ in recs}


def repack(vallist, ngenomes):
    """python/bns.cpp:130-150: `vallist` holds one array of ngenomes*(ngenomes-1)/2 upper-triangle values per k; the result
    has the reference's shape (pairs, len(vallist)) with list entry `kind` written to the flat positions
    kind * pairs + [0, pairs) -- the layout its loop produces."""
    npairs = (ngenomes * (ngenomes - 1)) >> 1
    nks = len(vallist)
    ret = np.zeros(npairs * nks, np.float32)
    for kind, vals in enumerate(vallist):
        v = np.asarray(vals, np.float32).reshape(-1)
        ret[kind * npairs:(kind + 1) * npairs] = v[:npairs]
    return ret.reshape(npairs, nks)
