"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the two CPU checkers.

* ``Oracle``  : oracle/_build/libbns_oracle.so, the plain-C restatement (bns_oracle.c). Built on demand
                with gcc; available everywhere.
* ``Ref``     : oracle/_ref/libbns_ref_v{3,4}.so, the UNMODIFIED reference headers behind
                oracle/ref_driver.cpp. Built only where /root/reference exists; the prebuilt
                binaries travel to the GPU box. ``load_ref()`` returns None when absent.

Both expose the same methods so a test can run the same call against either.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCORE_LEX, SCORE_ENTROPY = 0, 1
API_STRING, API_PATH, API_ITER = 0, 1, 2
CAST_SATURATE, CAST_WRAP = 0, 1

_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(ty)


def _gaps(k, gaps):
    if gaps is None:
        return None
    g = np.ascontiguousarray(np.asarray(gaps, dtype=np.uint16))
    assert g.size == k - 1, "gap vector must have k-1 entries"
    return g


def pack_reads(seqs):
    """list of str/bytes -> (uint8 bases, uint64 offsets[n+1])"""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return bases, offs


def host_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            return "avx512f" in f.read()
    except OSError:
        return False


class _Base:
    """Common surface; subclasses bind prefix `bo_` (oracle) or `bref_` (reference)."""

    prefix = ""
    has_cast_arg = False

    def __init__(self, lib):
        self.lib = lib
        p = self.prefix
        f = lambda n: getattr(lib, p + n)
        for n in ("lex_score", "wang64"):
            f(n).restype = C.c_uint64
            f(n).argtypes = [C.c_uint64]
        for n in ("rc", "canonical"):
            f(n).restype = C.c_uint64
            f(n).argtypes = [C.c_uint64, C.c_uint]
        f("parse_spacing").restype = C.c_int
        f("parse_spacing").argtypes = [C.c_char_p, C.c_uint, _u16p, C.c_int]
        f("encode").restype = C.c_int64
        cast = [C.c_int] if self.has_cast_arg else []
        f("encode").argtypes = [C.c_uint, C.c_uint, _u16p, C.c_int, C.c_int, C.c_int] + cast + \
            [C.c_char_p, C.c_uint64, _u64p, C.c_uint64]
        f("tax_load").restype = C.c_void_p
        f("tax_load").argtypes = [C.c_char_p]
        f("tax_from_pairs").restype = C.c_void_p
        f("tax_from_pairs").argtypes = [_u32p, _u32p, C.c_uint64]
        f("tax_size").restype = C.c_uint64
        f("tax_size").argtypes = [C.c_void_p]
        f("tax_pairs").restype = C.c_uint64
        f("tax_pairs").argtypes = [C.c_void_p, _u32p, _u32p, C.c_uint64]
        f("tax_free").argtypes = [C.c_void_p]
        f("lca").restype = C.c_uint32
        f("lca").argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        f("resolve").restype = C.c_uint32
        f("resolve").argtypes = [C.c_void_p, _u32p, _u16p, C.c_uint32]
        f("db_from_pairs").restype = C.c_void_p
        f("db_from_pairs").argtypes = [_u64p, _u32p, C.c_uint64]
        f("db_arrays").argtypes = [C.c_void_p, C.POINTER(_u64p), C.POINTER(_u32p), C.POINTER(_u32p), _u64p, _u64p]
        f("db_get").restype = C.c_int
        f("db_get").argtypes = [C.c_void_p, C.c_uint64, _u32p]
        f("db_free").argtypes = [C.c_void_p]
        f("classify").argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, _u16p, C.c_int, C.c_int, C.c_int] + cast + \
            [C.c_void_p, _u64p, C.c_uint64, C.c_int, _u32p, _u32p, _u32p, _u32p, _u64p, C.c_int]
        f("classify_text").restype = C.c_void_p
        f("classify_text").argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, _u16p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, _u64p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_uint64,
                                       C.c_int, _u64p, _u64p, _u64p]
        f("free").argtypes = [C.c_void_p]

    def _f(self, n):
        return getattr(self.lib, self.prefix + n)

    # scalar
    def lex_score(self, x): return self._f("lex_score")(x)
    def wang64(self, x): return self._f("wang64")(x)
    def rc(self, x, k): return self._f("rc")(x, k)
    def canonical(self, x, k): return self._f("canonical")(x, k)

    def parse_spacing(self, s, k):
        out = np.zeros(4 * k + 64, dtype=np.uint16)
        n = self._f("parse_spacing")(s.encode() if s is not None else None, k, _ptr(out, _u16p), out.size)
        return out[:n].copy()

    def encode(self, seq, k, w, gaps=None, score=SCORE_LEX, canon=True, api=API_STRING, cast_mode=None):
        if isinstance(seq, str):
            seq = seq.encode()
        seq = bytes(seq)
        g = _gaps(k, gaps)
        cap = len(seq) + 8
        out = np.zeros(cap, dtype=np.uint64)
        cast = [self.cast_mode if cast_mode is None else cast_mode] if self.has_cast_arg else []
        n = self._f("encode")(k, w, _ptr(g, _u16p), score, int(canon), api, *cast, seq, len(seq), _ptr(out, _u64p), cap)
        assert 0 <= n <= cap
        return out[:n].copy()

    # taxonomy
    def tax_from_pairs(self, child, parent):
        c = np.ascontiguousarray(child, dtype=np.uint32)
        p = np.ascontiguousarray(parent, dtype=np.uint32)
        return self._f("tax_from_pairs")(_ptr(c, _u32p), _ptr(p, _u32p), c.size)

    def tax_load(self, path):
        return self._f("tax_load")(path.encode())

    def tax_pairs(self, t):
        n = self._f("tax_size")(t)
        c = np.zeros(n, np.uint32)
        p = np.zeros(n, np.uint32)
        m = self._f("tax_pairs")(t, _ptr(c, _u32p), _ptr(p, _u32p), n)
        assert m == n
        o = np.argsort(c)
        return c[o], p[o]

    def tax_free(self, t): self._f("tax_free")(t)
    def lca(self, t, a, b): return self._f("lca")(t, a, b)

    def resolve(self, t, taxa, counts):
        a = np.ascontiguousarray(taxa, dtype=np.uint32)
        c = np.ascontiguousarray(counts, dtype=np.uint16)
        return self._f("resolve")(t, _ptr(a, _u32p), _ptr(c, _u16p), a.size)

    # database
    def db_from_pairs(self, keys, vals):
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        v = np.ascontiguousarray(vals, dtype=np.uint32)
        return self._f("db_from_pairs")(_ptr(k, _u64p), _ptr(v, _u32p), k.size)

    def db_arrays(self, db):
        """Borrowed numpy views of the raw khash arrays: keys, vals, flags, n_buckets, size."""
        kp, vp, fp = _u64p(), _u32p(), _u32p()
        nb, sz = C.c_uint64(), C.c_uint64()
        self._f("db_arrays")(db, C.byref(kp), C.byref(vp), C.byref(fp), C.byref(nb), C.byref(sz))
        n = nb.value
        if n == 0:
            return np.zeros(0, np.uint64), np.zeros(0, np.uint32), np.zeros(1, np.uint32), 0, 0
        keys = np.ctypeslib.as_array(kp, shape=(n,))
        vals = np.ctypeslib.as_array(vp, shape=(n,))
        flags = np.ctypeslib.as_array(fp, shape=(max(1, n >> 4),))
        return keys, vals, flags, n, sz.value

    def db_pairs(self, db):
        """Occupied (key, val) pairs sorted by key."""
        keys, vals, flags, n, _ = self.db_arrays(db)
        if n == 0:
            return np.zeros(0, np.uint64), np.zeros(0, np.uint32)
        idx = np.arange(n, dtype=np.uint64)
        fl = (flags[(idx >> np.uint64(4)).astype(np.int64)] >> ((idx & np.uint64(15)) << np.uint64(1)).astype(np.uint32)) & 3
        occ = fl == 0
        k, v = keys[occ].copy(), vals[occ].copy()
        o = np.argsort(k)
        return k[o], v[o]

    def db_get(self, db, key):
        v = C.c_uint32()
        return v.value if self._f("db_get")(db, key, C.byref(v)) else None

    def db_free(self, db): self._f("db_free")(db)

    def classify(self, db, tax, bases, offsets, k, w, gaps=None, score=SCORE_LEX, canon=True, api=API_STRING,
                 cast_mode=None, paired=False, want_taxa=False, nthreads=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        inc = 2 if paired else 1
        nrec = n // inc
        g = _gaps(k, gaps)
        taxon = np.zeros(nrec, np.uint32)
        nhit = np.zeros(nrec, np.uint32)
        nmiss = np.zeros(nrec, np.uint32)
        taxa = toffs = None
        if want_taxa:
            lens = (offsets[1:] - offsets[:-1]).astype(np.uint64)
            per = lens.reshape(nrec, inc).sum(axis=1) if nrec else np.zeros(0, np.uint64)
            toffs = np.zeros(nrec + 1, np.uint64)
            toffs[1:] = np.cumsum(per + np.uint64(2))
            taxa = np.zeros(int(toffs[-1]), np.uint32)
        cast = [self.cast_mode if cast_mode is None else cast_mode] if self.has_cast_arg else []
        self._f("classify")(db, tax, k, w, _ptr(g, _u16p), score, int(canon), api, *cast,
                            bases.ctypes.data, _ptr(offsets, _u64p), n, int(paired),
                            _ptr(taxon, _u32p), _ptr(nhit, _u32p), _ptr(nmiss, _u32p),
                            _ptr(taxa, _u32p), _ptr(toffs, _u64p), nthreads)
        if want_taxa:
            lists = [taxa[int(toffs[i]):int(toffs[i]) + int(nhit[i])].copy() for i in range(nrec)]
            return taxon, nhit, nmiss, lists
        return taxon, nhit, nmiss

    def classify_text(self, db, tax, bases, offsets, names, k, w, gaps=None, canon=True, emit_all=True,
                      emit_fastq=False, emit_kraken=True, quals=None, paired=False):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        g = _gaps(k, gaps)
        nm = (C.c_char_p * n)(*[x.encode() for x in names])
        ql = None
        if quals is not None:
            ql = (C.c_char_p * n)(*[None if q is None else q.encode() for q in quals])
        ln, nc, nu = C.c_uint64(), C.c_uint64(), C.c_uint64()
        p = self._f("classify_text")(db, tax, k, w, _ptr(g, _u16p), int(canon), int(emit_all), int(emit_fastq),
                                     int(emit_kraken), bases.ctypes.data, _ptr(offsets, _u64p), nm, ql, n,
                                     int(paired), C.byref(ln), C.byref(nc), C.byref(nu))
        txt = C.string_at(p, ln.value)
        self._f("free")(p)
        return txt, nc.value, nu.value


class Oracle(_Base):
    prefix = "bo_"
    has_cast_arg = True

    def __init__(self, lib):
        super().__init__(lib)
        lib.bo_cast_u64.restype = C.c_uint64
        lib.bo_cast_u64.argtypes = [C.c_double, C.c_int]
        lib.bo_host_cast_mode.restype = C.c_int
        lib.bo_spacer_info.argtypes = [C.c_uint, C.c_uint, _u16p, _u32p, _u32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.bo_db_new.restype = C.c_void_p
        lib.bo_db_from_arrays.restype = C.c_void_p
        lib.bo_db_from_arrays.argtypes = [_u64p, _u32p, _u32p, C.c_uint64]
        lib.bo_db_add_genome.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, _u16p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, _u64p, C.c_uint64, C.c_uint32]
        lib.bo_db_probe_count.restype = C.c_uint64
        lib.bo_db_probe_count.argtypes = [C.c_void_p, C.c_uint64]
        self.cast_mode = CAST_SATURATE
        self.kind = "port"

    def cast_u64(self, x, mode): return self.lib.bo_cast_u64(x, mode)
    def host_cast_mode(self): return self.lib.bo_host_cast_mode()

    def spacer(self, k, w, gaps=None):
        g = _gaps(k, gaps)
        c, wo = C.c_uint32(), C.c_uint32()
        us, uw = C.c_int(), C.c_int()
        self.lib.bo_spacer_info(k, w, _ptr(g, _u16p), C.byref(c), C.byref(wo), C.byref(us), C.byref(uw))
        return c.value, wo.value, bool(us.value), bool(uw.value)

    def db_new(self): return self.lib.bo_db_new()

    def db_from_arrays(self, keys, vals, flags, n_buckets):
        return self.lib.bo_db_from_arrays(_ptr(keys, _u64p), _ptr(vals, _u32p), _ptr(flags, _u32p), n_buckets)

    def db_add_genome(self, db, tax, records, taxid, k, w, gaps=None, score=SCORE_LEX, canon=True, cast_mode=None):
        """records: list of sequences (contigs) of one genome, or a (bases, offsets) pair."""
        bases, offs = records if isinstance(records, tuple) else pack_reads(records)
        g = _gaps(k, gaps)
        self.lib.bo_db_add_genome(db, tax, k, w, _ptr(g, _u16p), score, int(canon),
                                  self.cast_mode if cast_mode is None else cast_mode,
                                  bases.ctypes.data, _ptr(offs, _u64p), offs.size - 1, taxid)

    def db_probe_count(self, db, key): return self.lib.bo_db_probe_count(db, key)


class Ref(_Base):
    prefix = "bref_"
    has_cast_arg = False

    def __init__(self, lib, variant):
        super().__init__(lib)
        lib.bref_cast_saturates.restype = C.c_int
        lib.bref_build_info.restype = C.c_char_p
        lib.bref_spacer.argtypes = [C.c_uint, C.c_uint, _u16p, _u32p, _u32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.bref_db_build.restype = C.c_void_p
        lib.bref_db_build.argtypes = [C.c_void_p, C.c_uint, C.c_uint, _u16p, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_char_p), _u32p]
        self.variant = variant
        self.cast_mode = CAST_SATURATE if lib.bref_cast_saturates() else CAST_WRAP
        self.kind = "reference"

    def build_info(self): return self.lib.bref_build_info().decode()

    def spacer(self, k, w, gaps=None):
        g = _gaps(k, gaps)
        c, wo = C.c_uint32(), C.c_uint32()
        us, uw = C.c_int(), C.c_int()
        self.lib.bref_spacer(k, w, _ptr(g, _u16p), C.byref(c), C.byref(wo), C.byref(us), C.byref(uw))
        return c.value, wo.value, bool(us.value), bool(uw.value)

    def db_build(self, tax, paths, taxids, k, w, gaps=None, score=SCORE_LEX, canon=True):
        g = _gaps(k, gaps)
        pa = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        t = np.ascontiguousarray(taxids, dtype=np.uint32)
        return self.lib.bref_db_build(tax, k, w, _ptr(g, _u16p), score, int(canon), len(paths), pa, _ptr(t, _u32p))


_oracle = None
_refs = {}


def build_oracle(force=False):
    so = os.path.join(HERE, "_build", "libbns_oracle.so")
    src = os.path.join(HERE, "bns_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


def build_ref():
    """(Re)build oracle/_ref from the reference sources where they lie; no-op when they are absent."""
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def load_oracle():
    global _oracle
    if _oracle is None:
        _oracle = Oracle(C.CDLL(build_oracle()))
    return _oracle


def load_ref(variant=None):
    """variant 'v4' (AVX-512, saturating cast) / 'v3' (wrapping cast); default: best the host can run."""
    if variant is None:
        variant = "v4" if host_has_avx512() else "v3"
    if variant == "v4" and not host_has_avx512():
        return None
    if variant not in _refs:
        so = os.path.join(HERE, "_ref", "libbns_ref_%s.so" % variant)
        _refs[variant] = Ref(C.CDLL(so), variant) if os.path.exists(so) else None
    return _refs[variant]
