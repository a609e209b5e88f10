"""ctypes binding of libbonsai_b200.so (include/bonsai_b200.h). Used by the tests and bench.py; the C++ mirror
of the reference surface lives in include/bonsai_b200/bonsai.hpp.

There is no CPU fallback: if the library is missing or no CUDA device is usable, loading / opening raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BNS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbonsai_b200.so")

SCORE_LEX, SCORE_ENTROPY = 0, 1
API_STRING, API_PATH, API_ITER = 0, 1, 2
CAST_SATURATE, CAST_WRAP = 0, 1
MAX_K = 32

# every symbol include/bonsai_b200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "bns_b200_version", "bns_b200_strerror", "bns_b200_last_error", "bns_b200_open", "bns_b200_close",
    "bns_b200_geometry", "bns_b200_encode_bound", "bns_b200_load_table", "bns_b200_load_pairs",
    "bns_b200_load_pairs_device", "bns_b200_table_info_get", "bns_b200_lookup_batch", "bns_b200_lookup_sectors", "bns_b200_load_taxonomy",
    "bns_b200_load_taxonomy_file", "bns_b200_resolve_batch", "bns_b200_build_begin", "bns_b200_build_add_genome",
    "bns_b200_build_finish", "bns_b200_table_dump", "bns_b200_reconfigure", "bns_b200_db_export_header",
    "bns_b200_db_alloc_from_header", "bns_b200_db_segments", "bns_b200_db_commit", "bns_b200_encode_batch",
    "bns_b200_classify_batch", "bns_b200_classify_batch_ex", "bns_b200_classify_batch_runs", "bns_b200_classify_device", "bns_b200_sync", "bns_b200_stats_get",
    "bns_b200_stats_reset", "bns_b200_host_alloc", "bns_b200_host_free", "bns_b200_bench_gather",
    "bns_b200_open_multi", "bns_b200_replicate", "bns_b200_close_multi", "bns_b200_device_status", "bns_b200_classify_device_runs",
    "bns_b200_set_host_pack_threads", "bns_b200_host_pack_threads",
]


class Config(C.Structure):
    _fields_ = [("k", C.c_uint32), ("w", C.c_uint32), ("gaps", C.c_uint16 * MAX_K), ("score", C.c_uint32),
                ("canonicalize", C.c_uint32), ("api", C.c_uint32), ("entropy_cast", C.c_uint32),
                ("device", C.c_int32), ("n_gpus", C.c_uint32), ("host_pack_threads", C.c_uint32), ("reserved", C.c_uint32 * 5)]


class TableInfo(C.Structure):
    _fields_ = [("n_keys", C.c_uint64), ("n_buckets", C.c_uint64), ("bytes", C.c_uint64),
                ("bucket_bits", C.c_uint32), ("val_bits", C.c_uint32), ("n_values", C.c_uint32), ("max_disp", C.c_uint32),
                ("n_displaced", C.c_uint64), ("n_overflowed", C.c_uint64), ("layout", C.c_uint32), ("disp_bits", C.c_uint32),
                ("n_stash", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("n_classified", C.c_uint64), ("n_unclassified", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("reads_processed", C.c_uint64), ("bases_processed", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("last_kernel_ms", C.c_double), ("kernel_ms_total", C.c_double)]


class DbHeader(C.Structure):
    _fields_ = [("words", C.c_uint64 * 16)]


class BnsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bonsai_b200 error %d: %s" % (code, msg))
        self.code = code


class DevMem:
    """A raw device range as a __cuda_array_interface__ object: torch.as_tensor(DevMem(p, n), device="cuda") gives a
    uint8 view the host plumbing can hand to torch.distributed.broadcast (the load-time DB replication)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


_lib = None


def load_library(path=None):
    """dlopen the in-tree library and declare the prototypes. Raises OSError if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    lib = C.CDLL(path or LIB_PATH)
    vp, u64p, u32p, u16p, u8p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)
    lib.bns_b200_version.restype = C.c_char_p
    lib.bns_b200_strerror.restype = C.c_char_p
    lib.bns_b200_strerror.argtypes = [C.c_int]
    lib.bns_b200_last_error.restype = C.c_char_p
    lib.bns_b200_last_error.argtypes = [vp]
    lib.bns_b200_open.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.bns_b200_close.argtypes = [vp]
    lib.bns_b200_set_host_pack_threads.argtypes = [vp, C.c_uint32]
    lib.bns_b200_host_pack_threads.argtypes = [vp]
    lib.bns_b200_close.restype = None
    lib.bns_b200_geometry.argtypes = [vp, u32p, u32p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.bns_b200_encode_bound.restype = C.c_uint64
    lib.bns_b200_encode_bound.argtypes = [vp, C.c_uint64]
    lib.bns_b200_load_table.argtypes = [vp, vp, vp, vp, C.c_uint64]
    lib.bns_b200_load_pairs.argtypes = [vp, vp, vp, C.c_uint64]
    lib.bns_b200_load_pairs_device.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint32]
    lib.bns_b200_table_info_get.argtypes = [vp, C.POINTER(TableInfo)]
    lib.bns_b200_lookup_batch.argtypes = [vp, vp, C.c_uint64, vp, vp]
    lib.bns_b200_lookup_sectors.argtypes = [vp, vp, C.c_uint64, u64p]
    lib.bns_b200_build_begin.argtypes = [vp, C.c_uint64, vp, C.c_uint32]
    lib.bns_b200_build_add_genome.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint32]
    lib.bns_b200_build_finish.argtypes = [vp]
    lib.bns_b200_table_dump.argtypes = [vp, vp, vp, C.c_uint64, u64p]
    lib.bns_b200_reconfigure.argtypes = [vp, C.POINTER(Config)]
    lib.bns_b200_load_taxonomy.argtypes = [vp, vp, vp, C.c_uint64]
    lib.bns_b200_load_taxonomy_file.argtypes = [vp, C.c_char_p]
    lib.bns_b200_resolve_batch.argtypes = [vp, vp, vp, vp, C.c_uint64, vp]
    lib.bns_b200_db_export_header.argtypes = [vp, C.POINTER(DbHeader)]
    lib.bns_b200_db_alloc_from_header.argtypes = [vp, C.POINTER(DbHeader)]
    lib.bns_b200_db_segments.argtypes = [vp, C.POINTER(vp), u64p, C.c_int, C.POINTER(C.c_int)]
    lib.bns_b200_db_commit.argtypes = [vp]
    lib.bns_b200_encode_batch.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, vp]
    lib.bns_b200_classify_batch.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, vp, vp, vp, vp, vp]
    lib.bns_b200_classify_batch_ex.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.bns_b200_classify_batch_runs.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, vp, vp, vp, vp, vp, C.c_uint64, vp, vp, C.POINTER(C.c_uint64)]
    lib.bns_b200_classify_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.bns_b200_sync.argtypes = [vp]
    lib.bns_b200_device_status.argtypes = [vp]
    lib.bns_b200_classify_device_runs.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, vp, vp, vp, vp, C.c_uint64, vp, vp, vp, vp]
    lib.bns_b200_stats_get.argtypes = [vp, C.POINTER(Stats)]
    lib.bns_b200_stats_reset.argtypes = [vp]
    lib.bns_b200_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.bns_b200_host_free.argtypes = [vp]
    lib.bns_b200_bench_gather.argtypes = [vp, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]
    lib.bns_b200_open_multi.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_int), C.POINTER(vp), C.POINTER(C.c_int)]
    lib.bns_b200_replicate.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    lib.bns_b200_close_multi.argtypes = [C.POINTER(vp), C.c_int]
    lib.bns_b200_close_multi.restype = None
    if path is None:
        _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data


class Context:
    """One bns_b200 context = one ClassifierGeneric + its per-worker Encoder copy on one GPU."""

    @staticmethod
    def _config(k, w, gaps, score, canonicalize, api, entropy_cast, device, host_pack_threads=None):
        cfg = Config()
        cfg.k, cfg.w, cfg.score, cfg.canonicalize, cfg.api = k, w, score, int(bool(canonicalize)), api
        cfg.entropy_cast, cfg.device = entropy_cast, device
        # None: the library's default (this process's share of the hardware threads); 0: off; n: that many worker threads
        cfg.host_pack_threads = 0 if host_pack_threads is None else (0xffffffff if host_pack_threads <= 0 else int(host_pack_threads))
        if gaps is not None:
            assert len(gaps) == k - 1, "gap vector must have k-1 entries"
            for i, g in enumerate(gaps):
                cfg.gaps[i] = int(g)
        return cfg

    def __init__(self, k, w=0, gaps=None, score=SCORE_LEX, canonicalize=True, api=API_STRING,
                 entropy_cast=CAST_SATURATE, device=-1, host_pack_threads=None):
        self.lib = load_library()
        cfg = self._config(k, w, gaps, score, canonicalize, api, entropy_cast, device, host_pack_threads)
        h = C.c_void_p()
        rc = self.lib.bns_b200_open(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise BnsError(rc, self.lib.bns_b200_last_error(None).decode())
        self.h = h
        self.k = k

    @classmethod
    def _adopt(cls, handle, k):
        self = cls.__new__(cls)
        self.lib, self.h, self.k = load_library(), handle, k
        return self

    def _ck(self, rc):
        if rc != 0:
            raise BnsError(rc, self.lib.bns_b200_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.bns_b200_close(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def geometry(self):
        c, w = C.c_uint32(), C.c_uint32()
        us, uw, cn = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.bns_b200_geometry(self.h, C.byref(c), C.byref(w), C.byref(us), C.byref(uw), C.byref(cn)))
        return dict(c=c.value, w=w.value, unspaced=bool(us.value), unwindowed=bool(uw.value), canon=bool(cn.value))

    def encode_bound(self, n):
        return self.lib.bns_b200_encode_bound(self.h, n)

    # ---- database ----
    def load_table(self, keys, vals, flags, n_buckets):
        keys = np.ascontiguousarray(keys, np.uint64)
        vals = np.ascontiguousarray(vals, np.uint32)
        flags = np.ascontiguousarray(flags, np.uint32)
        self._ck(self.lib.bns_b200_load_table(self.h, _p(keys), _p(vals), _p(flags), n_buckets))

    def load_pairs(self, keys, vals):
        keys = np.ascontiguousarray(keys, np.uint64)
        vals = np.ascontiguousarray(vals, np.uint32)
        assert keys.size == vals.size
        self._ck(self.lib.bns_b200_load_pairs(self.h, _p(keys), _p(vals), keys.size))

    def load_pairs_device(self, d_keys_ptr, d_vals_ptr, n, values):
        values = np.ascontiguousarray(values, np.uint32)
        self._ck(self.lib.bns_b200_load_pairs_device(self.h, d_keys_ptr, d_vals_ptr, n, _p(values), values.size))

    def table_info(self):
        ti = TableInfo()
        self._ck(self.lib.bns_b200_table_info_get(self.h, C.byref(ti)))
        return {f: getattr(ti, f) for f, _ in TableInfo._fields_}

    def lookup(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        vals = np.zeros(keys.size, np.uint32)
        found = np.zeros(keys.size, np.uint8)
        self._ck(self.lib.bns_b200_lookup_batch(self.h, _p(keys), keys.size, _p(vals), _p(found)))
        return vals, found.astype(bool)

    def lookup_sectors(self, keys):
        """total 32-byte buckets touched by probing `keys` (p-bar = result / len(keys))"""
        keys = np.ascontiguousarray(keys, np.uint64)
        out = C.c_uint64()
        self._ck(self.lib.bns_b200_lookup_sectors(self.h, _p(keys), keys.size, C.byref(out)))
        return out.value

    # ---- database construction on the device ----
    def reconfigure(self, k, w=0, gaps=None, score=SCORE_LEX, canonicalize=True, api=API_STRING, entropy_cast=CAST_SATURATE):
        cfg = self._config(k, w, gaps, score, canonicalize, api, entropy_cast, -1)
        self._ck(self.lib.bns_b200_reconfigure(self.h, C.byref(cfg)))
        self.k = k

    def build_begin(self, max_kmers, taxids):
        t = np.ascontiguousarray(taxids, np.uint32)
        self._ck(self.lib.bns_b200_build_begin(self.h, int(max_kmers), _p(t), t.size))

    def build_add_genome(self, bases, offsets, taxid):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        self._ck(self.lib.bns_b200_build_add_genome(self.h, _p(bases), _p(offsets), offsets.size - 1, int(taxid)))

    def build_finish(self):
        self._ck(self.lib.bns_b200_build_finish(self.h))

    def table_dump(self):
        """-> (keys sorted uint64, vals uint32)"""
        n = C.c_uint64()
        self._ck(self.lib.bns_b200_table_dump(self.h, None, None, 0, C.byref(n)))
        keys = np.zeros(n.value, np.uint64)
        vals = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(self.lib.bns_b200_table_dump(self.h, _p(keys), _p(vals), n.value, C.byref(n)))
        o = np.argsort(keys)
        return keys[o], vals[o]

    # ---- taxonomy ----
    def load_taxonomy(self, child, parent):
        child = np.ascontiguousarray(child, np.uint32)
        parent = np.ascontiguousarray(parent, np.uint32)
        self._ck(self.lib.bns_b200_load_taxonomy(self.h, _p(child), _p(parent), child.size))

    def load_taxonomy_file(self, path):
        self._ck(self.lib.bns_b200_load_taxonomy_file(self.h, path.encode()))

    def resolve(self, lists):
        """lists: iterable of [(taxid, count), ...] -> uint32 taxon per list"""
        lists = [list(l) for l in lists]
        offs = np.zeros(len(lists) + 1, np.uint64)
        offs[1:] = np.cumsum([len(l) for l in lists])
        taxa = np.array([t for l in lists for t, _ in l], np.uint32)
        cnts = np.array([c for l in lists for _, c in l], np.uint16)
        out = np.zeros(len(lists), np.uint32)
        self._ck(self.lib.bns_b200_resolve_batch(self.h, _p(taxa), _p(cnts), _p(offs), len(lists), _p(out)))
        return out

    # ---- replication ----
    def db_export_header(self):
        h = DbHeader()
        self._ck(self.lib.bns_b200_db_export_header(self.h, C.byref(h)))
        return np.array(list(h.words), np.uint64)

    def db_alloc_from_header(self, words):
        h = DbHeader()
        for i, wd in enumerate(words):
            h.words[i] = int(wd)
        self._ck(self.lib.bns_b200_db_alloc_from_header(self.h, C.byref(h)))

    def db_segments(self):
        ptrs = (C.c_void_p * 8)()
        nbytes = (C.c_uint64 * 8)()
        n = C.c_int()
        self._ck(self.lib.bns_b200_db_segments(self.h, ptrs, nbytes, 8, C.byref(n)))
        return [(ptrs[i] or 0, nbytes[i]) for i in range(n.value)]

    def db_commit(self):
        self._ck(self.lib.bns_b200_db_commit(self.h))

    # ---- encode / classify ----
    def encode(self, bases, offsets):
        """-> (kmers uint64 flat, out_offsets uint64[n+1], counts uint32[n]); sequence r's k-mers are
        kmers[out_offsets[r] : out_offsets[r] + counts[r]]"""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = offsets.size - 1
        lens = offsets[1:] - offsets[:-1]
        c = self.geometry()["c"]
        bound = np.where(lens >= c, lens - np.uint64(c) + np.uint64(1), np.uint64(0)).astype(np.uint64)
        out_offs = np.zeros(n + 1, np.uint64)
        out_offs[1:] = np.cumsum(bound)
        kmers = np.zeros(int(out_offs[-1]) + 1, np.uint64)
        counts = np.zeros(n, np.uint32)
        self._ck(self.lib.bns_b200_encode_batch(self.h, _p(bases), _p(offsets), n, _p(kmers), _p(out_offs), _p(counts)))
        return kmers, out_offs, counts

    def encode_lists(self, bases, offsets):
        kmers, oo, cnt = self.encode(bases, offsets)
        return [kmers[int(oo[i]):int(oo[i]) + int(cnt[i])] for i in range(cnt.size)]

    def classify(self, bases, offsets, paired=False, want_counts=True, want_taxa=False):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = offsets.size - 1
        inc = 2 if paired else 1
        nrec = n // inc
        taxon = np.zeros(nrec, np.uint32)
        nhit = np.zeros(nrec, np.uint32) if want_counts else None
        nmiss = np.zeros(nrec, np.uint32) if want_counts else None
        taxa = toffs = None
        if want_taxa:
            lens = (offsets[1:] - offsets[:-1])[: nrec * inc].reshape(nrec, inc).sum(axis=1)
            toffs = np.zeros(nrec + 1, np.uint64)
            toffs[1:] = np.cumsum(lens + np.uint64(2))
            taxa = np.zeros(int(toffs[-1]) + 1, np.uint32)
        self._ck(self.lib.bns_b200_classify_batch(self.h, _p(bases), _p(offsets), n, int(paired), _p(taxon), _p(nhit),
                                                  _p(nmiss), _p(taxa), _p(toffs)))
        if want_taxa:
            lists = [taxa[int(toffs[i]):int(toffs[i]) + int(nhit[i])].copy() for i in range(nrec)]
            return taxon, nhit, nmiss, lists
        return taxon, nhit, nmiss

    def classify_runs(self, bases, offsets, paired=False, cap=None):
        """bns_b200_classify_batch_runs: (taxon, nhit, nmiss, [per record array of (taxid, run length) rows])"""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = offsets.size - 1
        inc = 2 if paired else 1
        nrec = n // inc
        taxon = np.zeros(nrec, np.uint32); nhit = np.zeros(nrec, np.uint32); nmiss = np.zeros(nrec, np.uint32)
        if cap is None:
            cap = int((offsets[nrec * inc] - offsets[0])) + 2 * nrec + (1 << 21)
        runs = np.zeros(max(cap, 1), np.uint64)
        pos = np.zeros(nrec, np.uint64); nruns = np.zeros(nrec, np.uint32)
        total = C.c_uint64(0)
        self._ck(self.lib.bns_b200_classify_batch_runs(self.h, _p(bases), _p(offsets), n, int(paired), _p(taxon), _p(nhit),
                                                       _p(nmiss), None, _p(runs), cap, _p(pos), _p(nruns), C.byref(total)))
        out = []
        for r in range(nrec):
            w = runs[int(pos[r]):int(pos[r]) + int(nruns[r])]
            out.append(np.stack([(w >> np.uint64(32)).astype(np.uint32), (w & np.uint64(0xffffffff)).astype(np.uint32)], axis=1))
        assert int(total.value) >= int(nruns.sum())          # the entries used: runs plus the stretches a warp left unused
        return taxon, nhit, nmiss, out

    def set_host_pack_threads(self, n):
        """worker threads the host-buffer classify calls pack bases with (0: every chunk crosses PCIe as ASCII)"""
        self._ck(self.lib.bns_b200_set_host_pack_threads(self.h, int(n)))

    def host_pack_threads(self):
        return int(self.lib.bns_b200_host_pack_threads(self.h))

    def classify_into(self, bases_ptr, offsets_ptr, n_reads, taxon_ptr, nhit_ptr=None, nmiss_ptr=None, paired=False):
        """bns_b200_classify_batch on raw HOST pointers (e.g. pinned buffers): blocks until outputs are written."""
        self._ck(self.lib.bns_b200_classify_batch(self.h, bases_ptr, offsets_ptr, n_reads, int(paired), taxon_ptr,
                                                  nhit_ptr, nmiss_ptr, None, None))

    def classify_device(self, d_bases, d_offsets, n_reads, d_taxon, d_nhit=0, d_nmiss=0, paired=False, stream=0,
                        d_taxa=0, d_taxa_offsets=0):
        """All arguments are raw device pointers (ints); asynchronous on `stream`."""
        self._ck(self.lib.bns_b200_classify_device(self.h, d_bases, d_offsets, n_reads, int(paired), d_taxon, d_nhit or None,
                                                   d_nmiss or None, d_taxa or None, d_taxa_offsets or None, stream or None))

    def classify_device_runs(self, d_bases, d_offsets, n_reads, d_taxon, d_nhit, d_runs, runs_cap, d_run_pos, d_nruns, d_total, d_nmiss=0,
                             paired=False, stream=0):
        """bns_b200_classify_device_runs: raw device pointers (ints); asynchronous on `stream`."""
        self._ck(self.lib.bns_b200_classify_device_runs(self.h, d_bases, d_offsets, n_reads, int(paired), d_taxon, d_nhit, d_nmiss or None,
                                                        d_runs, runs_cap, d_run_pos, d_nruns, d_total, stream or None))

    def sync(self):
        self._ck(self.lib.bns_b200_sync(self.h))

    def device_status(self):
        """raises what the asynchronous device-pointer calls latched (bns_b200_device_status)"""
        self._ck(self.lib.bns_b200_device_status(self.h))

    def stats(self):
        s = Stats()
        self._ck(self.lib.bns_b200_stats_get(self.h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in Stats._fields_}

    def stats_reset(self):
        self._ck(self.lib.bns_b200_stats_reset(self.h))

    def bench_gather(self, n_loads, seed=1):
        ms = C.c_double()
        self._ck(self.lib.bns_b200_bench_gather(self.h, n_loads, seed, C.byref(ms)))
        return ms.value


def open_multi(n_gpus, k, w=0, gaps=None, score=SCORE_LEX, canonicalize=True, api=API_STRING, entropy_cast=CAST_SATURATE, devices=None):
    """bns_b200_open_multi: one Context per GPU of this process (n_gpus = 0: every visible device)."""
    lib = load_library()
    cfg = Context._config(k, w, gaps, score, canonicalize, api, entropy_cast, -1)
    hs = (C.c_void_p * 64)()
    n = C.c_int()
    devs = None
    if devices is not None:
        devs = (C.c_int * len(devices))(*devices)
        n_gpus = len(devices)
    rc = lib.bns_b200_open_multi(C.byref(cfg), n_gpus, devs, hs, C.byref(n))
    if rc != 0:
        raise BnsError(rc, lib.bns_b200_last_error(None).decode())
    return [Context._adopt(C.c_void_p(hs[i]), k) for i in range(n.value)]


def replicate(ctxs, root=0):
    """bns_b200_replicate: the database of ctxs[root] into every other context (NCCL broadcast inside the library)."""
    lib = load_library()
    hs = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rc = lib.bns_b200_replicate(hs, len(ctxs), root)
    if rc != 0:
        raise BnsError(rc, lib.bns_b200_last_error(ctxs[root].h).decode())
