"""CPU, world_size 2 over gloo: the N > 1 host logic -- read sharding (order-preserving, no overlap) and the
load-time database replication protocol (header, segments, commit) -- with the CPU oracle standing in for the
per-rank classify step."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeCtx:
    """Implements the db_* surface of capi.Context over host memory."""

    def __init__(self, segs=None):
        self.segs = [np.frombuffer(bytes(s), np.uint8).copy() for s in (segs or [])]
        self.committed = segs is not None

    def db_export_header(self):
        h = np.zeros(16, np.uint64)
        h[0] = 0x42304e53424e5331
        for i, s in enumerate(self.segs):
            h[1 + i] = s.size
        return h

    def db_alloc_from_header(self, words):
        assert int(words[0]) == 0x42304e53424e5331
        self.segs = [np.zeros(int(words[1 + i]), np.uint8) for i in range(4)]

    def db_segments(self):
        return [(s.ctypes.data, s.size) for s in self.segs]

    def db_commit(self):
        self.committed = True


def _host_view(ptr, nbytes):
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
    return torch.frombuffer(buf, dtype=torch.uint8)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bonsai_b200 import sharding
    from oracle import pyoracle as po
    o = po.load_oracle()
    # --- DB replication: rank 0 owns the DB (keys/vals of a small k-mer set + the toy taxonomy)
    bases, offs, _ = H.make_reads(1201, seed=4)
    rng = np.random.default_rng(0)
    if rank == 0:
        km = np.unique(np.concatenate([o.encode(bytes(bases[int(offs[i]):int(offs[i + 1])]), 31, 31) for i in range(400)]))[::2]
        vals = rng.choice(np.array([2, 10, 11, 12, 13, 20], np.uint32), km.size)
        c, p = H.toy_tax_arrays()
        ctx = FakeCtx([km.tobytes(), vals.tobytes(), c.tobytes(), p.tobytes()])
    else:
        ctx = FakeCtx()
    moved = sharding.replicate_db(ctx, dist, rank, root=0, as_tensor=_host_view, header_device="cpu")
    assert ctx.committed and moved == sum(s.size for s in ctx.segs)
    km = ctx.segs[0].view(np.uint64); vals = ctx.segs[1].view(np.uint32)
    c = ctx.segs[2].view(np.uint32); p = ctx.segs[3].view(np.uint32)
    # --- read sharding: classify this rank's shard, gather in rank order
    D, T = o.db_from_pairs(km, vals), o.tax_from_pairs(c, p)
    (lo, hi), (r0, r1), (b0, b1) = sharding.shard_reads(offs, rank, world)
    t, h, m = o.classify(D, T, bases[b0:b1], offs[r0:r1 + 1] - offs[r0], 31, 31)
    out = [None] * world
    dist.all_gather_object(out, (lo, hi, t, h, m))
    if rank == 0:
        et, eh, em = o.classify(D, T, bases, offs, 31, 31)
        assert [x[0] for x in out] == [0] + [x[1] for x in out[:-1]] and out[-1][1] == offs.size - 1
        assert np.array_equal(np.concatenate([x[2] for x in out]), et)
        assert np.array_equal(np.concatenate([x[3] for x in out]), eh)
        assert np.array_equal(np.concatenate([x[4] for x in out]), em)
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_shard_ranges():
    from bonsai_b200 import sharding
    for n in (0, 1, 7, 10, 1001):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    offs = np.arange(0, 11 * 150, 150, dtype=np.uint64)
    assert sharding.shard_reads(offs, 1, 2, mates=2) == ((3, 5), (6, 10), (900, 1500))
