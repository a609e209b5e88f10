"""GPU (-m gpu), needs >= 2 devices (skipped otherwise): reads shard over the GPUs of one box, the database is replicated once
(SURVEY 8e). Three ways in:
  * bns_b200_open_multi + bns_b200_replicate (NCCL inside the library, one process): every replica must return rank 0's answers;
  * `bonsai classify --gpus 2`: byte-identical text to --gpus 1;
  * one process per GPU over torch.distributed (what bench.py does): a real NCCL broadcast of the four segments, checked against
    rank 0 and against the CPU oracle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ndev() < 2, reason="needs two CUDA devices")


def _small_db(oracle, genomes, tax):
    dbo = oracle.db_new()
    for gi, taxid in enumerate(H.GENOME_TAXIDS):
        b, _ = H.genome_records(genomes, gi)
        oracle.db_add_genome(dbo, tax, (b[:300_000].copy(), np.array([0, 300_000], np.uint64)), taxid, 31, 31)
    return dbo


@needs2
@pytest.mark.parametrize("how", ["nccl", "p2p"])
def test_open_multi_replicate_in_library(oracle, genomes, toy_tax, monkeypatch, how):
    from bonsai_b200 import capi
    if how == "p2p":
        monkeypatch.setenv("BNS_B200_REPLICATE", "p2p")           # peer copies instead of the NCCL broadcast
    dbo = _small_db(oracle, genomes, toy_tax)
    keys, vals = oracle.db_pairs(dbo)
    tc, tp = H.toy_tax_arrays()
    bases, offs, _ = H.make_reads(20000, seed=5, genomes=genomes)
    et, eh, em = oracle.classify(dbo, toy_tax, bases, offs, 31, 31)
    n = min(_ndev(), 4)
    ctxs = capi.open_multi(n, 31, 31)
    try:
        assert len(ctxs) == n
        ctxs[0].load_pairs(keys, vals)
        ctxs[0].load_taxonomy(tc, tp)
        capi.replicate(ctxs, root=0)
        for c in ctxs:
            t, h, m = c.classify(bases, offs)
            assert np.array_equal(t, et) and np.array_equal(h, eh) and np.array_equal(m, em)
            k2, v2 = c.table_dump()
            assert np.array_equal(k2, keys) and np.array_equal(v2, vals)
            assert c.table_info()["n_keys"] == keys.size
    finally:
        for c in ctxs:
            c.close()


@needs2
def test_cli_gpus_matches_single(tmp_path, oracle, genomes):
    from bonsai_b200 import build
    from helpers import write_fastq
    cli = build.build_cli()
    d = tmp_path
    nodes = d / "nodes.dmp"
    nodes.write_text("".join("%d\t|\t%d\t|\trank\t|\n" % cp for cp in H.TOY_TAX))
    args = []
    for gi, taxid in enumerate(H.GENOME_TAXIDS):
        b, _ = H.genome_records(genomes, gi)
        p = d / ("g%d.fa" % gi)
        p.write_text(">g%d\n%s\n" % (gi, bytes(b[:200_000]).decode()))
        args.append("%d=%s" % (taxid, p))
    db = d / "four.db"
    subprocess.check_call([cli, "build", "-k", "31", str(db), str(nodes)] + args)
    bases, offs, _ = H.make_reads(30000, seed=11, genomes=genomes)
    seqs = [bytes(bases[int(offs[i]):int(offs[i + 1])]).decode() for i in range(30000)]
    names = ["r%d" % i for i in range(30000)]
    write_fastq(d / "r.fq", names, seqs)
    outs = {}
    for g in (1, 2):
        r = subprocess.run([cli, "classify", "-a", "-p", "4", "-c", "300000", "--gpus", str(g), str(db), str(nodes), str(d / "r.fq")],
                           capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        outs[g] = (r.stdout, r.stderr.decode().strip().splitlines()[-1])
    assert outs[1][0] == outs[2][0] and outs[1][0].count(b"\n") == 30000, (outs[2][0][:400], outs[2][1])
    assert outs[1][1] == outs[2][1]                       # "classified N, unclassified M" adds up over the GPUs


@needs2
def test_torch_distributed_nccl_broadcast(tmp_path):
    """two ranks, real NCCL: rank 0 builds, sharding.replicate_db broadcasts, both ranks classify a common probe; rank 0 checks
    the probe against the oracle and the all-gathered digests against its own"""
    script = tmp_path / "rank.py"
    script.write_text('''
import os, sys, hashlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import helpers as H
from bonsai_b200 import capi, sharding
from oracle import pyoracle as po
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
g = H.load_genomes(); tc, tp = H.toy_tax_arrays()
o = po.load_oracle(); T = o.tax_from_pairs(tc, tp)
dbo = o.db_new()
for gi, taxid in enumerate(H.GENOME_TAXIDS):
    b, _ = H.genome_records(g, gi)
    o.db_add_genome(dbo, T, (b[:300000].copy(), np.array([0, 300000], np.uint64)), taxid, 31, 31)
keys, vals = o.db_pairs(dbo)
ctx = capi.Context(31, 31, device=rank)
if rank == 0:
    ctx.load_pairs(keys, vals); ctx.load_taxonomy(tc, tp)
moved = sharding.replicate_db(ctx, dist, rank, root=0, device=dev)
bases, offs, _ = H.make_reads(20000, seed=5, genomes=g)
t, h, m = ctx.classify(bases, offs)
et, eh, em = o.classify(dbo, T, bases, offs, 31, 31)
ok = bool(np.array_equal(t, et) and np.array_equal(h, eh) and np.array_equal(m, em))
k2, v2 = ctx.table_dump()
ok = ok and bool(np.array_equal(k2, keys) and np.array_equal(v2, vals))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("REPLICAS_OK" if int(flag.item()) == 1 else "REPLICAS_BAD", moved)
dist.destroy_process_group()
''' % (ROOT, ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "REPLICAS_OK" in r.stdout, r.stdout + r.stderr
