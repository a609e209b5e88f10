"""GPU (-m gpu): `bonsai build` + `bonsai classify` end to end (FASTA/FASTQ(.gz) ingest -> C ABI -> Kraken / FASTQ
text), byte-for-byte against the oracle's restatement of classify_seq + the reference's emitters."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from helpers import write_fastq
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(tmp_path_factory, oracle, genomes):
    from bonsai_b200 import build
    cli = build.build_cli()
    d = tmp_path_factory.mktemp("cli")
    nodes = d / "nodes.dmp"
    nodes.write_text("".join("%d\t|\t%d\t|\trank\t|\n" % cp for cp in H.TOY_TAX))
    tax = oracle.tax_load(str(nodes))
    # genomes: two 150 kb records per genome, FASTA (one gzipped)
    args, dbo = [], oracle.db_new()
    for gi, taxid in enumerate(H.GENOME_TAXIDS):
        b, off = H.genome_records(genomes, gi)
        recs = [bytes(b[:150_000]), bytes(b[200_000:350_000])]
        p = d / ("g%d.fa%s" % (gi, ".gz" if gi == 1 else ""))
        txt = "".join(">rec%d some comment\n%s\n" % (i, "\n".join(r.decode()[j:j + 70] for j in range(0, len(r), 70))) for i, r in enumerate(recs))
        (gzip.open(p, "wt") if gi == 1 else open(p, "w")).write(txt)
        args.append("%d=%s" % (taxid, p))
        oracle.db_add_genome(dbo, tax, recs, taxid, 31, 31)
    db = d / "four.db"
    subprocess.check_call([cli, "build", "-k", "31", str(db), str(nodes)] + args)
    return dict(cli=cli, dir=d, nodes=nodes, db=db, dbo=dbo, tax=tax)


def test_build_matches_oracle(setup, oracle):
    out = subprocess.check_output([setup["cli"], "dbcheck", str(setup["db"])], text=True)
    k, v = oracle.db_pairs(setup["dbo"])
    with np.errstate(over="ignore"):
        assert "occupied=%d " % k.size in out
        assert "key_xor=%016x" % int(np.bitwise_xor.reduce(k)) in out
        assert "key_sum=%016x" % int(k.sum(dtype=np.uint64)) in out
        assert "val_sum=%d" % int(v.sum(dtype=np.uint64)) in out


@pytest.mark.parametrize("mode", ["kraken_all", "kraken_classified_only", "fastq_all", "fastq_kraken", "kraken_nocanon",
                                  "kraken_all_p4", "fastq_all_p4", "kraken_all_plain_p4", "fastq_all_plain_p3"])
def test_classify_text(setup, oracle, genomes, mode):
    g = genomes
    rng = np.random.default_rng(3)
    seqs, names = [], []
    for i in range(1500):
        gi = i % 4
        b, _ = H.genome_records(g, gi)
        if i % 9 == 0:
            s = bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 150)])
        else:
            st = int(rng.integers(0, 340_000))
            s = bytearray(b[st:st + int(rng.integers(20, 260))].tobytes())
            if i % 13 == 0 and len(s) > 40:
                s[int(rng.integers(0, len(s)))] = ord("N")
            s = bytes(s)
        seqs.append(s.decode())
        names.append("read%d/1" % i if i % 2 else "read%d" % i)
    quals = ["".join(chr(33 + (j * 7 + i) % 40) for j in range(len(s))) for i, s in enumerate(seqs)]
    # *_plain_*: an uncompressed file goes through the parallel index (several small windows) instead of kseq
    fq = setup["dir"] / ("reads_%s.fq%s" % (mode, "" if "_plain_" in mode else ".gz"))
    use_q = mode != "kraken_classified_only"
    write_fastq(fq, names, seqs, quals if use_q else None)
    flags = {"kraken_all": ["-a"], "kraken_classified_only": [], "fastq_all": ["-a", "-f", "-K"], "fastq_kraken": ["-a", "-f", "-k"],
             "kraken_nocanon": ["-a", "-C"],
             # one big chunk, four formatter threads (the lean kernel for the FASTQ-style output without run lists)
             "kraken_all_p4": ["-a", "-p", "4"], "fastq_all_p4": ["-a", "-f", "-K", "-p", "4"],
             "kraken_all_plain_p4": ["-a", "-p", "4"], "fastq_all_plain_p3": ["-a", "-f", "-K", "-p", "3"]}[mode]
    chunk = "100000000" if mode.endswith("_p4") else "20000"
    kw = {"kraken_all": dict(emit_all=True, emit_fastq=False, emit_kraken=True),
          "kraken_classified_only": dict(emit_all=False, emit_fastq=False, emit_kraken=True),
          "fastq_all": dict(emit_all=True, emit_fastq=True, emit_kraken=False),
          "fastq_kraken": dict(emit_all=True, emit_fastq=True, emit_kraken=True),
          "kraken_nocanon": dict(emit_all=True, emit_fastq=False, emit_kraken=True, canon=False),
          "kraken_all_p4": dict(emit_all=True, emit_fastq=False, emit_kraken=True),
          "fastq_all_p4": dict(emit_all=True, emit_fastq=True, emit_kraken=False),
          "kraken_all_plain_p4": dict(emit_all=True, emit_fastq=False, emit_kraken=True),
          "fastq_all_plain_p3": dict(emit_all=True, emit_fastq=True, emit_kraken=False)}[mode]
    outp = setup["dir"] / ("out_%s.txt" % mode)
    r = subprocess.run([setup["cli"], "classify"] + flags + ["-c", chunk, "-o", str(outp), str(setup["db"]), str(setup["nodes"]), str(fq)],
                       capture_output=True, text=True, env=dict(os.environ, BNS_B200_FASTQ_WINDOW="60000"))
    assert r.returncode == 0, r.stderr
    bases, offs = po.pack_reads(seqs)
    trimmed = [n[:-2] if n.endswith("/1") else n for n in names]          # trim_readno
    exp, nc, nu = oracle.classify_text(setup["dbo"], setup["tax"], bases, offs, trimmed, 31, 31,
                                       quals=quals if use_q else None, **kw)
    assert open(outp, "rb").read() == exp
    assert "classified %d, unclassified %d" % (nc, nu) in r.stderr


def test_classify_paired(setup, oracle, genomes):
    b, _ = H.genome_records(genomes, 0)
    rng = np.random.default_rng(8)
    s1, s2, n1, n2 = [], [], [], []
    for i in range(400):
        st = int(rng.integers(0, 140_000))
        s1.append(b[st:st + 100].tobytes().decode())
        s2.append(b[st + 200:st + 320].tobytes().decode() if i % 5 else "ACGTN" * 20)
        n1.append("p%d/1" % i); n2.append("p%d/2" % i)
    f1, f2 = setup["dir"] / "r1.fa", setup["dir"] / "r2.fa"
    write_fastq(f1, n1, s1); write_fastq(f2, n2, s2)
    r = subprocess.run([setup["cli"], "classify", "-a", "-c", "5000", str(setup["db"]), str(setup["nodes"]), str(f1), str(f2)],
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    inter = [x for pair in zip(s1, s2) for x in pair]
    names = [x for pair in zip(["p%d" % i for i in range(400)], ["p%d" % i for i in range(400)]) for x in pair]
    bases, offs = po.pack_reads(inter)
    exp, _, _ = oracle.classify_text(setup["dbo"], setup["tax"], bases, offs, names, 31, 31, emit_all=True, emit_fastq=False,
                                     emit_kraken=True, paired=True)
    assert r.stdout == exp


def test_build_entropy_minimised(setup, oracle, genomes):
    """`bonsai build -e -w 50`: the entropy-minimised DB (the reference's path-overload encoder with ent_score) -- also
    checks that the C++ layer picks the cast mode a -march=native reference build would have on this host"""
    d = setup["dir"]
    args, dbo = [], oracle.db_new()
    mode = po.CAST_SATURATE if po.host_has_avx512() else po.CAST_WRAP
    for gi, taxid in enumerate(H.GENOME_TAXIDS):
        b, off = H.genome_records(genomes, gi)
        recs = [bytes(b[:150_000]), bytes(b[200_000:350_000])]
        args.append("%d=%s" % (taxid, d / ("g%d.fa%s" % (gi, ".gz" if gi == 1 else ""))))
        oracle.db_add_genome(dbo, setup["tax"], recs, taxid, 31, 50, None, po.SCORE_ENTROPY, True, cast_mode=mode)
    db = d / "ent.db"
    subprocess.check_call([setup["cli"], "build", "-k", "31", "-w", "50", "-e", str(db), str(setup["nodes"])] + args)
    out = subprocess.check_output([setup["cli"], "dbcheck", str(db)], text=True)
    k, v = oracle.db_pairs(dbo)
    with np.errstate(over="ignore"):
        assert "k=31 w=50" in out and "occupied=%d " % k.size in out
        assert "key_xor=%016x" % int(np.bitwise_xor.reduce(k)) in out
        assert "val_sum=%d" % int(v.sum(dtype=np.uint64)) in out


def test_classify_paired_plain_equals_gz(setup, genomes):
    """Mate files through the parallel index (plain FASTQ with small windows, gzip with small and with default inflate
    windows, BGZF, 4 threads) give the bytes the kseq path gives for the same files -- including the hand-over when the second
    file ends early (bseq_read's warning)."""
    b, _ = H.genome_records(genomes, 1)
    rng = np.random.default_rng(18)
    s1, s2, n1, n2 = [], [], [], []
    for i in range(900):
        st = int(rng.integers(0, 140_000))
        s1.append(b[st:st + int(rng.integers(40, 160))].tobytes().decode())
        s2.append(b[st + 200:st + 200 + int(rng.integers(40, 160))].tobytes().decode())
        n1.append("q%d/1" % i); n2.append("q%d/2" % i)
    q1 = ["I" * len(s) for s in s1]; q2 = ["@" + "F" * (len(s) - 1) for s in s2]
    outs = {}
    kinds = {"plain": (".fq", {"BNS_B200_FASTQ_WINDOW": "40000"}), "gz_kseq": (".fq.gz", {"BNS_B200_INGEST": "kseq"}),
             "gz_windows": (".fq.gz", {"BNS_B200_GZ_WINDOW": "50000"}), "gz": (".fq.gz", {}),
             "bgzf": (".fq.bgz", {"BNS_B200_GZ_WINDOW": "140000"})}          # blocks inflated in parallel out of the mapping
    for kind, (ext, env) in kinds.items():
        f1, f2 = setup["dir"] / ("pp1" + ext), setup["dir"] / ("pp2" + ext)
        write_fastq(f1, n1, s1, q1); write_fastq(f2, n2[:880], s2[:880], q2[:880])       # the mates' file is 20 records short
        r = subprocess.run([setup["cli"], "classify", "-a", "-f", "-k", "-p", "4", "-c", "30000", str(setup["db"]), str(setup["nodes"]), str(f1), str(f2)],
                           capture_output=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        assert b"the 2nd file has fewer sequences" in r.stderr
        outs[kind] = r.stdout
    # 9 lines per pair: the second mate's header repeats the classification line INCLUDING its newline (classifier.h:99-104)
    assert outs["gz_kseq"].count(b"\n") == 880 * 9
    for kind in kinds:
        assert outs[kind] == outs["gz_kseq"], kind
