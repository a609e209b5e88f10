"""CPU: the host-only parts of the C++ mirror (include/bonsai_b200/bonsai.hpp) through the `bonsai` CLI:
the database file is the raw khash_t(c) layout the reference intends (database.h:81-102, util.h:281-296), so the
restated kh_get must find every key in the arrays the writer produced."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest


@pytest.fixture(scope="module")
def cli():
    from bonsai_b200 import build
    return build.build_cli()


def read_db(path):
    raw = open(path, "rb").read()
    k, w = struct.unpack_from("<II", raw, 0)
    off = 8
    gaps = list(raw[off:off + k - 1]); off += k - 1
    nb, nocc, size, ub = struct.unpack_from("<QQQQ", raw, off); off += 32
    nfl = 1 if nb < 16 else nb >> 4
    flags = np.frombuffer(raw, np.uint32, nfl, off); off += 4 * nfl
    keys = np.frombuffer(raw, np.uint64, nb, off); off += 8 * nb
    vals = np.frombuffer(raw, np.uint32, nb, off); off += 4 * nb
    assert off == len(raw)
    return dict(k=k, w=w, gaps=gaps, n_buckets=nb, n_occupied=nocc, size=size, upper_bound=ub, flags=flags, keys=keys, vals=vals)


@pytest.mark.parametrize("n", [0, 1, 3, 1000, 200000])
def test_db_file_is_a_khash(cli, oracle, tmp_path, n):
    rng = np.random.default_rng(n)
    keys = np.unique(rng.integers(0, 2**62, n, dtype=np.uint64))
    vals = rng.integers(1, 1000, keys.size, dtype=np.uint64).astype(np.uint32)
    pairs = tmp_path / "pairs.bin"
    with open(pairs, "wb") as f:
        f.write(struct.pack("<Q", keys.size)); f.write(keys.tobytes()); f.write(vals.tobytes())
    db = tmp_path / "t.db"
    subprocess.check_call([cli, "dbwrite", str(db), "31", "31", str(pairs)])
    out = subprocess.check_output([cli, "dbcheck", str(db)], text=True)
    d = read_db(db)
    assert (d["k"], d["w"], d["gaps"], d["size"]) == (31, 31, [0] * 30, keys.size)
    assert d["n_buckets"] & (d["n_buckets"] - 1) == 0 and d["size"] <= d["upper_bound"] == int(d["n_buckets"] * 0.77 + 0.5)
    with np.errstate(over="ignore"):
        assert "occupied=%d " % keys.size in out
        assert "key_xor=%016x" % int(np.bitwise_xor.reduce(keys) if keys.size else 0) in out
        assert "val_sum=%d" % int(vals.sum(dtype=np.uint64)) in out
    # the reference's kh_get (restated in the oracle) over the written arrays
    D = oracle.db_from_arrays(d["keys"], d["vals"], d["flags"], d["n_buckets"])
    for i in rng.integers(0, max(keys.size, 1), min(keys.size, 2000)):
        assert oracle.db_get(D, int(keys[i])) == int(vals[i])
    for q in rng.integers(0, 2**62, 200, dtype=np.uint64):
        if q not in keys:
            assert oracle.db_get(D, int(q)) is None
    k2, v2 = oracle.db_pairs(D)
    assert np.array_equal(k2, keys) and np.array_equal(v2, vals)
    # gz variant reads back identically
    dbz = tmp_path / "t.db.gz"
    subprocess.check_call([cli, "dbwrite", str(dbz), "31", "31", str(pairs), "gz"])
    assert subprocess.check_output([cli, "dbcheck", str(dbz)], text=True) == out


def test_cli_usage(cli):
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode != 0 and "classify" in r.stderr
    r = subprocess.run([cli, "classify"], capture_output=True, text=True)
    assert r.returncode != 0 and "-k:\tEmit kraken-style output." in r.stderr


def test_hist(cli, tmp_path):
    keys = np.arange(1, 1001, dtype=np.uint64) * np.uint64(7919)
    vals = np.where(np.arange(1000) % 10 == 0, 20, np.where(np.arange(1000) % 3 == 0, 11, 12)).astype(np.uint32)
    pairs = tmp_path / "p.bin"
    with open(pairs, "wb") as f:
        f.write(struct.pack("<Q", keys.size)); f.write(keys.tobytes()); f.write(vals.tobytes())
    db = tmp_path / "h.db"
    subprocess.check_call([cli, "dbwrite", str(db), "31", "50", str(pairs)])
    out = subprocess.check_output([cli, "hist", str(db)], text=True).splitlines()
    assert out[0] == "Name\tCount"
    rows = [tuple(map(int, l.split("\t"))) for l in out[1:]]
    u, c = np.unique(vals, return_counts=True)
    assert sorted(rows) == sorted(zip(u.tolist(), c.tolist()))
    assert [r[1] for r in rows] == sorted(r[1] for r in rows)


@pytest.mark.parametrize("scanner", ["avx2", "scalar"])
def test_parallel_ingest_index_matches_kseq(tmp_path, scanner):
    """detail::SimpleFile (parallel index of plain, gzip or BGZF 4-line FASTQ / 2-line FASTA, hand-over to kseq elsewhere) yields
    the same records as the kseq state machine, window by window and batch by batch (single and mate files), with the AVX2 line
    scanner (where the CPU has it) and with the memchr-per-line one: tests/host/ingest_index.cpp, host code only."""
    import shutil
    import subprocess
    gxx = shutil.which("g++") or "/usr/bin/g++"
    exe = str(tmp_path / "ingest_index")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "ingest_index.cpp")
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-o", exe, src, "-lz", "-lpthread"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    env = dict(os.environ)
    if scanner == "scalar":
        env["BNS_B200_INDEX_SCALAR"] = "1"
    r = subprocess.run([exe, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout and r.stdout.count(" ok") == 80, r.stdout


def test_cli_host_pipeline_ingest_kinds(tmp_path, cli):
    """`bonsai classify`'s host pipeline end to end without a device: the CLI source is linked against a DUMMY ABI
    (tests/host/abi_stub.cpp: results are an arbitrary function of each record's bases) and must print byte-identical
    text whichever way the same records are read -- kseq, or the parallel index over plain, gzip and BGZF files, at several
    window / batch sizes and thread counts, single and mate files (the mates' file 7 records short), Kraken and FASTQ text."""
    import hashlib
    import shutil
    from helpers import write_fastq
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    gxx = shutil.which("g++") or "/usr/bin/g++"
    exe = str(tmp_path / "bonsai_stub")
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-o", exe, os.path.join(root, "bonsai_b200", "csrc", "cli", "bonsai_main.cpp"),
                        os.path.join(here, "host", "abi_stub.cpp"), "-lz", "-lpthread"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    d = tmp_path
    keys = np.arange(1, 100, dtype=np.uint64)
    with open(d / "pairs.bin", "wb") as f:
        f.write(struct.pack("<Q", keys.size)); f.write(keys.tobytes()); f.write(np.full(keys.size, 11, np.uint32).tobytes())
    subprocess.check_call([cli, "dbwrite", str(d / "t.db"), "31", "31", str(d / "pairs.bin")])
    (d / "nodes.dmp").write_text("1\t|\t1\t|\n2\t|\t1\t|\n10\t|\t2\t|\n11\t|\t10\t|\n12\t|\t10\t|\n")
    rng = np.random.default_rng(3)

    def mk(n, tag):
        names = ["read%d/%s" % (i, tag) for i in range(n)]
        seqs, quals = [], []
        for l in rng.integers(20, 260, n):
            seqs.append(np.frombuffer(b"ACGTN", np.uint8)[rng.choice(5, p=[.245, .245, .245, .245, .02], size=int(l))].tobytes().decode())
            quals.append(rng.integers(33, 93, int(l), dtype=np.uint8).tobytes().decode())
        return names, seqs, quals
    a, b = mk(9000, "1"), mk(8993, "2")
    for ext in (".fq", ".fq.gz", ".fq.bgz"):
        write_fastq(d / ("r1" + ext), *a); write_fastq(d / ("r2" + ext), *b)
    for ext in (".fa", ".fa.gz", ".fa.bgz"):
        write_fastq(d / ("s" + ext), a[0], a[1])

    def run(env, flags, *files):
        r = subprocess.run([exe, "classify"] + flags + [str(d / "t.db"), str(d / "nodes.dmp")] + [str(d / f) for f in files],
                           capture_output=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        return hashlib.md5(r.stdout).hexdigest(), r.stdout.count(b"\n"), r.stderr

    kinds = [({"BNS_B200_INGEST": "kseq"}, ".gz", "50000"), ({"BNS_B200_FASTQ_WINDOW": "150000"}, "", "50000"), ({}, "", "3000000"),
             ({}, ".gz", "50000"), ({"BNS_B200_GZ_WINDOW": "100000"}, ".gz", "50000"), ({"BNS_B200_GZ_WINDOW": "140000"}, ".bgz", "50000"),
             ({}, ".bgz", "400000")]
    for flags, per_record in ((["-a", "-p", "4"], 1), (["-a", "-f", "-k", "-p", "3"], 4), (["-p", "5"], None)):
        single = [run(env, flags + ["-c", c], "r1.fq" + ext) for env, ext, c in kinds]
        assert len({x[0] for x in single}) == 1, (flags, single)
        assert per_record is None or single[0][1] == 9000 * per_record
        paired = [run(env, flags + ["-c", c], "r1.fq" + ext, "r2.fq" + ext) for env, ext, c in kinds]
        paired.append(run({"BNS_B200_GZ_WINDOW": "140000"}, flags + ["-c", "50000"], "r1.fq", "r2.fq.bgz"))
        assert len({x[0] for x in paired}) == 1, (flags, paired)
        assert all(b"the 2nd file has fewer sequences" in x[2] for x in paired)
    fasta = [run(env, ["-a", "-f", "-k", "-p", "4", "-c", c], "s.fa" + ext) for env, ext, c in kinds]
    assert len({x[0] for x in fasta}) == 1 and fasta[0][1] == 9000 * 4
    # --gpus N: the batches are dealt round-robin to N worker contexts and the text is reassembled in batch order
    one = run({}, ["-a", "-p", "4", "-c", "50000"], "r1.fq", "r2.fq")
    for g in ("2", "3", "5"):
        many = run({}, ["-a", "-p", "4", "-c", "50000", "--gpus", g], "r1.fq", "r2.fq")
        assert many[0] == one[0] and many[1] == one[1], g
        assert many[2].split(b"classified")[-2:] == one[2].split(b"classified")[-2:]      # the counters add up over the contexts


def test_reference_style_caller_compiles_and_runs(tmp_path, cli):
    """tests/host/ref_style_main.cpp spells the classify path the way the reference's own classify_main does (Database<khash_t(c)>,
    ClassifierGeneric(db.db_, ...), khash_t(p) *build_parent_map, process_dataset, classify_seqs(.., ks::string &, .., ForPool &),
    kh_destroy(p, ..)): it must compile against bonsai.hpp unchanged and run (here against the dummy ABI)."""
    import shutil
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "ref_style")
    r = subprocess.run([shutil.which("g++") or "/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-o", exe, os.path.join(here, "host", "ref_style_main.cpp"),
                        os.path.join(here, "host", "abi_stub.cpp"), "-lz", "-lpthread"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    keys = np.arange(1, 50, dtype=np.uint64)
    with open(tmp_path / "pairs.bin", "wb") as f:
        f.write(struct.pack("<Q", keys.size)); f.write(keys.tobytes()); f.write(np.full(keys.size, 11, np.uint32).tobytes())
    subprocess.check_call([cli, "dbwrite", str(tmp_path / "t.db"), "31", "31", str(tmp_path / "pairs.bin")])
    (tmp_path / "nodes.dmp").write_text("1\t|\t1\t|\n2\t|\t1\t|\n10\t|\t2\t|\n11\t|\t10\t|\n")
    (tmp_path / "r.fq").write_text("".join("@q%d\n%s\n+\n%s\n" % (i, "ACGTTGCA" * 10, "I" * 80) for i in range(300)))
    r = subprocess.run([exe, str(tmp_path / "t.db"), str(tmp_path / "nodes.dmp"), str(tmp_path / "r.fq")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 302 and all(l[0] in "CU" and l.split("\t")[1].startswith(("q", "a", "b")) for l in lines)


def test_cpp_encoder_surface_compiles(tmp_path):
    """tests/host/encoder_api.cpp (the GPU test of the C++ Encoder mirror, incl. assign / next_kmer / next_minimizer) compiles
    and links against the dummy ABI; without a device every case reports the library's refusal instead of a stream."""
    import shutil
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "encoder_api")
    r = subprocess.run([shutil.which("g++") or "/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(here, "host", "encoder_api.cpp"),
                        os.path.join(here, "host", "abi_stub.cpp"), "-lz", "-lpthread"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    (tmp_path / "cases.txt").write_text("2 31 31 0 0 - %s 0\n3 5 9 0 0 1,0,2,0 - 0\n" % ("ACGT" * 10))
    r = subprocess.run([exe, str(tmp_path / "cases.txt")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "2 cases, 1 failures" in r.stdout and "exception" in r.stdout, r.stdout     # the dummy ABI has no encoder; an empty string needs none


def test_cpp_encoder_file_overloads_batching(tmp_path):
    """Encoder::for_each over a path / std::string / gzFile / list and for_each_canon / for_each_uncanon gather the records of a
    file into batches; against the dummy ABI, every batch size must hand fn what record-by-record calls do, in file order, from
    plain / gzip / xz files and with the canonical context forced either way: tests/host/encoder_files.cpp, host code only."""
    import shutil
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "encoder_files")
    r = subprocess.run([shutil.which("g++") or "/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(here, "host", "encoder_files.cpp"),
                        os.path.join(here, "host", "abi_stub.cpp"), "-lz", "-lpthread"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout and r.stdout.count(" ok") >= 48, r.stdout


@pytest.mark.parametrize("isa", ["native", "avx2", "scalar"])
def test_host_packer_matches_its_format(tmp_path, isa):
    """bonsai_b200/csrc/bns_pack.cpp (ASCII bases -> 2-bit units + suspicious bits + exception words, what the packed-input
    kernel reads) against a byte-at-a-time model of the format, for every instruction-set body this machine can run, and cut
    into pieces on the worker pool: tests/host/pack_check.cpp, host code only."""
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(here, "..", "bonsai_b200", "csrc")
    exe = str(tmp_path / "pack_check")
    r = subprocess.run([shutil.which("g++") or "/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-I", csrc, "-o", exe,
                        os.path.join(here, "host", "pack_check.cpp"), os.path.join(csrc, "bns_pack.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ)
    if isa != "native":
        env["BNS_B200_PACK_ISA"] = isa
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout + r.stderr
