#!/usr/bin/env python
"""Regenerates tests/golden/{genomes4.npz,golden.json}. Runs ONLY in the build container: it reads the
reference's data fixtures (test/GCF_*.fna.gz, test/phix.fa) and calls the UNMODIFIED reference code
through oracle/_ref (oracle/ref_driver.cpp). The outputs are what travels; nothing at test time reads
/root/reference.

    python tests/golden/make_golden.py
"""
import glob
import gzip
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po  # noqa: E402
import helpers as H  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def read_fasta(path):
    op = gzip.open if path.endswith(".gz") else open
    recs, cur = [], []
    with op(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if cur:
                    recs.append("".join(cur))
                cur = []
            else:
                cur.append(line.strip())
    if cur:
        recs.append("".join(cur))
    return recs


def make_genomes():
    paths = sorted(glob.glob(os.path.join(REF, "test", "GCF_*.fna.gz")))
    assert len(paths) == 4
    lens, gidx, seqs = [], [], []
    for gi, p in enumerate(paths):
        for r in read_fasta(p):
            lens.append(len(r)); gidx.append(gi); seqs.append(r)
    allb = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    phix = np.frombuffer("".join(read_fasta(os.path.join(REF, "test", "phix.fa"))).encode(), dtype=np.uint8)
    np.savez(os.path.join(OUT, "genomes4.npz"), packed=H.pack2bit(allb), contig_len=np.array(lens, np.uint32),
             contig_genome=np.array(gidx, np.uint8), phix_packed=H.pack2bit(phix), phix_len=np.int64(phix.size))
    return paths


def hx(a):
    return [format(int(x), "x") for x in a]


def pairs_digest(k, v):
    h = hashlib.md5()
    h.update(np.ascontiguousarray(k).tobytes()); h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest()


def main():
    paths = make_genomes()
    H._genomes = None
    g = H.load_genomes()
    R4, R3 = po.load_ref("v4"), po.load_ref("v3")
    assert R4 is not None and R3 is not None, "run `make -C oracle ref` first"
    G = {"reference": "dnbaker/bonsai@6741de9c", "built_with": [R4.build_info(), R3.build_info()]}

    # ---- scalars ------------------------------------------------------------------------------
    xs = [0, 1, 2, 0x0123456789abcdef, 2**62 - 1, 2**64 - 1, 0x9E3779B97F4A7C15]
    G["lex_score"] = {format(x, "x"): format(R4.lex_score(x), "x") for x in xs}
    G["wang64"] = {format(x, "x"): format(R4.wang64(x), "x") for x in xs}
    G["rc"] = [[format(x, "x"), k, format(R4.rc(x & ((1 << (2 * k)) - 1) if k < 32 else x, k), "x"),
                format(R4.canonical(x & ((1 << (2 * k)) - 1) if k < 32 else x, k), "x")]
               for x in xs for k in (1, 5, 16, 31, 32)]
    # ---- spacer -------------------------------------------------------------------------------
    sp_cases = [(31, 31, None), (31, 50, None), (31, 0, None), (31, 31, H.SPACED_GAPS), (31, 60, H.SPACED_GAPS),
                (5, 8, None), (5, 5, [1, 2, 0, 0]), (13, 20, [0] * 12), (2, 2, [3])]
    G["spacer"] = [[k, w, gaps, list(R4.spacer(k, w, gaps))] for k, w, gaps in sp_cases]
    G["parse_spacing"] = [[s, k, [int(x) for x in R4.parse_spacing(s, k)]]
                          for s, k in (("", 5), ("1x3,0x5", 9), ("0x2,1,2x2,0", 7), ("3", 2), ("1,2,3", 4))]
    # ---- small encoder vectors (SURVEY C.5 strings and a few more) -------------------------------
    small = ["ACGTACGTAC", "acgtNACGTTT", "ACGT", "ACGTU", "AAAAANAAAAACCCCC", "TTTTTTTTTT", "AAAAAAAAAACGT",
             "ACATGCTAGCATGCTGACTGACTGATCGATCGTA", "", "N", "ACGTNNNNACGTACGTNACGTACG",
             "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTACGT", "GATTACAGATTACAGATTACAnGATTACAGATTACA"]
    enc = []
    for s in small:
        for (k, w, gaps) in ((5, 5, None), (5, 8, None), (5, 5, [1, 2, 0, 0]), (5, 12, [1, 2, 0, 0]), (32, 32, None), (32, 36, None)):
            for score in (0, 1):
                for canon in (0, 1):
                    for api in (0, 1):
                        e4 = hx(R4.encode(s, k, w, gaps, score, canon, api))
                        e3 = hx(R3.encode(s, k, w, gaps, score, canon, api))
                        enc.append(dict(seq=s, k=k, w=w, gaps=gaps, score=score, canon=canon, api=api,
                                        saturate=e4, wrap=e3 if e3 != e4 else None))
    G["encode_small"] = enc
    # ---- streams over seeded reads + phiX ----------------------------------------------------------
    bases, offs, origin = H.make_reads(2000, seed=42)
    rd = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(2000)]
    G["reads2000_md5"] = hashlib.md5(bases.tobytes()).hexdigest()
    modes = [("lex_w31_canon", 31, 31, None, 0, 1, 0), ("lex_w31_nocanon", 31, 31, None, 0, 0, 0),
             ("lex_w50_canon", 31, 50, None, 0, 1, 0), ("lex_w50_nocanon", 31, 50, None, 0, 0, 0),
             ("ent_w50_canon", 31, 50, None, 1, 1, 0), ("ent_w50_nocanon", 31, 50, None, 1, 0, 0),
             ("spaced_string_api", 31, 31, H.SPACED_GAPS, 0, 0, 0), ("spaced_path_api", 31, 31, H.SPACED_GAPS, 0, 0, 1),
             ("spaced_w60_path_api", 31, 60, H.SPACED_GAPS, 0, 0, 1), ("ent_w50_canon_path_api", 31, 50, None, 1, 1, 1),
             ("lex_w50_canon_path_api", 31, 50, None, 0, 1, 1)]
    streams = {}
    for name, k, w, gaps, score, canon, api in modes:
        for R, tag in ((R4, "saturate"), (R3, "wrap")):
            if tag == "wrap" and score == 0:
                continue
            allk = np.concatenate([R.encode(s, k, w, gaps, score, canon, api) for s in rd])
            px = R.encode(bytes(g["phix"]), k, w, gaps, score, canon, api)
            streams[name + ":" + tag] = dict(k=k, w=w, gaps=gaps, score=score, canon=canon, api=api,
                                             reads=list(H.digest(allk)), phix=list(H.digest(px)),
                                             phix_distinct=int(np.unique(px).size))
    G["streams"] = streams
    # ---- taxonomy / resolve_tree / lca ----------------------------------------------------------
    tc, tp = H.toy_tax_arrays()
    T = R4.tax_from_pairs(tc, tp)
    res_cases = [[], [(11, 5)], [(11, 5), (12, 5)], [(11, 3), (10, 2), (12, 4)], [(10, 5), (11, 1)], [(20, 3), (11, 3)],
                 [(11, 5), (12, 5), (13, 6)], [(13, 6), (11, 6), (12, 5)], [(2, 10), (11, 1), (20, 1)],
                 [(10, 4), (11, 2), (12, 2)], [(1, 3)], [(11, 2), (20, 2), (10, 1)], [(12, 65535), (11, 65535), (10, 1)]]
    rng = np.random.default_rng(7)
    for _ in range(200):
        n = int(rng.integers(1, 7))
        ids = rng.choice([1, 2, 10, 11, 12, 13, 20], n, replace=False)
        res_cases.append([(int(t), int(rng.integers(1, 6))) for t in ids])
    G["resolve"] = [[c, int(R4.resolve(T, [a for a, _ in c], [b for _, b in c]))] for c in res_cases]
    ids = [0, 1, 2, 10, 11, 12, 13, 20]
    G["lca"] = [[a, b, int(R4.lca(T, a, b))] for a in ids for b in ids]
    # ---- databases (reference path-overload encoders + update_lca_map) --------------------------------
    dbs = {}
    dbspec = {"lex_k31_w31": (31, 31, None, 0, 1), "ent_k31_w50": (31, 50, None, 1, 1),
              "spaced_k31_c40": (31, 31, H.SPACED_GAPS, 0, 0)}
    handles = {}
    for name, (k, w, gaps, score, canon) in dbspec.items():
        db = R4.db_build(T, paths, H.GENOME_TAXIDS, k, w, gaps, score, canon)
        kk, vv = R4.db_pairs(db)
        vals, cnts = np.unique(vv, return_counts=True)
        dbs[name] = dict(k=k, w=w, gaps=gaps, score=score, canon=canon, size=int(kk.size),
                         hist={int(a): int(b) for a, b in zip(vals, cnts)}, md5=pairs_digest(kk, vv),
                         n_buckets=int(R4.db_arrays(db)[3]))
        handles[name] = db
        print(name, dbs[name]["size"], dbs[name]["hist"], flush=True)
    G["dbs"] = dbs
    # ---- classification of the 2000 reads ------------------------------------------------------
    names = ["r%d_%s" % (i, "rand" if origin[i] < 0 else "g%d" % origin[i]) for i in range(2000)]
    cls = {}
    for cname, dbname, (k, w, gaps, canon, api) in (
            ("config1_lex_w31", "lex_k31_w31", (31, 31, None, 1, 0)),
            ("config2_entdb", "ent_k31_w50", (31, 31, None, 1, 0)),
            ("config4_spaced", "spaced_k31_c40", (31, 31, H.SPACED_GAPS, 0, 1)),
            ("windowed_lex_w50_on_reads", "lex_k31_w31", (31, 50, None, 1, 0)),
            ("windowed_lex_w50_nocanon", "lex_k31_w31", (31, 50, None, 0, 0))):
        taxon, nhit, nmiss, lists = R4.classify(handles[dbname], T, bases, offs, k, w, gaps, 0, canon, api, want_taxa=True)
        flat = np.concatenate(lists) if lists else np.zeros(0, np.uint32)
        cls[cname] = dict(db=dbname, k=k, w=w, gaps=gaps, canon=canon, api=api, taxon=[int(x) for x in taxon],
                          nhit=[int(x) for x in nhit], nmiss=[int(x) for x in nmiss],
                          taxa_md5=hashlib.md5(flat.tobytes()).hexdigest())
        print(cname, np.unique(taxon, return_counts=True), flush=True)
    # phiX (config 1)
    pb, poff = po.pack_reads([bytes(g["phix"])])
    taxon, nhit, nmiss = R4.classify(handles["lex_k31_w31"], T, pb, poff, 31, 31)
    cls["phix"] = dict(taxon=int(taxon[0]), nhit=int(nhit[0]), nmiss=int(nmiss[0]))
    # paired: reads (2i, 2i+1) as mates
    taxon, nhit, nmiss = R4.classify(handles["lex_k31_w31"], T, bases, offs, 31, 31, paired=True)
    cls["paired_lex_w31"] = dict(taxon=[int(x) for x in taxon], nhit=[int(x) for x in nhit], nmiss=[int(x) for x in nmiss])
    G["classify"] = cls
    # ---- text (unmodified classify_seq) ----------------------------------------------------------
    text = {}
    for tname, kw in (("kraken_all", dict(emit_all=True, emit_fastq=False, emit_kraken=True)),
                      ("kraken_classified_only", dict(emit_all=False, emit_fastq=False, emit_kraken=True)),
                      ("fastq_all", dict(emit_all=True, emit_fastq=True, emit_kraken=False)),
                      ("fastq_kraken_all", dict(emit_all=True, emit_fastq=True, emit_kraken=True)),
                      # paired FASTQ is NOT pinned: append_fastq_classification keeps raw pointers (cms/cme,
                      # classifier.h:79,91) into a buffer that reallocates while the comment is written, so
                      # the unmodified reference emits garbage of arbitrary length there (observed 839 MB).
                      ("kraken_paired", dict(emit_all=True, emit_fastq=False, emit_kraken=True, paired=True))):
        txt, nc, nu = R4.classify_text(handles["lex_k31_w31"], T, bases, offs, names, 31, 31, **kw)
        text[tname] = dict(md5=hashlib.md5(txt).hexdigest(), n_classified=nc, n_unclassified=nu, length=len(txt),
                           head=txt.decode().split("\n")[:4])
    txt, nc, nu = R4.classify_text(handles["lex_k31_w31"], T, pb, poff, ["phix"], 31, 31)
    text["phix_kraken"] = txt.decode()
    G["text"] = text
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(G, f, indent=0, separators=(",", ":"))
    print("wrote golden.json", os.path.getsize(os.path.join(OUT, "golden.json")))


if __name__ == "__main__":
    main()
