"""CPU: the plain-C restatement (oracle/bns_oracle.c) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py) and against the reference's own known-answer tests."""
import hashlib

import numpy as np
import pytest

import helpers as H
from oracle import pyoracle as po


def hx(a):
    return [format(int(x), "x") for x in a]


def test_scalars(oracle, golden):
    for x, y in golden["lex_score"].items():
        assert format(oracle.lex_score(int(x, 16)), "x") == y
    # SURVEY 8-a5 verified constants
    assert oracle.lex_score(0) == 0x6815802df2ee00c6
    assert oracle.lex_score(1) == 0xf113896f27a3db06
    assert oracle.lex_score(0x0123456789abcdef) == 0x19eaaf3c48964130
    for x, y in golden["wang64"].items():
        assert format(oracle.wang64(int(x, 16)), "x") == y
    for x, k, r, c in golden["rc"]:
        v = int(x, 16)
        v = v & ((1 << (2 * k)) - 1) if k < 32 else v
        assert format(oracle.rc(v, k), "x") == r
        assert format(oracle.canonical(v, k), "x") == c
    assert oracle.canonical(2**64 - 1, 31) == 0


def test_spacer(oracle, golden):
    for k, w, gaps, exp in golden["spacer"]:
        assert list(oracle.spacer(k, w, gaps)) == exp
    for s, k, exp in golden["parse_spacing"]:
        assert [int(x) for x in oracle.parse_spacing(s, k)] == exp


def test_encode_small(oracle, golden):
    for e in golden["encode_small"]:
        a = hx(oracle.encode(e["seq"], e["k"], e["w"], e["gaps"], e["score"], e["canon"], e["api"], cast_mode=po.CAST_SATURATE))
        assert a == e["saturate"], e
        b = hx(oracle.encode(e["seq"], e["k"], e["w"], e["gaps"], e["score"], e["canon"], e["api"], cast_mode=po.CAST_WRAP))
        assert b == (e["wrap"] if e["wrap"] is not None else e["saturate"]), e


def test_survey_c5_vectors(oracle):
    """SURVEY App. C.5 (reference output recorded by the survey, independent of make_golden.py)."""
    enc = lambda s, k, w, **kw: hx(oracle.encode(s, k, w, **kw))
    assert enc("ACGTACGTAC", 5, 5) == "6c 1b1 1b1 6c 6c 1b1".split()
    assert enc("ACGTACGTAC", 5, 5, canon=False) == "6c 1b1 2c6 31b 6c 1b1".split()
    assert enc("ACGTACGTAC", 5, 8) == ["6c"] * 3
    assert enc("ACGTACGTAC", 5, 8, canon=False) == ["31b"] * 3
    assert enc("acgtNACGTTT", 5, 5) == ["1b", "6"]
    assert enc("acgtNACGTTT", 5, 5, canon=False) == ["6f", "1bf"]
    assert enc("acgtNACGTTT", 5, 8) == ["0", "0", "1b", "1b"]
    assert enc("acgtNACGTTT", 5, 8, canon=False) == ["1bf"]
    for s in ("ACGT", "ACGTU"):
        for canon in (False, True):
            for w in (5, 8):
                assert enc(s, 5, w, canon=canon) == []
    assert enc("AAAAANAAAAACCCCC", 5, 8) == "0 0 0 0 0 0 15 15 15".split()
    assert enc("AAAAANAAAAACCCCC", 5, 8, canon=False) == "0 15 15 15".split()
    assert enc("TTTTTTTTTT", 5, 5) == ["0"] * 6
    assert enc("TTTTTTTTTT", 5, 8, canon=False) == ["3ff"] * 3
    ent = dict(score=po.SCORE_ENTROPY, cast_mode=po.CAST_SATURATE)
    assert enc("ACGTACGTAC", 5, 8, canon=False, **ent) == ["6c"] * 3
    assert enc("AAAAAAAAAACGT", 5, 8, **ent) == ["0"] * 6
    assert enc("AAAAANAAAAACCCCC", 5, 8, **ent) == "0 0 1 155".split()


def test_reference_known_answers(oracle, genomes):
    """The assertions of the reference's own test/encoding.cpp that touch this path."""
    phix = bytes(genomes["phix"])
    assert len(phix) == 5386
    km = oracle.encode(phix, 31, 31, canon=False)
    assert np.unique(km).size == 5356                                   # test/encoding.cpp:122
    can = oracle.encode(phix, 31, 31, canon=True)
    assert set(can.tolist()) == {oracle.canonical(int(x), 31) for x in km}   # :146-147
    for w in (32, 55, 100, 300, 500):                                    # :65-88, spaced k=31
        gaps = [0] * 30
        gaps[0], gaps[1] = 1, 2                                          # the test's `{1,2,0...}` pattern
        c = 31 + 3
        n = oracle.encode(phix, 31, w, gaps, canon=False, api=po.API_PATH).size
        assert n == len(phix) - max(w, c) + 1
    ent = oracle.encode(phix, 31, 60, score=po.SCORE_ENTROPY, canon=True, api=po.API_PATH)
    assert np.unique(ent).size < 5353                                    # :194
    # 20-mer has no comb-52 k-mer, 52-mer has (test/encoding.cpp:48-64 analogue)
    gaps = [1] * 21 + [0] * 9
    c = oracle.spacer(31, 31, gaps)[0]
    assert c == 52
    assert oracle.encode("ACGT" * 5, 31, 31, gaps, canon=False, api=po.API_PATH).size == 0
    assert oracle.encode("ACGT" * 13, 31, 31, gaps, canon=False, api=po.API_PATH).size == 1


def test_reads_fixture_is_stable(reads2000, golden):
    assert hashlib.md5(reads2000[0].tobytes()).hexdigest() == golden["reads2000_md5"]


@pytest.mark.parametrize("tag", ["saturate", "wrap"])
def test_streams(oracle, golden, genomes, reads2000, tag):
    bases, offs, _ = reads2000
    rd = [bytes(bases[int(offs[i]):int(offs[i + 1])]) for i in range(2000)]
    cm = po.CAST_SATURATE if tag == "saturate" else po.CAST_WRAP
    n = 0
    for name, s in golden["streams"].items():
        if not name.endswith(":" + tag):
            continue
        n += 1
        allk = np.concatenate([oracle.encode(r, s["k"], s["w"], s["gaps"], s["score"], s["canon"], s["api"], cast_mode=cm) for r in rd])
        assert list(H.digest(allk)) == s["reads"], name
        px = oracle.encode(bytes(genomes["phix"]), s["k"], s["w"], s["gaps"], s["score"], s["canon"], s["api"], cast_mode=cm)
        assert list(H.digest(px)) == s["phix"], name
        assert int(np.unique(px).size) == s["phix_distinct"]
    assert n >= 3


def test_survey_digests(golden):
    """k-mer stream digest of phiX, n = 5356."""
    s = golden["streams"]["lex_w31_canon:saturate"]
    assert s["phix"] == [5356, 0x44d3fbe1abece0f7, 0xa1b246f9dd42748f]
    # and the spaced string-API quirk: nothing is emitted
    assert golden["streams"]["spaced_string_api:saturate"]["reads"][0] == 0


def test_resolve_and_lca(oracle, golden, toy_tax):
    for case, exp in golden["resolve"]:
        assert oracle.resolve(toy_tax, [a for a, _ in case], [b for _, b in case]) == exp, case
    for a, b, exp in golden["lca"]:
        assert oracle.lca(toy_tax, a, b) == exp
    # SURVEY C.5 spot values
    assert oracle.resolve(toy_tax, [11, 12], [5, 5]) == 10
    assert oracle.resolve(toy_tax, [2, 11, 20], [10, 1, 1]) == 2
    assert oracle.lca(toy_tax, 99, 11) == 0xFFFFFFFF


def test_taxonomy_file(oracle, tmp_path):
    p = tmp_path / "nodes.dmp"
    p.write_text("".join("%d\t|\t%d\t|\trank\t|\n" % cp for cp in H.TOY_TAX) + "# comment\n\n")
    t = oracle.tax_load(str(p))
    c, par = oracle.tax_pairs(t)
    assert c.tolist() == [1, 2, 10, 11, 12, 13, 20]
    assert par.tolist() == [0, 1, 2, 10, 10, 10, 2]      # 1 -> 0 forced (util.h:780)
    oracle.tax_free(t)


@pytest.mark.parametrize("name", ["lex_k31_w31", "ent_k31_w50", "spaced_k31_c40"])
def test_db_build(oracle, golden, dbcache, name):
    spec = golden["dbs"][name]
    k, v = oracle.db_pairs(dbcache.get(name))
    assert k.size == spec["size"]
    vals, cnts = np.unique(v, return_counts=True)
    assert {int(a): int(b) for a, b in zip(vals, cnts)} == {int(a): b for a, b in spec["hist"].items()}
    h = hashlib.md5()
    h.update(k.tobytes()); h.update(v.tobytes())
    assert h.hexdigest() == spec["md5"]


def test_classify(oracle, golden, dbcache, toy_tax, reads2000, genomes):
    bases, offs, _ = reads2000
    for cname, c in golden["classify"].items():
        if cname in ("phix", "paired_lex_w31"):
            continue
        taxon, nhit, nmiss, lists = oracle.classify(dbcache.get(c["db"]), toy_tax, bases, offs, c["k"], c["w"], c["gaps"],
                                                    po.SCORE_LEX, c["canon"], c["api"], want_taxa=True)
        assert taxon.tolist() == c["taxon"], cname
        assert nhit.tolist() == c["nhit"], cname
        assert nmiss.tolist() == c["nmiss"], cname
        flat = np.concatenate(lists)
        assert hashlib.md5(flat.tobytes()).hexdigest() == c["taxa_md5"], cname
    db = dbcache.get("lex_k31_w31")
    pb, poff = po.pack_reads([bytes(genomes["phix"])])
    taxon, nhit, nmiss = oracle.classify(db, toy_tax, pb, poff, 31, 31)
    assert (int(taxon[0]), int(nhit[0]), int(nmiss[0])) == (0, 0, 5356)
    c = golden["classify"]["paired_lex_w31"]
    taxon, nhit, nmiss = oracle.classify(db, toy_tax, bases, offs, 31, 31, paired=True)
    assert taxon.tolist() == c["taxon"] and nhit.tolist() == c["nhit"] and nmiss.tolist() == c["nmiss"]


def test_text(oracle, golden, dbcache, toy_tax, reads2000, genomes):
    bases, offs, origin = reads2000
    names = ["r%d_%s" % (i, "rand" if origin[i] < 0 else "g%d" % origin[i]) for i in range(2000)]
    db = dbcache.get("lex_k31_w31")
    modes = dict(kraken_all=dict(emit_all=True, emit_fastq=False, emit_kraken=True),
                 kraken_classified_only=dict(emit_all=False, emit_fastq=False, emit_kraken=True),
                 fastq_all=dict(emit_all=True, emit_fastq=True, emit_kraken=False),
                 fastq_kraken_all=dict(emit_all=True, emit_fastq=True, emit_kraken=True),
                 kraken_paired=dict(emit_all=True, emit_fastq=False, emit_kraken=True, paired=True))
    for name, kw in modes.items():
        exp = golden["text"][name]
        txt, nc, nu = oracle.classify_text(db, toy_tax, bases, offs, names, 31, 31, **kw)
        assert (nc, nu, len(txt)) == (exp["n_classified"], exp["n_unclassified"], exp["length"]), name
        assert hashlib.md5(txt).hexdigest() == exp["md5"], name
    pb, poff = po.pack_reads([bytes(genomes["phix"])])
    txt, _, _ = oracle.classify_text(db, toy_tax, pb, poff, ["phix"], 31, 31)
    assert txt.decode() == golden["text"]["phix_kraken"] == "U\tphix\t0\t5386\tM:5356\t0:0\n"


def test_iterator_surface_golden(oracle):
    """assign / has_next_kmer / next_minimizer / next_canonicalized_minimizer call by call (encoder.h:201-206,594-628):
    the restatement against vectors produced by the unmodified reference headers (tests/golden/make_golden_iter.py)."""
    import json
    import os
    with open(os.path.join(H.GOLDEN, "golden_iter.json")) as f:
        gold = json.load(f)
    assert gold["cast"] == "saturate" and len(gold["cases"]) == 224
    for c in gold["cases"]:
        got = oracle.encode(c["seq"], c["k"], c["w"], c["gaps"], c["score"], c["canon"], po.API_ITER, cast_mode=po.CAST_SATURATE)
        assert ["%x" % int(x) for x in got] == c["values"], {k: c[k] for k in ("k", "w", "gaps", "score", "canon", "seq")}
        # one value per call from the first full window on: l - w + 1 of them (has_next_kmer counts positions l - c + 1)
        cc, ww = oracle.spacer(c["k"], c["w"], c["gaps"])[:2]
        assert len(got) == max(0, len(c["seq"]) - ww + 1)


def test_iterator_surface_reference_pins(oracle):
    """test/encoding.cpp:17-47: the first next_kmer() / next_minimizer() of a 34-base string under the {1,2,0...} comb spells
    the string with its skipped bases dashed out (Spacer::to_string)."""
    test = "ACATGCTAGCATGCTGACTGACTGATCGATCGTA"
    gaps = [1, 2] + [0] * 28
    first = oracle.encode(test, 31, 31, gaps, canon=False, api=po.API_ITER)
    assert first.size == 1
    offs = np.concatenate([[0], np.cumsum(np.array(gaps) + 1)])
    spelled = ["-"] * 34
    for j, o in enumerate(offs):
        spelled[int(o)] = "ACGT"[(int(first[0]) >> (2 * (30 - j))) & 3]
    expect = list(test)
    expect[1] = expect[3] = expect[4] = "-"
    assert spelled == expect
