"""CPU: live differential test of the C restatement against the UNMODIFIED reference headers
(oracle/_ref, built by oracle/Makefile where /root/reference exists; prebuilt on the GPU box).
Skipped when the reference library is absent -- the golden vectors still pin the oracle then."""
import random

import numpy as np
import pytest

import helpers as H
from oracle import pyoracle as po


def _ref(variant):
    r = po.load_ref(variant)
    if r is None:
        pytest.skip("oracle/_ref/libbns_ref_%s.so not available on this host" % variant)
    return r


@pytest.mark.parametrize("variant", ["v4", "v3"])
def test_encode_fuzz(oracle, variant):
    R = _ref(variant)
    assert R.cast_mode == (po.CAST_SATURATE if variant == "v4" else po.CAST_WRAP)
    rng = random.Random(1234 + (variant == "v3"))
    for it in range(400):
        k = rng.choice([1, 2, 3, 5, 7, 13, 16, 21, 31, 31, 31, 32])
        gaps = None
        if rng.random() < 0.3 and k > 1:
            gaps = [rng.choice([0, 0, 0, 1, 2, 3]) for _ in range(k - 1)]
        c = k + (sum(gaps) if gaps else 0)
        w = rng.choice([0, k, c, c + 1, c + 3, c + 19, c + 50])
        L = rng.choice([0, 1, k - 1, k, c, c + 1, c + 5, 60, 150, 300])
        alphabet = rng.choice(["ACGT", "ACGT", "ACGTN", "ACGTacgtNnUu-", "AT", "A", "T", "AC"])
        seq = "".join(rng.choice(alphabet) for _ in range(L))
        if rng.random() < 0.2 and L > 40:
            p = rng.randrange(L - 35)
            seq = seq[:p] + rng.choice("ACGT") * 35 + seq[p + 35:]
        for score in (0, 1):
            for canon in (0, 1):
                for api in (0, 1, 2):                        # 2: next_minimizer / next_canonicalized_minimizer call by call
                    a = R.encode(seq, k, w, gaps, score, canon, api)
                    b = oracle.encode(seq, k, w, gaps, score, canon, api, cast_mode=R.cast_mode)
                    assert np.array_equal(a, b), dict(k=k, w=w, gaps=gaps, score=score, canon=canon, api=api, seq=seq)


def test_cast_restatement(oracle):
    """bo_cast_u64 == what each -march level of the reference build does with (u64)double."""
    vals = [0.0, -0.0, 0.5, -0.5, -0.999, -1.0, -1.5, -5.5, 1e18, 9.3e18, 1.8e19, 1.9e19, 3e19, -9.3e18, -1e19,
            float(2**63), float(2**64), -float(2**63), 123456789.75, float("nan"), float("inf"), float("-inf")]
    exp_sat = {-5.5: 2**64 - 1, -0.5: 0, 1e18: 10**18, 3e19: 2**64 - 1, float(2**63): 2**63}
    exp_wrap = {-5.5: 2**64 - 5, -0.5: 0, 1e18: 10**18, 3e19: 0, float(2**63): 2**63, -1e19: 2**63}
    for v in vals:
        if v in exp_sat:
            assert oracle.cast_u64(v, po.CAST_SATURATE) == exp_sat[v], v
        if v in exp_wrap:
            assert oracle.cast_u64(v, po.CAST_WRAP) == exp_wrap[v], v


def test_resolve_fuzz(oracle):
    R = _ref(None)
    rng = np.random.default_rng(5)
    # random tree over 60 nodes rooted at 1, taxids scattered
    ids = np.unique(rng.integers(2, 5000, 80))[:59]
    nodes = np.concatenate([[1], ids]).astype(np.uint32)
    parent = np.zeros(nodes.size, np.uint32)
    parent[0] = 1
    for i in range(1, nodes.size):
        parent[i] = nodes[rng.integers(0, i)]
    To, Tr = oracle.tax_from_pairs(nodes, parent), R.tax_from_pairs(nodes, parent)
    for _ in range(2000):
        n = int(rng.integers(1, 9))
        taxa = rng.choice(nodes, n, replace=False)
        cnt = rng.integers(1, 5, n)
        assert oracle.resolve(To, taxa, cnt) == R.resolve(Tr, taxa, cnt)
    for _ in range(2000):
        a, b = (int(x) for x in rng.choice(np.concatenate([[0], nodes]), 2))
        assert oracle.lca(To, a, b) == R.lca(Tr, a, b)


def test_classify_and_text_vs_ref(oracle, reads2000):
    R = _ref(None)
    bases, offs, origin = reads2000
    n = 600
    bases, offs = bases[: int(offs[n])], offs[: n + 1]
    rng = np.random.default_rng(11)
    # DB: every 3rd canonical 31-mer of the first 300 reads, random toy taxids
    km = np.unique(np.concatenate([oracle.encode(bytes(bases[int(offs[i]):int(offs[i + 1])]), 31, 31) for i in range(300)]))[::3]
    vals = rng.choice(np.array([2, 10, 11, 12, 13, 20], np.uint32), km.size)
    c, p = H.toy_tax_arrays()
    To, Tr = oracle.tax_from_pairs(c, p), R.tax_from_pairs(c, p)
    Do, Dr = oracle.db_from_pairs(km, vals), R.db_from_pairs(km, vals)
    ko, vo = oracle.db_pairs(Do)
    kr, vr = R.db_pairs(Dr)
    assert np.array_equal(ko, kr) and np.array_equal(vo, vr)
    # the restated kh_get works on the reference's raw arrays too
    keys, vs, flags, nb, _ = R.db_arrays(Dr)
    Dw = oracle.db_from_arrays(keys, vs, flags, nb)
    for key in list(km[:50]) + [1, 2, 3]:
        assert oracle.db_get(Dw, int(key)) == R.db_get(Dr, int(key)) == oracle.db_get(Do, int(key))
    for (k, w, canon, api, paired) in ((31, 31, 1, 0, False), (31, 31, 0, 0, False), (31, 40, 1, 0, False),
                                       (31, 40, 0, 0, False), (31, 31, 1, 0, True)):
        a = R.classify(Dr, Tr, bases, offs, k, w, None, 0, canon, api, paired=paired, want_taxa=True)
        b = oracle.classify(Do, To, bases, offs, k, w, None, 0, canon, api, paired=paired, want_taxa=True)
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x, y)
        assert all(np.array_equal(x, y) for x, y in zip(a[3], b[3]))
    names = ["read%d/1" % i for i in range(n)]
    quals = [("I" * int(offs[i + 1] - offs[i])) if i % 2 else None for i in range(n)]
    for kw in (dict(emit_all=True, emit_fastq=False, emit_kraken=True), dict(emit_all=False, emit_fastq=True, emit_kraken=True),
               dict(emit_all=True, emit_fastq=True, emit_kraken=False), dict(emit_all=True, emit_fastq=False, emit_kraken=False),
               dict(emit_all=True, emit_fastq=False, emit_kraken=True, paired=True)):
        assert R.classify_text(Dr, Tr, bases, offs, names, 31, 31, quals=quals, **kw) == \
            oracle.classify_text(Do, To, bases, offs, names, 31, 31, quals=quals, **kw), kw
