"""Golden vectors for the call-by-call Encoder surface (assign / has_next_kmer / next_kmer / next_minimizer /
next_canonicalized_minimizer, encoder.h:201-206,594-628), produced by the UNMODIFIED reference headers behind
oracle/ref_driver.cpp (oracle/_ref/libbns_ref_v4.so, saturating cast). Run in the container that has /root/reference:

    python tests/golden/make_golden_iter.py        # writes tests/golden/golden_iter.json

Each case: the values of one call per position from the first full window on (api 2 of the driver), as hex."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402


def cases():
    rng = random.Random(77)
    out = []
    spaced = [0] * 30
    spaced[0], spaced[1] = 1, 2
    six = [0] * 30
    for i, g in ((2, 1), (7, 2), (11, 1), (15, 1), (20, 3), (25, 1)):
        six[i] = g
    configs = [(31, 0, None), (31, 50, None), (31, 33, None), (21, 40, None), (31, 0, spaced), (31, 45, spaced), (31, 0, six),
               (31, 60, six), (16, 16, None), (32, 32, None), (32, 40, None), (5, 9, [1, 0, 2, 0]), (1, 1, None), (13, 64, None)]
    for k, w, gaps in configs:
        for score in (0, 1):
            for canon in (0, 1):
                for trial in range(4):
                    L = rng.choice([0, k - 1, k, k + 3, 40, 75, 150, 151, 300])
                    alphabet = rng.choice(["ACGT", "ACGT", "ACGTN", "ACGTacgtn", "T", "AT"])
                    seq = "".join(rng.choice(alphabet) for _ in range(L))
                    if trial == 3 and L > 40:
                        p = rng.randrange(L - 36)
                        seq = seq[:p] + "T" * 36 + seq[p + 36:]
                    out.append(dict(k=k, w=w, gaps=gaps, score=score, canon=canon, seq=seq))
    return out


def main():
    R = po.load_ref("v4")
    assert R is not None, "oracle/_ref/libbns_ref_v4.so is needed (make -C oracle)"
    cs = cases()
    for c in cs:
        v = R.encode(c["seq"], c["k"], c["w"], c["gaps"], c["score"], c["canon"], po.API_ITER)
        c["values"] = ["%x" % int(x) for x in v]
    # the reference's own pins for this surface, test/encoding.cpp:17-47: first next_kmer / next_minimizer of a 34-base string
    with open(os.path.join(HERE, "golden_iter.json"), "w") as f:
        json.dump(dict(cast="saturate", cases=cs), f, separators=(",", ":"))
    print(len(cs), "cases,", sum(len(c["values"]) for c in cs), "values")


if __name__ == "__main__":
    main()
