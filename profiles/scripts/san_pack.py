"""compute-sanitizer target: the host-buffer classify call with packing threads (packed-input kernel variants, both layouts,
single reads, mate pairs, spaced seeds) on a few thousand ragged reads; compares with the ASCII-only call.
    compute-sanitizer --tool memcheck|racecheck python profiles/scripts/san_pack.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["BNS_B200_PACK_MIN_BASES"] = "1"
os.environ["BNS_B200_PACK_CHUNK_READS"] = "600"
import helpers as H                      # noqa: E402
from bonsai_b200 import capi, dbbuild, workload as W   # noqa: E402

g = W.load_genomes()
tc, tp = W.toy_tax_arrays()
genomes = []
for gi in range(4):
    b, off = W.genome_records(g, gi)
    genomes.append((b[:150_000].copy(), np.array([0, 150_000], np.uint64)))
rb, ro, _ = H.make_reads(4000, seed=3, ragged=True)
fb, fo, _ = H.make_reads(3000, seed=4)
ok = True
for layout in ("hash", "minimizer"):
    os.environ["BNS_B200_LAYOUT"] = layout
    for mode in ("pack", "hybrid"):
        os.environ["BNS_B200_HOST_PACK_MODE"] = mode
        with capi.Context(31, 31, device=0) as bctx:
            dbbuild.build_on_device(bctx, genomes, W.GENOME_TAXIDS, tc, tp, 31, 31)
            keys, vals = bctx.table_dump()
        with capi.Context(31, 31, device=0, host_pack_threads=3) as ctx, capi.Context(31, 31, device=0, host_pack_threads=0) as plain:
            for c in (ctx, plain):
                c.load_pairs(keys, vals)
                c.load_taxonomy(tc, tp)
            for bases, offs, paired in ((rb, ro, False), (fb, fo, False), (rb, ro[:((ro.size - 1) // 2) * 2 + 1], True)):
                a = ctx.classify(bases[:int(offs[-1])], offs, paired=paired)
                b = plain.classify(bases[:int(offs[-1])], offs, paired=paired)
                same = all(np.array_equal(x, y) for x, y in zip(a, b))
                ok = ok and same
                print(layout, mode, "paired" if paired else "single", offs.size - 1, "reads:", "same" if same else "DIFFERENT", flush=True)
print("san_pack:", "ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
