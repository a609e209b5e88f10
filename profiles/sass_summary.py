#!/usr/bin/env python
"""Opcode census of the SASS in bonsai_b200/libbonsai_b200.so (no GPU needed): what proves the kernels are sm_100a-era
integer / memory code -- LDG.E.256 probes, LDGSTS (cp.async) staging, UBLKCP + SYNCS (TMA bulk copy onto an mbarrier),
REDUX / VIMNMX3 / SHFL warp arithmetic, 64-bit CAS inserts -- and that no tensor op exists on this path (there is none to use).

    python profiles/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bonsai_b200", "libbonsai_b200.so")
KEY = ["LDG.E.256", "LDG.E.128", "LDG.E.64", "LDGSTS", "UBLKCP", "SYNCS", "REDUX", "VIMNMX3", "VIMNMX", "SHFL", "VOTE", "MATCH", "ATOMG", "ATOM", "RED",
       "IMAD", "LOP3", "SHF", "ISETP", "SEL", "BREV", "POPC", "FLO", "PRMT", "STL", "LDL", "CCTL", "HMMA", "UTC", "LDTM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    groups = collections.OrderedDict()
    for mangled, name in zip(per, demangle):
        short = re.sub(r"\(.*", "", name).replace("void bns::", "")
        base = re.sub(r"<.*", "", short)
        g = groups.setdefault(base, {"n": 0, "insts": 0, "ops": collections.Counter(), "variants": []})
        g["n"] += 1
        g["insts"] += sum(per[mangled].values())
        g["ops"].update(per[mangled])
        g["variants"].append((short, sum(per[mangled].values())))
    print("# SASS opcode census of %s (cuobjdump -sass; sm_100a)" % os.path.relpath(LIB, ROOT))
    total = collections.Counter()
    for base, g in groups.items():
        total.update(g["ops"])
        print("\n%s: %d variant(s), %d SASS instructions in all" % (base, g["n"], g["insts"]))
        keyed = collections.Counter()
        for op, c in g["ops"].items():
            if op.startswith("LDG."):                              # by width: LDG.E.NA.ENL2.256.CONSTANT -> LDG.256
                w = re.search(r"\.(256|128|64|U8|U16|S8|S16)\b", op)
                keyed["LDG." + (w.group(1) if w else "32")] += c
                continue
            for k in KEY:
                if op.startswith(k):
                    keyed[k] += c
                    break
        print("   " + "  ".join("%s x%d" % kv for kv in sorted(keyed.items(), key=lambda kv: -kv[1])))
        if g["n"] <= 6:
            for v, c in g["variants"]:
                print("   %6d  %s" % (c, v))
    tensor = sum(c for op, c in total.items() if op.startswith(("HMMA", "UTC", "LDTM", "STTM", "HGMMA", "QGMMA", "IGMMA", "IMMA")))
    print("\ntensor-core / TMEM instructions in the library: %d (this path is 64-bit integer hashing and table probing)" % tensor)


if __name__ == "__main__":
    main()
