#!/bin/bash
# A/B of library variants: stress (2^28 keys, minimizer layout) and config2 (L2-resident hash layout).  r2_ab2.sh TAG lib1.so lib2.so ...
TAG=$1; shift
mkdir -p gpurun_out
for LIB in "$@"; do
  NAME=$(basename $LIB .so)
  BNS_B200_LIB=$PWD/$LIB python bench.py --workload stress --stress-keys 268435456 --steps 13 --warmup 3 --e2e-steps 0 --check-reads 100000 > gpurun_out/ab_${TAG}_s_$NAME.json 2> gpurun_out/ab_${TAG}_s_$NAME.err
  BNS_B200_LIB=$PWD/$LIB python bench.py --steps 50 --warmup 5 --e2e-steps 0 --no-sub --no-cpu-baseline --check-reads 200000 > gpurun_out/ab_${TAG}_c2_$NAME.json 2> gpurun_out/ab_${TAG}_c2_$NAME.err
  python - <<PY
import json
for w in ("s","c2"):
    try:
        d=json.loads(open("gpurun_out/ab_${TAG}_%s_$NAME.json" % w).read().strip().splitlines()[-1])
        print("$NAME", w, round(d["value"],1), "Mreads/s", d["config"]["db_layout"], "match", (d.get("oracle_check") or {}).get("taxids_match"))
    except Exception as e:
        print("$NAME", w, "failed", e); print(open("gpurun_out/ab_${TAG}_%s_$NAME.err" % w).read()[-800:])
PY
done
