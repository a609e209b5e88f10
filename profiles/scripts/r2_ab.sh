#!/bin/bash
# A/B of library variants on the stress workload: r2_ab.sh TAG KEYS lib1.so lib2.so ...
TAG=$1; KEYS=$2; shift 2
mkdir -p gpurun_out
for LIB in "$@"; do
  NAME=$(basename $LIB .so)
  BNS_B200_LIB=$PWD/$LIB python bench.py --workload stress --stress-keys $KEYS --steps 13 --warmup 3 --e2e-steps 0 --check-reads 100000 > gpurun_out/ab_${TAG}_$NAME.json 2> gpurun_out/ab_${TAG}_$NAME.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${TAG}_$NAME.json").read().strip().splitlines()[-1])
    print("$NAME $KEYS:", round(d["value"],1), "Mreads/s", d["config"]["db_layout"], "displaced", d["config"]["db_displaced"], "pbar", round(d["roofline"]["sectors_per_lookup"],4), "match", d.get("oracle_check",{}).get("taxids_match"), d["config"].get("stress_reads_classified_as_expected"))
except Exception as e:
    print("$NAME failed", e); print(open("gpurun_out/ab_${TAG}_$NAME.err").read()[-800:])
PY
done
