#!/bin/bash
# round 2: minimizer-layout iteration -- tests of the layout, stress at 2^28 with two table loads, one full ncu capture.   r2_loc2.sh TAG
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "minimizer or lookup or layout or classify_golden or replication or build" 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
run() {  # name keys [env...]
  NAME=$1; KEYS=$2; shift 2
  env "$@" python bench.py --workload stress --stress-keys $KEYS --steps 13 --warmup 3 --e2e-steps 0 --check-reads 200000 > gpurun_out/stress_${TAG}_$NAME.json 2> gpurun_out/stress_${TAG}_$NAME.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/stress_${TAG}_$NAME.json").read().strip().splitlines()[-1])
    print("$NAME $KEYS:", round(d["value"],1), "Mreads/s", d["config"]["db_layout"], round(d["config"]["db_table_mb"]), "MB displaced", d["config"]["db_displaced"], "pbar", round(d["roofline"]["sectors_per_lookup"],4), "match", d.get("oracle_check",{}).get("taxids_match"), d["config"].get("stress_reads_classified_as_expected"))
except Exception as e:
    print("$NAME failed", e); print(open("gpurun_out/stress_${TAG}_$NAME.err").read()[-1500:])
PY
}
run base28 268435456 X=1
run half28 268435456 BNS_B200_LOC_LOAD=0.5
run base30 1073741824 X=1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bns_classify_u -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
  python bench.py --workload stress --stress-keys 268435456 --steps 2 --warmup 1 --reads 4000000 --e2e-steps 0 --check-reads 0 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
