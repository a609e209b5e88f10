#!/bin/bash
# round 2: stash for the minimizer layout -- layout tests, config1db in both layouts, stress.   r2_stash.sh TAG
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "minimizer or layout or lookup or runs or replication or multigpu" 2>&1 | tail -25 > gpurun_out/pytest_$TAG.log; tail -12 gpurun_out/pytest_$TAG.log
run() {  # name workload [env...]
  NAME=$1; WL=$2; shift 2
  env "$@" python bench.py --workload $WL --stress-keys 268435456 --steps 20 --warmup 3 --e2e-steps 0 --check-reads 200000 --no-cpu-baseline --no-sub > gpurun_out/st_${TAG}_$NAME.json 2> gpurun_out/st_${TAG}_$NAME.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/st_${TAG}_$NAME.json").read().strip().splitlines()[-1])
    print("$NAME:", round(d["value"],1), "Mreads/s", d["config"]["db_layout"], round(d["config"]["db_table_mb"]), "MB displaced", d["config"]["db_displaced"], "pbar", round(d["roofline"]["sectors_per_lookup"],4), "match", (d.get("oracle_check") or {}).get("taxids_match"))
except Exception as e:
    print("$NAME failed", e); print(open("gpurun_out/st_${TAG}_$NAME.err").read()[-1500:])
PY
}
run c1db_hash config1db X=1
run c1db_min config1db BNS_B200_LAYOUT=minimizer
run c1db_min_half config1db BNS_B200_LAYOUT=minimizer BNS_B200_LOC_LOAD=0.625
run stress28 stress X=1
