#!/bin/bash
# round 2: minimizer-layout iteration -- tests, stress bench at two sizes, one full ncu capture.   r2_loc.sh TAG
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
for KEYS in 268435456 1073741824; do
  python bench.py --workload stress --stress-keys $KEYS --steps 13 --warmup 3 --e2e-steps 0 --check-reads 200000 > gpurun_out/stress_${TAG}_$KEYS.json 2> gpurun_out/stress_${TAG}_$KEYS.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/stress_${TAG}_$KEYS.json").read().strip().splitlines()[-1])
print("stress $KEYS:", round(d["value"],1), "Mreads/s", d["config"]["db_layout"], "displaced", d["config"]["db_displaced"], "pbar", round(d["roofline"]["sectors_per_lookup"],4), "match", d.get("oracle_check",{}).get("taxids_match"), d["config"].get("stress_reads_classified_as_expected"))
PY
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bns_classify_u -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
  python bench.py --workload stress --stress-keys 268435456 --steps 2 --warmup 1 --reads 4000000 --e2e-steps 0 --check-reads 0 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
