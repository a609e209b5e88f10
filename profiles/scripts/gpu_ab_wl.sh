#!/bin/bash
# gpu_ab_wl.sh TAG "ENV=.. ENV=.. -- bench args" ...   one bench run per argument
TAG=$1; shift
mkdir -p gpurun_out
n=0
for spec in "$@"; do
  n=$((n+1))
  envs="${spec%%--*}"; args="${spec#*--}"
  env $envs timeout 900 python bench.py --no-cpu-baseline $args > gpurun_out/wl_${TAG}_$n.json 2> gpurun_out/wl_${TAG}_$n.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/wl_${TAG}_$n.json")); r=d["roofline"]
    print("[$spec]", d["config"].get("db_layout"), "value %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"], "kernel_ms %.3f"%r["kernel_ms"], "frac %.3f"%r["frac"], "pbar %.4f"%r["sectors_per_lookup"], "tableMB %.0f"%d["config"]["db_table_mb"], "gatherfrac %.3f"%r["frac_of_random_gather"], "uncls", d["n_unclassified"], "match", d["e2e"]["taxids_match_device_path"], d["config"].get("stress_reads_classified_as_expected"))
except Exception as e:
    print("[$spec] FAILED", e); print(open("gpurun_out/wl_${TAG}_$n.err").read()[-1500:])
P
done
