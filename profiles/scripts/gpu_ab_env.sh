#!/bin/bash
# A/B of environment settings on the default bench: gpu_ab_env.sh TAG "NAME=VAL ..." "NAME=VAL ..."   ("-" = no setting)
TAG=$1; shift
mkdir -p gpurun_out
n=0
for envs in "$@"; do
  n=$((n+1))
  if [ "$envs" = "-" ]; then envs=""; fi
  env $envs timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/abe_${TAG}_$n.json 2> gpurun_out/abe_${TAG}_$n.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/abe_${TAG}_$n.json"))
    print("[$envs]", "value %.1f"%d["value"], "kernel_ms %.3f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "pbar %.4f"%d["roofline"]["sectors_per_lookup"], "tableMB %.0f"%d["config"]["db_table_mb"], "uncls", d["n_unclassified"], "match", d["e2e"]["taxids_match_device_path"])
except Exception as e:
    print("[$envs] FAILED", e); print(open("gpurun_out/abe_${TAG}_$n.err").read()[-2000:])
P
done
