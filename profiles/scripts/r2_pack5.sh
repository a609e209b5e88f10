#!/bin/bash
# staging written with / without non-temporal stores, chunk sizes: e2e of the headline workload
mkdir -p gpurun_out
for cfg in "x 262144" "1 262144" "1 131072" "x 131072" "1 65536"; do
  set -- $cfg
  if [ "$1" = "1" ]; then export BNS_B200_PACK_NO_NT=1; else unset BNS_B200_PACK_NO_NT; fi
  BNS_B200_PACK_CHUNK_READS=$2 python bench.py --no-sub --no-cpu-baseline --steps 20 --e2e-steps 10 > gpurun_out/pack5_$1_$2.json 2> gpurun_out/pack5_$1_$2.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/pack5_$1_$2.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("no_nt=$1 chunk=$2: e2e %.1f Mreads/s  h2d %.0f MB/step (%.1f GB/s)  ascii-only %.1f  match %s" % (e["value"], e["h2d_bytes_per_step"] / 1e6, e["h2d_gbs"], e["ascii_only"]["value"], e["taxids_match_device_path"]))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/pack5_$1_$2.err").read()[-1500:])
PY
done
