#!/bin/bash
# host packing (bns_pack.cpp + the packed-input kernel): parity tests, then the e2e of bench.py with packing off / hybrid / pack-only
mkdir -p gpurun_out
nproc; lscpu | grep -i "model name"
python -m pytest tests -m gpu -x -q -k "host_packed or classify_golden or ragged or full_size" 2>&1 | tail -5
for cfg in "0 hybrid 262144" "15 hybrid 262144" "15 pack 262144" "15 hybrid 131072" "15 hybrid 524288" "11 hybrid 262144" "7 hybrid 262144"; do
  set -- $cfg
  echo "== BNS_B200_HOST_PACK=$1 mode=$2 chunk=$3"
  BNS_B200_HOST_PACK=$1 BNS_B200_HOST_PACK_MODE=$2 BNS_B200_PACK_CHUNK_READS=$3 python bench.py --no-sub --no-cpu-baseline --steps 20 --e2e-steps 10 > gpurun_out/pack_$1_$2_$3.json 2> gpurun_out/pack_$1_$2_$3.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/pack_$1_$2_$3.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value %.0f  e2e %.1f Mreads/s  h2d %.0f MB/step (%.1f GB/s)  match %s  runs-e2e %s" % (d["value"], e["value"], e["h2d_bytes_per_step"] / 1e6, e["h2d_gbs"], e["taxids_match_device_path"], e.get("runs", {}).get("e2e_mreads_s")))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/pack_$1_$2_$3.err").read()[-2000:])
PY
done
