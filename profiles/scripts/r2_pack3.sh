#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "host_packed" 2>&1 | tail -3
for cfg in "15 hybrid 262144" "15 pack 262144" "15 hybrid 131072" "15 pack 131072" "11 hybrid 131072" "15 hybrid 524288"; do
  set -- $cfg
  echo "== BNS_B200_HOST_PACK=$1 mode=$2 chunk=$3"
  BNS_B200_VERBOSE=1 BNS_B200_HOST_PACK=$1 BNS_B200_HOST_PACK_MODE=$2 BNS_B200_PACK_CHUNK_READS=$3 python bench.py --no-sub --no-cpu-baseline --steps 20 --e2e-steps 10 > gpurun_out/pack3_$1_$2_$3.json 2> gpurun_out/pack3_$1_$2_$3.err
  grep "classify_batch" gpurun_out/pack3_$1_$2_$3.err | tail -2
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/pack3_$1_$2_$3.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value %.0f  e2e %.1f Mreads/s  h2d %.0f MB/step (%.1f GB/s)  match %s" % (d["value"], e["value"], e["h2d_bytes_per_step"] / 1e6, e["h2d_gbs"], e["taxids_match_device_path"]))
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/pack3_$1_$2_$3.err").read()[-2000:])
PY
done
