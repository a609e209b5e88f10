#!/bin/bash
# A/B of library variants on the default bench: gpu_ab.sh TAG lib1 lib2 ...   ("default" = in-tree library)
TAG=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename $lib .so)
  if [ "$lib" = "default" ]; then unset BNS_B200_LIB; else export BNS_B200_LIB=$PWD/$lib; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${TAG}_$name.json 2> gpurun_out/ab_${TAG}_$name.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/ab_${TAG}_$name.json"))
    print("$name", "value %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"], "kernel_ms %.3f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "uncls", d["n_unclassified"], "match", d["e2e"]["taxids_match_device_path"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/ab_${TAG}_$name.err").read()[-2000:])
P
done
