#!/bin/bash
# warps per SM of the lean kernel: 8 warps x 3 CTAs (80 registers) against 7 x 4 and 9 x 3 (72 registers) and 12 x 2
mkdir -p gpurun_out
for lib in bonsai_b200/libbonsai_b200.so bonsai_b200/variants/w20c1.so bonsai_b200/variants/w16c1.so; do
  name=$(basename $lib .so)
  for wl in config2 stress; do
    BNS_B200_LIB=$PWD/$lib python bench.py --workload $wl --stress-keys 268435456 --no-sub --no-cpu-baseline --steps 20 --warmup 3 --e2e-steps 0 --check-reads 100000 > gpurun_out/occ_${name}_$wl.json 2> gpurun_out/occ_${name}_$wl.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/occ_${name}_$wl.json").read().strip().splitlines()[-1])
    print("$name $wl: %.1f Mreads/s  kernel %.3f ms  match %s" % (d["value"], d["roofline"]["kernel_ms"], (d.get("oracle_check") or {}).get("taxids_match")))
except Exception as e:
    print("$name $wl failed", e); print(open("gpurun_out/occ_${name}_$wl.err").read()[-800:])
PY
  done
done
