#!/bin/bash
# A/B of prefetch variants of the two-stage kernel, then an ncu capture of it (2^28 keys)
mkdir -p gpurun_out
KEYS=268435456
for lib in bonsai_b200/libbonsai_b200.so bonsai_b200/variants/lp_nopf.so bonsai_b200/variants/lp_pf09.so bonsai_b200/variants/lp_pf01.so; do
  name=$(basename $lib .so)
  BNS_B200_LIB=$PWD/$lib python bench.py --workload stress --stress-keys $KEYS --steps 13 --warmup 3 --e2e-steps 0 --check-reads 100000 > gpurun_out/lp2_$name.json 2> gpurun_out/lp2_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/lp2_$name.json").read().strip().splitlines()[-1])
    print("$name: %.1f Mreads/s  kernel %.3f ms  match %s" % (d["value"], d["roofline"]["kernel_ms"], d.get("oracle_check", {}).get("taxids_match")))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/lp2_$name.err").read()[-1500:])
PY
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bns_classify_loc -s 3 -c 1 -o gpurun_out/prof_locpipe python bench.py --workload stress --stress-keys $KEYS --reads 4000000 --steps 3 --warmup 3 --e2e-steps 0 --check-reads 0 --no-cpu-baseline > gpurun_out/prof_locpipe.log 2>&1
tail -3 gpurun_out/prof_locpipe.log
