#!/bin/bash
# the two-stage minimizer-layout kernel (bns_classify_loc.cuh): parity tests, then A/B against the lean kernel on the stress workload
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "minimizer or stress or host_packed or full_size" 2>&1 | tail -4
for pipe in 0 1; do
  for keys in 268435456 1073741824; do
    BNS_B200_LOC_PIPE=$pipe python bench.py --workload stress --stress-keys $keys --steps 13 --warmup 3 --e2e-steps 0 --check-reads 200000 > gpurun_out/locpipe_${pipe}_$keys.json 2> gpurun_out/locpipe_${pipe}_$keys.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/locpipe_${pipe}_$keys.json").read().strip().splitlines()[-1])
    print("pipe=$pipe keys=$keys: %.1f Mreads/s  kernel %.3f ms  match %s  expected-classification %s" % (d["value"], d["roofline"]["kernel_ms"], d.get("oracle_check", {}).get("taxids_match"), d["config"].get("stress_reads_classified_as_expected")))
except Exception as e:
    print("pipe=$pipe keys=$keys failed", e); print(open("gpurun_out/locpipe_${pipe}_$keys.err").read()[-1500:])
PY
  done
done
