#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "host_packed or golden or runs or ragged or paired or full_size" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r02_wave.json 2> gpurun_out/bench_r02_wave.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r02_wave.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "ascii", round(d["e2e"]["ascii_only"]["value"],1), "runs dev", round(d["e2e"]["runs"]["device_mreads_s"],1), "runs e2e", round(d["e2e"]["runs"]["e2e_mreads_s"],1))
for k in ("stress", "config4", "config1db"):
    r = d.get(k, {}); e = r.get("e2e") or {}
    print(k, round(r.get("value",0),1), "frac", round(r.get("frac_of_random_gather",0),3), "e2e", round(e.get("value",0),1), "ascii", round((e.get("ascii_only") or {}).get("value",0),1), r.get("taxids_match"), r.get("error"))
PY
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:\(bool\)0, \(bool\)0, \(bool\)1>' -s 4 -c 1 -f -o gpurun_out/prof_r02w24_c2_packed \
  env BNS_B200_HOST_PACK_MODE=pack python bench.py --steps 1 --warmup 1 --reads 4000000 --no-sub --no-cpu-baseline --e2e-steps 1 --check-reads 0 > gpurun_out/ncu_full_r02w24_c2_packed.log 2>&1
tail -2 gpurun_out/ncu_full_r02w24_c2_packed.log | cut -c1-200
