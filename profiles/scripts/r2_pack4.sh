#!/bin/bash
# packing, final check: default bench line (all sub-records) with the buffers preallocated, then two ranks on one box
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "host_packed" 2>&1 | tail -2
BNS_B200_VERBOSE=1 python bench.py > gpurun_out/bench_r02_pack_n1.json 2> gpurun_out/bench_r02_pack_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r02_pack_n1.json").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 1), "ascii", round(d["e2e"]["ascii_only"]["value"], 1), "h2d/step", d["e2e"]["h2d_bytes_per_step"], "clk", d["clocks"]["samples"])
for k in ("stress", "config4", "config1db"):
    r = d.get(k, {})
    e = r.get("e2e") or {}
    print(k, round(r.get("value", 0), 1), "e2e", round(e.get("value", 0), 1), "ascii", (e.get("ascii_only") or {}).get("value"), r.get("taxids_match"), r.get("error"))
PY
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-sub > gpurun_out/bench_r02_pack_n2.json 2> gpurun_out/bench_r02_pack_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r02_pack_n2.json").read().strip().splitlines()[-1])
print("N=2 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ascii", round(d["e2e"]["ascii_only"]["value"], 1), "threads", d["e2e"]["host_pack_threads"], "h2d/step", d["e2e"]["h2d_bytes_per_step"], "replicas", d.get("replicas_match"))
PY
fi
