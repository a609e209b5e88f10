#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --no-sub --steps 50 > gpurun_out/bench_r02_end_n4.json 2> gpurun_out/bench_r02_end_n4.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_end_n4.json").read().strip().splitlines()[-1])
    print("N=4 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ascii", round(d["e2e"]["ascii_only"]["value"], 1), "threads", d["e2e"]["host_pack_threads"], "ceiling", d["e2e"].get("h2d_ceiling_gbs"), "replicas", d.get("replicas_match"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r02_end_n4.err").read()[-2000:])
PY
