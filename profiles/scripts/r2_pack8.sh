#!/bin/bash
# eight ranks on one box: e2e with host packing (threads per rank = hardware threads / 8 - 1) next to the ASCII-only call
mkdir -p gpurun_out
nproc; nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --no-sub --steps 50 > gpurun_out/bench_r02_pack_n8.json 2> gpurun_out/bench_r02_pack_n8.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_pack_n8.json").read().strip().splitlines()[-1])
    print("N=8 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ascii", round(d["e2e"]["ascii_only"]["value"], 1), "threads", d["e2e"]["host_pack_threads"], "h2d/step", d["e2e"]["h2d_bytes_per_step"], "ceiling", d["e2e"].get("h2d_ceiling_gbs"), "replicas", d.get("replicas_match"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r02_pack_n8.err").read()[-2000:])
PY
