#!/bin/bash
# round 2: everything that needs two GPUs (run with gpurun --gpus 2) + the whole GPU suite.   r2_multi.sh TAG
TAG=$1
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -4 gpurun_out/pytest_$TAG.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench2_$TAG.json 2> gpurun_out/bench2_$TAG.err
tail -c 400 gpurun_out/bench2_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench2_$TAG.json").read().strip().splitlines()[-1])
print("N=2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "h2d ceiling", round(d["e2e"]["h2d_ceiling_gbs"],1), "replicas_match", d.get("replicas_match"), "check", (d.get("oracle_check") or {}).get("taxids_match"))
for k in ("stress","config4","config1db"):
    r=d.get(k,{})
    print(" ", k, round(r.get("value",0),1), "e2e", round((r.get("e2e") or {}).get("value",0),1), "replicas", r.get("replicas_match"), "check", r.get("taxids_match"), r.get("error"))
PY
