#!/bin/bash
# round 2: the 8-GPU evidence (run with gpurun --gpus 8): topology, multi-GPU tests, the bench at N=8 with every sub-record,
# `bonsai classify --gpus 8` against --gpus 1.   r2_8gpu.sh TAG
TAG=$1
cd /root/repo; mkdir -p gpurun_out
LOG=gpurun_out/gpu8_$TAG.log
{
nvidia-smi -L | head -8; nproc
nvidia-smi topo -m 2>/dev/null | head -14
python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench8_$TAG.json 2> gpurun_out/bench8_$TAG.err
tail -c 300 gpurun_out/bench8_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench8_$TAG.json").read().strip().splitlines()[-1])
print("N=8 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "h2d", round(d["e2e"]["h2d_gbs"],1), "of ceiling", round(d["e2e"]["h2d_ceiling_gbs"],1), "replicas_match", d.get("replicas_match"), "check", (d.get("oracle_check") or {}).get("taxids_match"), "bcast ms", d["config"]["db_broadcast_ms"])
for k in ("stress","config4","config1db"):
    r=d.get(k,{})
    print(" ", k, r.get("error") or (round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "replicas", r.get("replicas_match"), "check", r.get("taxids_match"), "bcast ms", round(r["config"]["db_broadcast_ms"],1), r["clocks"]["samples"], r["clocks"]["reasons"]))
PY
# CLI: 4 M reads
python - <<PY
import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import helpers as H
g=H.load_genomes(); n=4000000
b,o,_=H.make_reads(n, seed=5, genomes=g)
arr=b.reshape(-1,150)
rec=np.empty((n, 2+8+1+150+3+150+1), np.uint8)
rec[:,0]=ord('@'); rec[:,1]=ord('r')
idx=np.arange(n, dtype=np.uint64); hexd=np.frombuffer(b"0123456789abcdef", np.uint8)
for j in range(8): rec[:,2+j]=hexd[((idx >> np.uint64(4*(7-j))) & np.uint64(15)).astype(np.int64)]
rec[:,10]=10; rec[:,11:161]=arr; rec[:,161]=10; rec[:,162]=ord('+'); rec[:,163]=10; rec[:,164:314]=ord('I'); rec[:,314]=10
rec.tofile('/tmp/reads.fq')
open('/tmp/nodes.dmp','w').write(''.join('%d\t|\t%d\t|\trank\t|\n'%cp for cp in H.TOY_TAX))
for gi in range(4):
    bb,off=H.genome_records(g,gi)
    with open('/tmp/g%d.fa'%gi,'w') as f:
        for r in range(len(off)-1): f.write('>c%d\n%s\n'%(r,bb[int(off[r]):int(off[r+1])].tobytes().decode()))
PY
./bonsai_b200/bin/bonsai build -k 31 /tmp/db.bin /tmp/nodes.dmp 11=/tmp/g0.fa 12=/tmp/g1.fa 13=/tmp/g2.fa 20=/tmp/g3.fa 2>&1 | tail -1
run() { local t0=$(date +%s.%N); BNS_B200_VERBOSE=1 "$@" ./bonsai_b200/bin/bonsai classify -a -p 32 -c 33554432 ${GP} -o /tmp/out_$N.txt /tmp/db.bin /tmp/nodes.dmp /tmp/reads.fq 2>&1 | grep -E "^\[bonsai|Successfully" ; local t1=$(date +%s.%N); echo "== $N: $(python -c "print('%.2f s' % ($t1-$t0))")"; }
N=g1 GP="" run env X=1
N=g8nccl GP="--gpus 8" run env X=1
N=g8p2p GP="--gpus 8" run env BNS_B200_REPLICATE=p2p
md5sum /tmp/out_g1.txt /tmp/out_g8nccl.txt /tmp/out_g8p2p.txt
} > $LOG 2>&1
cat $LOG
