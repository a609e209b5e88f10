#!/bin/bash
# round 2: `bonsai classify` end to end from files in the page cache (8 M reads plain, 2 M reads gzip / BGZF), with the
# BNS_B200_VERBOSE timers.   r2_cli.sh TAG [GPUS]
TAG=$1; GPUS=${2:-1}
cd /root/repo
mkdir -p gpurun_out
LOG=gpurun_out/cli_$TAG.log
{
nproc
bash profiles/scripts/cli_bench2.sh 8000000 2>&1 | tail -8
python - <<PY
import sys, gzip, zlib
sys.path.insert(0,'tests')
import helpers as H
data = open('/tmp/reads.fq','rb').read(2000000*315)
open('/tmp/reads2m.fq','wb').write(data)
import subprocess
open('/tmp/reads2m.fq.gz','wb').write(gzip.compress(data, 1))
open('/tmp/reads2m.fq.bgz','wb').write(H.bgzf_bytes(data))
PY
run() { local f=$1; shift; local n=$1; shift; local t0=$(date +%s.%N); BNS_B200_VERBOSE=1 ./bonsai_b200/bin/bonsai classify "$@" -o /tmp/out.txt /tmp/db.bin /tmp/nodes.dmp $f 2>&1 | grep "^\[" ; local t1=$(date +%s.%N); echo "== $n reads, $* : $(python -c "print('%.2f s  %.2f Mreads/s' % ($t1-$t0, $n/($t1-$t0)/1e6))")"; }
run /tmp/reads.fq 8000000 -a -K -c 67108864 -p 16
run /tmp/reads.fq 8000000 -a -K -c 67108864 -p 32
run /tmp/reads.fq 8000000 -a -c 67108864 -p 32
run /tmp/reads.fq 8000000 -a -f -K -c 67108864 -p 32
run /tmp/reads2m.fq 2000000 -a -c 67108864 -p 32
run /tmp/reads2m.fq.gz 2000000 -a -c 67108864 -p 32
run /tmp/reads2m.fq.bgz 2000000 -a -c 67108864 -p 32
BNS_B200_INGEST=kseq run /tmp/reads2m.fq 2000000 -a -c 67108864 -p 32
if [ "$GPUS" -gt 1 ]; then
  run /tmp/reads.fq 8000000 -a -K -c 67108864 -p 32 --gpus $GPUS
  run /tmp/reads.fq 8000000 -a -c 67108864 -p 32 --gpus $GPUS
fi
} > $LOG 2>&1
cat $LOG
