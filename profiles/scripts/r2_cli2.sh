#!/bin/bash
# CLI timing, Kraken text and taxon only, from the 8 M-read FASTQ (cli_bench2.sh makes the inputs)   r2_cli2.sh TAG
TAG=$1
cd /root/repo; mkdir -p gpurun_out
{
bash profiles/scripts/cli_bench2.sh 8000000 2>&1 | tail -2
run() { local f=$1; shift; local n=$1; shift; local t0=$(date +%s.%N); BNS_B200_VERBOSE=1 ./bonsai_b200/bin/bonsai classify "$@" -o /tmp/out.txt /tmp/db.bin /tmp/nodes.dmp $f 2>&1 | grep "^\[" ; local t1=$(date +%s.%N); echo "== $n reads, $* : $(python -c "print('%.2f s  %.2f Mreads/s' % ($t1-$t0, $n/($t1-$t0)/1e6))")"; }
run /tmp/reads.fq 8000000 -a -c 67108864 -p 16
run /tmp/reads.fq 8000000 -a -c 67108864 -p 16
run /tmp/reads.fq 8000000 -a -K -c 67108864 -p 16
run /tmp/reads.fq 8000000 -a -f -K -c 67108864 -p 16
ls -la /tmp/out.txt
} > gpurun_out/cli_$TAG.log 2>&1
cat gpurun_out/cli_$TAG.log
