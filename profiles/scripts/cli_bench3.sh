#!/bin/bash
# where the CLI's time goes (BNS_B200_VERBOSE timers): uses the files cli_bench2.sh leaves in /tmp (run it first in the same call)
cd /root/repo
head -c 31500000 /tmp/reads.fq > /tmp/reads_small.fq     # 100 k reads
run() { local f=$1; shift; local n=$1; shift; local t0=$(date +%s.%N); BNS_B200_VERBOSE=1 ./bonsai_b200/bin/bonsai classify "$@" -o /tmp/out.txt /tmp/db.bin /tmp/nodes.dmp $f 2>&1 | grep "^\[" ; local t1=$(date +%s.%N); echo "== $n reads, $* : $(python -c "print('%.2f s' % ($t1-$t0))")"; }
run /tmp/reads_small.fq 100k -a -K -c 16777216 -p 16
run /tmp/reads_small.fq 100k -a -c 16777216 -p 16
run /tmp/reads.fq 8M -a -K -c 16777216 -p 16
run /tmp/reads.fq 8M -a -c 16777216 -p 16
