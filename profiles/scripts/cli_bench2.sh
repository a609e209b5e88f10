#!/bin/bash
# CLI throughput from a FASTQ file in the page cache: cli_bench2.sh [N reads]
set -e
N=${1:-8000000}
cd /root/repo
python - <<PY
import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import helpers as H
g=H.load_genomes()
n=$N
b,o,_=H.make_reads(n, seed=5, genomes=g)
arr=b.reshape(-1,150)
# vectorised FASTQ: "@r" + 8 hex digits + "\n" + seq + "\n+\n" + qual + "\n"
rec=np.empty((n, 2+8+1+150+3+150+1), np.uint8)
rec[:,0]=ord('@'); rec[:,1]=ord('r')
idx=np.arange(n, dtype=np.uint64)
hexd=np.frombuffer(b"0123456789abcdef", np.uint8)
for j in range(8):
    rec[:,2+j]=hexd[((idx >> np.uint64(4*(7-j))) & np.uint64(15)).astype(np.int64)]
rec[:,10]=10; rec[:,11:161]=arr; rec[:,161]=10; rec[:,162]=ord('+'); rec[:,163]=10; rec[:,164:314]=ord('I'); rec[:,314]=10
rec.tofile('/tmp/reads.fq')
open('/tmp/nodes.dmp','w').write(''.join('%d\t|\t%d\t|\trank\t|\n'%cp for cp in H.TOY_TAX))
for gi in range(4):
    bb,off=H.genome_records(g,gi)
    with open('/tmp/g%d.fa'%gi,'w') as f:
        for r in range(len(off)-1):
            f.write('>c%d\n%s\n'%(r,bb[int(off[r]):int(off[r+1])].tobytes().decode()))
PY
./bonsai_b200/bin/bonsai build -k 31 -w 50 -e /tmp/db.bin /tmp/nodes.dmp 11=/tmp/g0.fa 12=/tmp/g1.fa 13=/tmp/g2.fa 20=/tmp/g3.fa 2>/dev/null
ls -la /tmp/reads.fq | awk '{print "fastq bytes", $5}'
cat /tmp/reads.fq > /dev/null
run() { local t0=$(date +%s.%N); ./bonsai_b200/bin/bonsai classify "$@" -o /tmp/out.txt /tmp/db.bin /tmp/nodes.dmp /tmp/reads.fq 2>/dev/null; local t1=$(date +%s.%N); echo "$* : $(python -c "print('%.2f s  %.2f Mreads/s' % ($t1-$t0, $N/($t1-$t0)/1e6))")"; }
run -a
run -a -c 67108864
run -a -c 67108864 -p 16
run -a -f -K -c 67108864 -p 16
run -a -K -c 67108864 -p 16
