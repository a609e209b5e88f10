#!/bin/bash
# where the CLI's wall time outside its own timers goes: a 1000-read input (start-up + exit only), loader statistics
cd /root/repo; mkdir -p gpurun_out
python - <<PY
import sys
sys.path.insert(0,'tests')
import helpers as H, numpy as np
g=H.load_genomes()
b,o,_=H.make_reads(1000, seed=5, genomes=g)
with open('/tmp/small.fq','w') as f:
    for i in range(1000):
        s=b[int(o[i]):int(o[i+1])].tobytes().decode(); f.write('@r%d\n%s\n+\n%s\n'%(i,s,'I'*len(s)))
open('/tmp/nodes.dmp','w').write(''.join('%d\t|\t%d\t|\trank\t|\n'%cp for cp in H.TOY_TAX))
for gi in range(4):
    bb,off=H.genome_records(g,gi)
    with open('/tmp/g%d.fa'%gi,'w') as f:
        for r in range(len(off)-1):
            f.write('>c%d\n%s\n'%(r,bb[int(off[r]):int(off[r+1])].tobytes().decode()))
PY
./bonsai_b200/bin/bonsai build -k 31 -w 50 -e /tmp/db.bin /tmp/nodes.dmp 11=/tmp/g0.fa 12=/tmp/g1.fa 13=/tmp/g2.fa 20=/tmp/g3.fa 2>/dev/null
for i in 1 2 3; do
  t0=$(date +%s.%N); BNS_B200_VERBOSE=1 ./bonsai_b200/bin/bonsai classify -a -p 16 -o /tmp/out.txt /tmp/db.bin /tmp/nodes.dmp /tmp/small.fq 2>&1 | grep "bonsai classify\]"; t1=$(date +%s.%N)
  python -c "print('wall %.3f s' % ($t1-$t0))"
done
LD_DEBUG=statistics ./bonsai_b200/bin/bonsai classify -a -p 16 -o /tmp/out.txt /tmp/db.bin /tmp/nodes.dmp /tmp/small.fq 2>&1 | grep -i "total startup time\|relocation\|load" | head -8
t0=$(date +%s.%N); ./bonsai_b200/bin/bonsai hist /tmp/db.bin > /dev/null 2>&1; t1=$(date +%s.%N); python -c "print('bonsai hist (no device): wall %.3f s' % ($t1-$t0))"
