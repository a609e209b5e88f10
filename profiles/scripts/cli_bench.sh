set -e
cd /root/repo
python - <<'PY'
import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import helpers as H
g=H.load_genomes()
b,o,_=H.make_reads(2_000_000, seed=5, genomes=g)
with open('/tmp/reads.fq','w') as f:
    arr=b.reshape(-1,150)
    q='I'*150
    for i in range(arr.shape[0]):
        f.write('@r%d\n%s\n+\n%s\n'%(i,arr[i].tobytes().decode(),q))
open('/tmp/nodes.dmp','w').write(''.join('%d\t|\t%d\t|\trank\t|\n'%cp for cp in H.TOY_TAX))
for gi in range(4):
    bb,off=H.genome_records(g,gi)
    with open('/tmp/g%d.fa'%gi,'w') as f:
        for r in range(len(off)-1):
            f.write('>c%d\n%s\n'%(r,bb[int(off[r]):int(off[r+1])].tobytes().decode()))
PY
time ./bonsai_b200/bin/bonsai build -k 31 -w 50 -e /tmp/db.bin /tmp/nodes.dmp 11=/tmp/g0.fa 12=/tmp/g1.fa 13=/tmp/g2.fa 20=/tmp/g3.fa
ls -la /tmp/db.bin /tmp/reads.fq
time ./bonsai_b200/bin/bonsai classify -a -o /tmp/out.kraken /tmp/db.bin /tmp/nodes.dmp /tmp/reads.fq
time ./bonsai_b200/bin/bonsai classify -a -c 33554432 -o /tmp/out2.kraken /tmp/db.bin /tmp/nodes.dmp /tmp/reads.fq
cmp /tmp/out.kraken /tmp/out2.kraken && head -3 /tmp/out.kraken && wc -l /tmp/out.kraken
