#!/bin/bash
# last tree of round 2 (SV variants): r2_final_prof3.sh c2|stress -- one full ncu capture per call (two reports exceed what a call brings back)
cd /root/repo; mkdir -p gpurun_out
if [ "$1" = "c2" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bns_ -c 400 --csv --log-file gpurun_out/launches_r02_sv.csv \
  python bench.py --steps 2 --warmup 1 --no-sub --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_r02_sv.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bns_classify_u -s 3 -c 1 -f -o gpurun_out/prof_r02sv_c2 \
  python bench.py --steps 2 --warmup 1 --reads 4000000 --no-sub --no-cpu-baseline --e2e-steps 0 --check-reads 0 > gpurun_out/ncu_full_r02sv_c2.log 2>&1
else
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bns_classify_u -s 3 -c 1 -f -o gpurun_out/prof_r02sv_stress \
  python bench.py --workload stress --stress-keys 268435456 --steps 2 --warmup 1 --reads 4000000 --e2e-steps 0 --check-reads 0 > gpurun_out/ncu_full_r02sv_stress.log 2>&1
fi
ls -la gpurun_out/prof_r02sv_*.ncu-rep
