#!/bin/bash
# round 2, final tree (24-warp CTAs, host packing): launch list of the bench command (our kernels only), one full ncu capture each of
# the headline kernel (config2, L2-resident hash layout), its packed-input variant (a 2^18-read chunk of the e2e call) and the stress
# kernel (2^28 keys, minimizer layout).   r2_final_prof2.sh TAG
TAG=$1
cd /root/repo; mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bns_ -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-sub --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_$TAG.log 2>&1
tail -1 gpurun_out/launches_$TAG.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bns_classify_u -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_c2 \
  python bench.py --steps 2 --warmup 1 --reads 4000000 --no-sub --no-cpu-baseline --e2e-steps 0 --check-reads 0 > gpurun_out/ncu_full_${TAG}_c2.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}_c2.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:bns_classify_u_kernel<0, true, 31, false, 0, false, false, true>' -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_c2_packed \
  env BNS_B200_HOST_PACK_MODE=pack python bench.py --steps 1 --warmup 1 --reads 4000000 --no-sub --no-cpu-baseline --e2e-steps 1 --check-reads 0 > gpurun_out/ncu_full_${TAG}_c2_packed.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}_c2_packed.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bns_classify_u -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_stress \
  python bench.py --workload stress --stress-keys 268435456 --steps 2 --warmup 1 --reads 4000000 --e2e-steps 0 --check-reads 0 > gpurun_out/ncu_full_${TAG}_stress.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}_stress.log | cut -c1-200
ls -la gpurun_out/prof_${TAG}_*.ncu-rep
