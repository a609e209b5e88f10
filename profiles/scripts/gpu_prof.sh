#!/bin/bash
# one full ncu capture of the classify kernel on the default bench: gpu_prof.sh TAG [extra bench args]
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:${KREGEX:-bns_classify_u} -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
  python bench.py --steps 2 --warmup 1 --reads 4000000 --no-cpu-baseline "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
