#!/bin/bash
# GPU box: parity tests, default bench, launch list and one full ncu capture of the classify kernel.
# usage: gpu_check.sh TAG [notests]
TAG=${1:-x}
mkdir -p gpurun_out
if [ "$2" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
  tail -3 gpurun_out/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --reads 4000000 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bns_classify -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
  python bench.py --steps 2 --warmup 1 --reads 4000000 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
