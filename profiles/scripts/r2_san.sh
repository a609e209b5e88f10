#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/scripts/san_pack.py > gpurun_out/san_pack_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/san_pack_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/scripts/san_pack.py > gpurun_out/san_pack_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/san_pack_racecheck.log
