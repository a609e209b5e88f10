#!/bin/bash
# quick GPU check: the whole GPU suite (first failure shown), optional bench of run lists.  r2_t.sh TAG [pytest -k expr]
TAG=$1; K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -40 > gpurun_out/pytest_$TAG.log; else python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_$TAG.log; fi
tail -25 gpurun_out/pytest_$TAG.log
