#!/bin/bash
# host-only: scaling of the parallel FASTQ ingest on the GPU box's cores (the file cli_bench2.sh leaves in /tmp).  r2_reader.sh TAG
TAG=$1
cd /root/repo; mkdir -p gpurun_out
g++ -O2 -std=c++17 -o /tmp/reader_bench tests/host/reader_bench.cpp tests/host/abi_stub.cpp -lz -lpthread
[ -f /tmp/reads.fq ] || bash profiles/scripts/cli_bench2.sh 8000000 > /dev/null 2>&1
{ nproc; /tmp/reader_bench /tmp/reads.fq 67108864 1 2 4 8 16 32; /tmp/reader_bench /tmp/reads.fq 16777216 16 32; } > gpurun_out/reader_$TAG.log 2>&1
cat gpurun_out/reader_$TAG.log
