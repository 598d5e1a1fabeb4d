#!/bin/bash
# one full-set ncu capture of an ACTIVE association launch (launch #10 = first iteration of the 2nd registration)
mkdir -p gpurun_out
TAG=${1:-r01}
SKIP=${2:-10}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_assoc -s $SKIP -c 1 -f -o gpurun_out/assoc_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu2_$TAG.log 2>&1
ls -la gpurun_out | tail -3
