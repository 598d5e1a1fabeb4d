#!/bin/bash
# ncu evidence for the association kernel: launch list of a short bench run + one full-set capture.  1 GPU only.
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_assoc -s 2 -c 1 -f -o gpurun_out/assoc_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu2_$TAG.log 2>&1
ls -la gpurun_out | tail -4
