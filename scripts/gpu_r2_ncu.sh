#!/bin/bash
# full-set ncu capture of one registration launch (the 3rd k_register launch of debug_timing.py) on a workload
mkdir -p gpurun_out
TAG=${1:-r02}
CFG=${2:-4}
export KICP_SPIN_TIMEOUT_MS=60000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_register -s 2 -c 1 -f -o gpurun_out/reg_${TAG}_cfg$CFG \
    python scripts/debug_timing.py $CFG > gpurun_out/ncu_${TAG}_cfg$CFG.log 2>&1
tail -3 gpurun_out/ncu_${TAG}_cfg$CFG.log
ls -la gpurun_out/*.ncu-rep | tail -3
