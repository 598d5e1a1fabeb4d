#!/bin/bash
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=5000
for c in 4; do KICP_LIB=$PWD/kinematic-icp_b200/lib/ab/libkicp_prof.so timeout 120 python scripts/debug_prof.py $c; done 2>&1 | tee gpurun_out/r2_prof.log
