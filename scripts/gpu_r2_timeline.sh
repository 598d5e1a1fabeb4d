#!/bin/bash
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=5000
export KICP_LIB=$PWD/kinematic-icp_b200/lib/ab/libkicp_prof.so
{
timeout 150 python scripts/debug_timeline.py 4 flush
timeout 150 python scripts/debug_timeline.py 4
timeout 100 python scripts/debug_timeline.py 2
} 2>&1 | tee gpurun_out/r2_timeline.log
