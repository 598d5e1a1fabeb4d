#!/bin/bash
mkdir -p gpurun_out
export KICP_SPIN_TIMEOUT_MS=5000
{
echo "== default"; for c in 4 3; do timeout 120 python scripts/debug_timing.py $c | grep -E "total|sum per pass"; done
for so in kinematic-icp_b200/lib/ab/*.so; do echo "== $so"; for c in 4 3; do KICP_LIB=$PWD/$so timeout 120 python scripts/debug_timing.py $c | grep -E "total|sum per pass"; done; done
} 2>&1 | tee gpurun_out/r2_ab.log
